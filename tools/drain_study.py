"""Would a difficulty predictor + longest-first ordering shrink the drain of the last wave?  (dev study)
Solves the bench batch, then simulates the persistent-kernel schedule (740 slots pulling problems in a given order, cost = iterations)
for the natural order, the oracle order (sorted by true iterations) and orders given by cheap input features."""
import sys, dataclasses, heapq
import numpy as np
sys.path.insert(0, ".")
import torch
from rrt_mpc_b200 import MPCController, SolverSettings, MPCConfig
from rrt_mpc_b200.synthetic import make_batch
N, B = 50, 65536
par = dataclasses.replace(MPCConfig(horizon=N).to_parameters(0.8), du_bounds=((-12., 12.), (-0.02, 0.02)))
x0, ref, up = make_batch(B, N, seed=3)
for early in (True, False):
    ctl = MPCController(par, SolverSettings(eps_abs=1e-6, eps_rel=1e-6, polish_passes=5, polish_retry=4, early_polish=early, keep_iterate=False), max_batch=B)
    r = ctl.solve_batch(torch.as_tensor(x0).cuda(), torch.as_tensor(ref).cuda(), u_prev=torch.as_tensor(up).cuda())
    it = r.iters.cpu().numpy().astype(np.float64) + 20.0 * r.info.cpu().numpy()[:, 1]       # + ~20 iterations' worth per factorisation / polish pass
    def makespan(order, slots=740):
        h = [0.0] * slots
        heapq.heapify(h)
        for b in order:
            t = heapq.heappop(h); heapq.heappush(h, t + it[b])
        return max(h)
    ideal = it.sum() / 740
    yaw = np.unwrap(ref[:, :, 2], axis=1)
    feats = {
        "heading error |yaw0 - x0.yaw|": np.abs(np.angle(np.exp(1j * (ref[:, 0, 2] - x0[:, 2])))),
        "position error |x0 - ref0|": np.hypot(x0[:, 0] - ref[:, 0, 0], x0[:, 1] - ref[:, 0, 1]),
        "total turning sum|dyaw|": np.abs(np.diff(yaw, axis=1)).sum(axis=1),
        "max |dyaw|": np.abs(np.diff(yaw, axis=1)).max(axis=1),
        "speed error |v0 - vref0|": np.abs(x0[:, 3] - ref[:, 0, 3]),
        "|u_prev steer|": np.abs(up[:, 1]),
    }
    print(f"early={early}: mean iters-equivalent {it.mean():.1f}, max {it.max():.0f}; ideal {ideal:.0f}; natural order {makespan(range(B)) / ideal:.4f} x ideal; "
          f"oracle longest-first {makespan(np.argsort(-it)) / ideal:.4f}")
    for name, f in feats.items():
        c = np.corrcoef(f, it)[0, 1]
        print(f"   {name:32s} corr {c:+.3f}   longest-first by it: {makespan(np.argsort(-f)) / ideal:.4f} x ideal")
