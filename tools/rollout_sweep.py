import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
from rrt_mpc_b200 import MPCConfig, SolverSettings, TrajectoryTracker
B, T = 8192, 500
d = np.load("tests/golden/default_scenario.npz")
path = np.array(d["path"])
rng = np.random.default_rng(4)
noise = rng.normal(size=(B,) + path.shape) * 0.15
noise[:, 0] = 0.0
paths = [path + noise[b] for b in range(B)]
starts = path[0] + rng.normal(size=(B, 2)) * 0.5
goals = np.full((B, 2), 1e9)
for retry, eps, early in ((4, 1e-6, False), (2, 1e-6, False), (0, 1e-6, False), (2, 1e-6, True), (2, 1e-3, False)):
    tr = TrajectoryTracker(MPCConfig(sim_steps=T), None, settings=SolverSettings(eps_abs=eps, eps_rel=eps, polish_passes=5, polish_retry=retry, early_polish=early))
    tr.track_batch(paths[:512], starts[:512], goals[:512], map_resolution=0.8, warm_start=True, sim_steps=50)
    torch.cuda.synchronize()
    t = time.perf_counter()
    res = tr.track_batch(paths, starts, goals, map_resolution=0.8, warm_start=True, record_step_time=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    it = res.step_iters[res.step_status != 0]
    ns = np.asarray(res.step_ns, dtype=np.float64); ns = ns[ns > 0] / 1e3
    print(f"retry={retry} eps={eps} early={early}: {int(res.n_steps.sum())/dt:.0f} steps/s, wall {dt:.2f}s, iters mean {it.mean():.1f} p99 {np.percentile(it,99):.0f} max {it.max()}, "
          f"lat us p50 {np.percentile(ns,50):.0f} p99 {np.percentile(ns,99):.0f} max {ns.max():.0f}, status!=1: {(res.step_status[res.step_status!=0]!=1).sum()}, aborted {int(res.aborted.sum())}")
