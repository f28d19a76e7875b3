"""Quick GPU sanity run (dev tool): solve a few problems through the C ABI and compare with the numpy oracle."""
import sys, time, dataclasses
import numpy as np
sys.path.insert(0, ".")
import torch
from rrt_mpc_b200 import MPCController, MPCParameters, SolverSettings, MPCConfig
from rrt_mpc_b200.synthetic import make_batch
from oracle import mpc_numpy as M

def run(N, B, ncheck, eps=1e-6, passes=3):
    cfg = MPCConfig(horizon=N)
    par = cfg.to_parameters(0.8)
    op = M.Params(horizon=N)
    if N == 50:
        par = dataclasses.replace(par, du_bounds=((-12., 12.), (-0.02, 0.02)))
        op = dataclasses.replace(op, du_bounds=((-12., 12.), (-0.02, 0.02)))
    x0, ref, up = make_batch(B, N, seed=3 if N == 50 else 2)
    ctl = MPCController(par, SolverSettings(eps_abs=eps, eps_rel=eps, polish_passes=passes), max_batch=B)
    t = time.time(); res = ctl.solve_batch(x0, ref, u_prev=up); t1 = time.time() - t
    t = time.time(); res = ctl.solve_batch(x0, ref, u_prev=up); t2 = time.time() - t
    print(f"N={N} B={B}: host-path first {t1*1e3:.1f} ms, second {t2*1e3:.1f} ms -> {B/t2:.0f} solves/s; per_sm={ctl.problems_per_sm()} ws={ctl.workspace_doubles()}")
    print("  status counts", dict(zip(*np.unique(res.status, return_counts=True))), "iters mean %.1f max %d" % (res.iters.mean(), res.iters.max()),
          "polish passes", dict(zip(*np.unique(res.info[:, 2], return_counts=True))))
    worst = 0
    for b in range(ncheck):
        u0n, Xn, Un, _ = M.solve_kkt_newton(x0[b], ref[b], up[b], op)
        e = np.abs(res.u0[b] - u0n).max(); worst = max(worst, e)
        ex = np.abs(res.Xp[b] - Xn).max()
        if b < 4 or e > 1e-6:
            print(f"  b={b} it={res.iters[b]} st={res.status[b]} info={res.info[b]} u0 err {e:.2e} X err {ex:.2e} pri {res.pri_res[b]:.1e} dua {res.dua_res[b]:.1e}")
    print("  worst u0 err vs KKT-Newton over", ncheck, ":", worst)
    # device path timing
    d = lambda a: torch.as_tensor(a).cuda()
    dx0, dref, dup = d(x0), d(ref), d(up)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r2 = ctl.solve_batch(dx0, dref, u_prev=dup); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    it = r2.iters.double().mean().item()
    print(f"  device path: {ms:.2f} ms -> {B/ms*1e3:.0f} solves/s, mean iters {it:.1f}, {ms*1e3/B/it*1e3:.1f} ns per problem-iteration (chip-wide)")
    assert np.array_equal(r2.status.cpu().numpy(), res.status)
    assert np.abs(r2.u0.cpu().numpy() - res.u0).max() == 0.0

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run(20, 4096, 24)
    run(50, 4096, 12)
    run(5, 8, 8, eps=1e-3, passes=1)
