"""Accuracy census at full size (dev tool): how many problems end unpolished / with residual, and their u0 error."""
import sys, dataclasses
import numpy as np
sys.path.insert(0, ".")
import torch
from rrt_mpc_b200 import MPCController, SolverSettings, MPCConfig
from rrt_mpc_b200.synthetic import make_batch
from oracle import mpc_numpy as O
N, B = int(sys.argv[1]), int(sys.argv[2])
du = 0.02 if N == 50 else 0.15
par = dataclasses.replace(MPCConfig(horizon=N).to_parameters(0.8), du_bounds=((-12., 12.), (-du, du)))
op = dataclasses.replace(O.Params(horizon=N), du_bounds=((-12., 12.), (-du, du)))
x0, ref, up = make_batch(B, N, seed=3 if N == 50 else 2)
d = lambda a: torch.as_tensor(a).cuda()
dx0, dref, dup = d(x0), d(ref), d(up)
for passes in (3, 6, 12):
    ctl = MPCController(par, SolverSettings(eps_abs=1e-6, eps_rel=1e-6, polish_passes=passes), max_batch=B)
    r = ctl.solve_batch(dx0, dref, u_prev=dup); torch.cuda.synchronize()
    info = r.info.cpu().numpy(); pri = r.pri_res.cpu().numpy(); dua = r.dua_res.cpu().numpy(); u0 = r.u0.cpu().numpy(); it = r.iters.cpu().numpy()
    npol = info[:, 2]
    print(f"passes={passes}: accepted-pass histogram {dict(zip(*np.unique(npol, return_counts=True)))}; pri>1e-8: {(pri > 1e-8).sum()}, dua>1e-7: {(dua > 1e-7).sum()}")
    bad = np.flatnonzero((npol == 0) | (pri > 1e-8))
    rng = np.random.default_rng(0)
    errs = []
    for b in rng.choice(bad, min(40, len(bad)), replace=False):
        u0n, *_ = O.solve_kkt_newton(x0[b], ref[b], up[b], op)
        errs.append((np.abs(u0[b] - u0n).max(), int(npol[b]), pri[b], dua[b], int(it[b]), int(b)))
    errs.sort(reverse=True)
    print("  worst of sampled bad:", [(f"{e:.1e}", n, f"{p:.1e}", f"{q:.1e}", i, b) for e, n, p, q, i, b in errs[:8]])
    print("  sampled bad: frac with u0 err > 1e-5:", np.mean([e[0] > 1e-5 for e in errs]), " median err", np.median([e[0] for e in errs]))
