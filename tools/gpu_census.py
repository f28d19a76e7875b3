"""Iteration / solve census of a batch (dev tool): python tools/gpu_census.py N B [early]"""
import sys, dataclasses, os
import numpy as np
sys.path.insert(0, ".")
import torch
from rrt_mpc_b200 import MPCController, SolverSettings, MPCConfig
from rrt_mpc_b200.synthetic import make_batch
N, B = int(sys.argv[1]), int(sys.argv[2]); early = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
par = MPCConfig(horizon=N).to_parameters(0.8)
if N == 50: par = dataclasses.replace(par, du_bounds=((-12., 12.), (-0.02, 0.02)))
x0, ref, up = make_batch(B, N, seed=3 if N == 50 else 2)
kw = dict(eps_abs=1e-6, eps_rel=1e-6, polish_passes=5, polish_retry=2, early_polish=int(sys.argv[3]) if len(sys.argv) > 3 else 1)
for k, v in os.environ.items():
    if k.startswith("SET_"): kw[k[4:].lower()] = type(getattr(SolverSettings(), k[4:].lower()))(float(v))
ctl = MPCController(par, SolverSettings(**kw), max_batch=B)
d = lambda a: torch.as_tensor(a).cuda()
r = ctl.solve_batch(d(x0), d(ref), u_prev=d(up)); torch.cuda.synchronize()
it = r.iters.cpu().numpy(); info = r.info.cpu().numpy()
print(f"settings {kw}")
print(f"iters mean {it.mean():.1f} median {np.median(it):.0f} p90 {np.percentile(it, 90):.0f} p99 {np.percentile(it, 99):.0f} max {it.max()}")
print(f"solves mean {info[:,3].mean():.1f}  (overhead over iters {info[:,3].mean()/it.mean()-1:.1%}); factorisations mean {info[:,1].mean():.2f}; rho updates {info[:,0].mean():.2f}; accepted polish passes {info[:,2].mean():.2f}")
h, e = np.histogram(it, bins=[0, 50, 75, 100, 125, 150, 200, 250, 300, 400, 500, 750, 1000, 2000, 100000])
tot = it.sum()
for lo, hi, c in zip(e[:-1], e[1:], h):
    sel = (it >= lo) & (it < hi)
    print(f"  iters [{lo:5d},{hi:6d}): {c:6d} problems ({c/B:6.1%}), {it[sel].sum()/tot:6.1%} of all iterations")
print("status", dict(zip(*np.unique(r.status.cpu().numpy(), return_counts=True))))
