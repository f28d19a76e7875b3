"""Small single-launch case for ncu (dev tool): python tools/prof_case.py N B eps"""
import sys, dataclasses, os
import numpy as np
sys.path.insert(0, ".")
import torch
from rrt_mpc_b200 import MPCController, SolverSettings, MPCConfig
from rrt_mpc_b200.synthetic import make_batch
N, B, eps = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
par = MPCConfig(horizon=N).to_parameters(0.8)
if N == 50:
    par = dataclasses.replace(par, du_bounds=((-12., 12.), (-0.02, 0.02)))
x0, ref, up = make_batch(B, N, seed=3 if N == 50 else 2)
ctl = MPCController(par, SolverSettings(eps_abs=eps, eps_rel=eps, polish_passes=5, polish_retry=2, keep_iterate=bool(int(os.environ.get('KEEP','0'))), early_polish=bool(int(os.environ.get('EARLY','1'))),
                                         check_termination=int(os.environ.get('CHECK', '25'))), max_batch=B)
d = lambda a: torch.as_tensor(a).cuda()
dx0, dref, dup = d(x0), d(ref), d(up)
for _ in range(2):
    r = ctl.solve_batch(dx0, dref, u_prev=dup)
torch.cuda.synchronize()
print("iters mean", r.iters.double().mean().item(), "solved", int((r.status == 1).sum()))
