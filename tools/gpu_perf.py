"""Device-path timing (dev tool): python tools/gpu_perf.py N B [eps]"""
import sys, dataclasses, os
import numpy as np
sys.path.insert(0, ".")
import torch
import rrt_mpc_b200._lib as _L
if os.environ.get('CUDAMPC_LIB'): _L.LIB_PATH = os.path.abspath(os.environ['CUDAMPC_LIB'])
from rrt_mpc_b200 import MPCController, SolverSettings, MPCConfig
from rrt_mpc_b200.synthetic import make_batch
N, B = int(sys.argv[1]), int(sys.argv[2]); eps = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
par = MPCConfig(horizon=N).to_parameters(0.8)
if N == 50: par = dataclasses.replace(par, du_bounds=((-12., 12.), (-0.02, 0.02)))
x0, ref, up = make_batch(B, N, seed=3 if N == 50 else 2)
kw = dict(eps_abs=eps, eps_rel=eps, polish_passes=5, polish_retry=2, early_polish=bool(int(os.environ.get('EARLY','1'))))
for k, v in os.environ.items():     # SET_<FIELD>=value overrides a SolverSettings field
    if k.startswith("SET_"): kw[k[4:].lower()] = int(v) if k == "SET_EARLY_POLISH" else type(getattr(SolverSettings(), k[4:].lower()))(float(v))
ctl = MPCController(par, SolverSettings(**kw), max_batch=B)
d = lambda a: torch.as_tensor(a).cuda()
dx0, dref, dup = d(x0), d(ref), d(up)
r = ctl.solve_batch(dx0, dref, u_prev=dup); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = ctl.solve_batch(dx0, dref, u_prev=dup); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
import subprocess
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
it = r.iters.double().mean().item()
cyc = best * 1e-3 / (B * it) * 148 * 1.8e9
print(f"[{" ".join(k[4:].lower()+"="+v for k,v in os.environ.items() if k.startswith("SET_")) or "default"}] N={N} B={B} per_sm={ctl.problems_per_sm()}: {best:.2f} ms -> {B/best*1e3:.0f} solves/s, mean iters {it:.1f}, "
      f"{cyc:.0f} SM-cycles per problem-iteration, solved {(r.status==1).sum().item()} | {clk}")
