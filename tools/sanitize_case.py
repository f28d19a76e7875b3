"""Tiny end-to-end case for compute-sanitizer (dev tool)."""
import sys, dataclasses
import numpy as np
sys.path.insert(0, ".")
from rrt_mpc_b200 import MPCController, SolverSettings, MPCConfig, TrajectoryTracker
from rrt_mpc_b200.synthetic import make_batch
for N, B in ((20, 24), (50, 12), (5, 3)):
    par = MPCConfig(horizon=N).to_parameters(0.8)
    x0, ref, up = make_batch(B, N, seed=2)
    ctl = MPCController(par, SolverSettings(eps_abs=1e-4, eps_rel=1e-4, polish_passes=3, polish_retry=1, early_polish=True), max_batch=B)
    r = ctl.solve_batch(x0, ref, u_prev=up)
    print(N, B, r.status.tolist()[:4], r.iters.mean())
    A, Bm, c = ctl.linearize_batch(ref)
d = np.load("tests/golden/default_scenario.npz")
tr = TrajectoryTracker(MPCConfig(sim_steps=6), None, settings=SolverSettings(eps_abs=1e-3, eps_rel=1e-3))
res = tr.track_batch([[tuple(p) for p in d["path"]]] * 3, [d["start"]] * 3, [d["goal"]] * 3, map_resolution=0.8)
print("rollout steps", res.n_steps.tolist())
