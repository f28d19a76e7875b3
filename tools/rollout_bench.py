"""BASELINE.json configs[3]: closed-loop rollout of 8,192 vehicles along perturbed copies of the default RRT* path,
500 steps, per-step relinearisation, warm start, all on the device (cudampc_rollout_batch).  Prints one JSON line.
usage: python tools/rollout_bench.py [vehicles] [steps]"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from rrt_mpc_b200 import MPCConfig, SolverSettings, TrajectoryTracker

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
d = np.load("tests/golden/default_scenario.npz")
path = np.array(d["path"])
rng = np.random.default_rng(4)
paths, starts = [], []
for b in range(B):
    pts = path + rng.normal(size=path.shape) * 0.15
    pts[0] = path[0]
    paths.append([tuple(p) for p in pts]); starts.append(pts[0] + rng.normal(size=2) * 0.5)
out = {}
for name, goals, early in (("stop_at_goal", np.tile(d["goal"], (B, 1)), True), ("all_steps", np.full((B, 2), 1e9), True),
                           ("all_steps_osqp_literal", np.full((B, 2), 1e9), False)):
    tr = TrajectoryTracker(MPCConfig(sim_steps=T), None, settings=SolverSettings(eps_abs=1e-6, eps_rel=1e-6, polish_passes=5, polish_retry=2, early_polish=early))
    tr.track_batch(paths[:256], starts[:256], goals[:256], map_resolution=0.8, warm_start=True)        # warm-up
    torch.cuda.synchronize()
    t = time.perf_counter()
    res = tr.track_batch(paths, starts, goals, map_resolution=0.8, warm_start=True)
    dt = time.perf_counter() - t
    steps = int(res.n_steps.sum())
    out[name] = {"vehicles": B, "sim_steps": T, "wall_s": dt, "vehicle_steps": steps, "vehicle_steps_per_s": steps / dt,
                 "mean_steps_per_vehicle": float(res.n_steps.mean()), "goal_reached_frac": float(res.goal_reached.mean()),
                 "aborted": int(res.aborted.sum()), "mean_iters_per_step": float(res.step_iters[res.step_status != 0].mean())}
print(json.dumps(out))
