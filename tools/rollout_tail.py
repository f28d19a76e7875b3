import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from rrt_mpc_b200 import MPCConfig, SolverSettings, TrajectoryTracker
B, T = 2048, 500
d = np.load("tests/golden/default_scenario.npz")
path = np.array(d["path"])
rng = np.random.default_rng(4)
noise = rng.normal(size=(B,) + path.shape) * 0.15
noise[:, 0] = 0.0
paths = [path + noise[b] for b in range(B)]
starts = path[0] + rng.normal(size=(B, 2)) * 0.5
goals = np.full((B, 2), 1e9)
tr = TrajectoryTracker(MPCConfig(sim_steps=T), None, settings=SolverSettings(eps_abs=1e-6, eps_rel=1e-6, polish_passes=5, polish_retry=2, early_polish=False))
res = tr.track_batch(paths, starts, goals, map_resolution=0.8, warm_start=True)
it = res.step_iters
print("iters by step-index decile (mean / p99 / max):")
for a in range(0, T, 50):
    blk = it[:, a:a+50]
    print(a, round(blk.mean(),1), int(np.percentile(blk,99)), int(blk.max()))
tot = it.sum(axis=1); order = np.argsort(-tot)
print("iterations per vehicle: mean", tot.mean(), "p50", np.percentile(tot,50), "p99", np.percentile(tot,99), "max", tot.max(), "top5", tot[order[:5]].tolist())
b0 = order[0]
print("slowest vehicle", b0, "iters per step (first 130):", it[b0,:130].tolist())
slow = np.argwhere(it > 5000)
print("slow steps:", len(slow), "vehicles:", len(set(slow[:,0])), "first few:", slow[:10].tolist())
b, s = slow[0] if len(slow) else (0,0)
print("vehicle", b, "iters around:", it[b, max(0,s-3):s+5].tolist())
st = res.states[b]
print("state at slow step:", st[s-1] if s>0 else None, "speed", st[s-1][3] if s>0 else None)
print("mean speed by decile:", [round(float(np.nanmean(res.states[:, a:a+50, 3])),3) for a in range(0,T,50)])
