"""Seeded synthetic tracking problems of the BASELINE.json shapes.

Config 2: 4,096 problems, horizon 20, default limits.            ``make_batch(4096, 20, seed=2)``
Config 3: 65,536 problems, horizon 50, steering-rate +-0.02.       ``make_batch(65536, 50, seed=3)``
Config 5: 1,048,576 problems, horizon 50 (seed 5), sharded by rank.

Each problem: a smooth random-curvature path -> ``build_reference`` (the reference's own producer,
mirrored in ref_builder.py) -> a random window of N+1 rows; ``x0 = ref[0] + noise``; random ``u_prev``.
Paths come from a seeded pool (``pool`` of them) so that a million problems generate in seconds;
window start, state noise and ``u_prev`` are drawn per problem.
"""
from __future__ import annotations

import numpy as np

from .ref_builder import build_reference


def random_path(rng: np.random.Generator, length_px: float, ds: float = 1.0) -> np.ndarray:
    """Smooth path: heading integrates a slowly varying random curvature (arcs / S-curves)."""
    n = int(length_px / ds) + 1
    knots = max(3, n // 40 + 2)
    kappa_knots = rng.uniform(-0.035, 0.035, size=knots)          # 1/px  (radius >= ~28 px)
    kappa = np.interp(np.linspace(0, knots - 1, n), np.arange(knots), kappa_knots)
    yaw = rng.uniform(-np.pi, np.pi) + np.cumsum(kappa) * ds
    x = 100.0 + np.cumsum(np.cos(yaw)) * ds
    y = 100.0 + np.cumsum(np.sin(yaw)) * ds
    return np.column_stack((x, y))


def make_batch(batch: int, horizon: int, seed: int, *, v_px_s: float = 15.0, dt: float = 0.1,
               pool: int = 256, start: int = 0, count: int | None = None):
    """Return ``x0 (B,4)``, ``ref (B,N+1,4)``, ``u_prev (B,2)`` for problems ``start .. start+count``
    of the seeded batch (so that ranks can generate disjoint shards of one global batch)."""
    count = batch - start if count is None else count
    rng = np.random.default_rng(seed)
    need = 2.0 * (horizon + 1)
    refs = []
    for _ in range(pool):
        length = rng.uniform(max(60.0, 1.5 * need), max(120.0, 2.5 * need))
        refs.append(build_reference(random_path(rng, length), v_px_s, horizon, dt))
    # per-problem draws come from a counter-based stream so that shards agree with the full batch
    x0 = np.empty((count, 4)); ref = np.empty((count, horizon + 1, 4)); u_prev = np.empty((count, 2))
    chunk = 4096
    for c0 in range((start // chunk) * chunk, start + count, chunk):
        crng = np.random.default_rng([seed, c0 // chunk])
        pid = crng.integers(0, pool, size=chunk)
        frac = crng.random(chunk)
        noise = crng.normal(size=(chunk, 4)) * np.array([1.0, 1.0, 0.1, 2.0])
        up = crng.uniform([-5.0, -0.2], [5.0, 0.2], size=(chunk, 2))
        lo, hi = max(c0, start), min(c0 + chunk, start + count)
        for g in range(lo, hi):
            j = g - c0
            r = refs[pid[j]]
            s0 = int(frac[j] * max(1, len(r) - (horizon + 1) + 1))
            s0 = min(s0, len(r) - (horizon + 1))
            w = r[s0:s0 + horizon + 1]
            ref[g - start] = w
            x0[g - start] = w[0] + noise[j]
            u_prev[g - start] = up[j]
    return x0, ref, u_prev
