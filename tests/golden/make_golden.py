"""Generate the golden fixtures of tests/golden/ (run in the BUILD container, where /root/reference exists).

What comes from the REAL reference (imported through a 3-line shim because the reference's src/__init__.py
eagerly imports cvxpy, which is not installable offline):
  * linearize / f_discrete outputs            <- src/control/vehicle_model.py
  * build_reference outputs                   <- src/control/ref_builder.py + src/common/geometry.py
  * the default-config scenario (map, RRT* path, ref_global, start, goal)
                                              <- src/maps/*, src/planning/rrt_star.py with src/config.py defaults
What does NOT (the reference's solver stack, cvxpy -> osqp, is absent): the QP optimum.  For those cases the
fixture holds the unique minimiser certified by oracle.mpc_numpy.solve_kkt_newton (an active-set Newton method
that shares no code with the ADMM paths) - see DESIGN.md "parity".

usage: python tests/golden/make_golden.py
"""
import dataclasses
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

shim = types.ModuleType("src")
shim.__path__ = [os.path.join(REF, "src")]
sys.modules["src"] = shim
from src.control.vehicle_model import f_discrete, linearize          # noqa: E402
from src.control.ref_builder import build_reference                  # noqa: E402
from src.maps.generator import MapGenerator                          # noqa: E402
from src.maps.inflate import inflate_grayscale_map, to_occupancy_grid  # noqa: E402
from src.planning.rrt_star import PlannerParameters, RRTStarPlanner  # noqa: E402

from oracle import mpc_numpy as O                                    # noqa: E402
from rrt_mpc_b200.synthetic import make_batch, random_path           # noqa: E402


def golden_vehicle_model():
    rng = np.random.default_rng(11)
    n = 256
    x = np.column_stack((rng.uniform(-300, 300, n), rng.uniform(-300, 300, n), rng.uniform(-7, 7, n), rng.uniform(-5, 40, n)))
    u = np.column_stack((rng.uniform(-35, 35, n), rng.uniform(-0.7, 0.7, n)))
    x[0], u[0] = [1.0, 2.0, 0.2, 8.0], [1.0, 0.1]           # the vector quoted in SURVEY.md §8c
    dt_L = np.column_stack((rng.choice([0.05, 0.1, 0.2], n), rng.uniform(2.0, 15.0, n)))
    dt_L[0] = [0.1, 5.0]
    A = np.zeros((n, 4, 4)); B = np.zeros((n, 4, 2)); fx = np.zeros((n, 4)); f = np.zeros((n, 4))
    for i in range(n):
        A[i], B[i], fx[i] = linearize(x[i], u[i], dt_L[i, 0], dt_L[i, 1])
        f[i] = f_discrete(x[i], u[i], dt_L[i, 0], dt_L[i, 1])
    np.savez(os.path.join(HERE, "vehicle_model.npz"), x=x, u=u, dt_L=dt_L, A=A, B=B, fx=fx, f=f)


def golden_linearize_window():
    """(A_k,B_k,c_k) along a window exactly as mpc_controller.py:59-70,108-109 forms them, using the REAL linearize."""
    out = {}
    for name, N, seed in (("n20", 20, 2), ("n50", 50, 3)):
        x0, ref, up = make_batch(24, N, seed)
        # add windows that wrap through +-pi to exercise np.unwrap
        ref = ref.copy()
        ref[0, :, 2] = ((ref[0, :, 2] + np.pi) % (2 * np.pi)) - np.pi
        ref[1, :, 2] = ((ref[1, :, 2] + 3.0 + np.pi) % (2 * np.pi)) - np.pi
        As = np.zeros((len(ref), N, 4, 4)); Bs = np.zeros((len(ref), N, 4, 2)); cs = np.zeros((len(ref), N, 4))
        for b in range(len(ref)):
            r = np.copy(ref[b]); r[:, 2] = np.unwrap(r[:, 2])
            xlin = r[0]; ulin = np.zeros(2)
            for k in range(N):
                A, B, fx = linearize(xlin, ulin, 0.1, 2.8 / 0.8)           # wheelbase_px of config.py:79-81
                As[b, k], Bs[b, k], cs[b, k] = A, B, fx - A @ xlin - B @ ulin
                xlin = r[k]
        out.update({f"{name}_ref": ref, f"{name}_A": As, f"{name}_B": Bs, f"{name}_c": cs})
    np.savez(os.path.join(HERE, "linearize_window.npz"), **out)


def golden_ref_builder():
    rng = np.random.default_rng(5)
    out = {}
    for i in range(12):
        path = random_path(rng, rng.uniform(20, 300), ds=rng.choice([0.5, 1.0, 3.0]))
        if i == 0:
            path = path[:2]
        if i == 1:
            path = np.array([[0.0, 0.0], [1.0, 0.0]])         # shorter than one step
        v, N, dt = [(15.0, 15, 0.1), (15.0, 50, 0.1), (28.0, 12, 0.1), (40.0, 20, 0.05)][i % 4]
        out[f"path{i}"] = path
        out[f"args{i}"] = np.array([v, N, dt])
        out[f"ref{i}"] = build_reference([tuple(p) for p in path], v, int(N), dt)
    np.savez(os.path.join(HERE, "ref_builder.npz"), **out)


def golden_default_scenario():
    """config.py defaults: 80x80 map (seed 4), inflation 0.75 m at 0.8 m/px, start (70,70), goal (10,10), RRT* seed 13."""
    base = MapGenerator((80.0, 80.0), 1.0, seed=4).generate()
    raw = (base * 255).astype(np.uint8)                      # what save_grayscale / load_grayscale round-trips
    inflated = inflate_grayscale_map(raw, 0.75, 0.8)
    occupancy = np.flipud(to_occupancy_grid(inflated))
    start = (70, 70)
    goal = (occupancy.shape[1] - 70, occupancy.shape[0] - 70)
    params = PlannerParameters(step=3.0, goal_radius=10.0, max_iterations=2000, rewire_radius=20.0, goal_sample_rate=0.1,
                               random_seed=13, prune_path=True, spline_samples=20, spline_alpha=0.5, dedupe_tolerance=1e-9,
                               collision_step=0.75)
    plan = RRTStarPlanner(occupancy, params).plan(start, goal)
    assert plan.success
    path = np.array(plan.path, dtype=float)
    ref_global = build_reference(plan.path, 15.0, 15, 0.1)
    np.savez(os.path.join(HERE, "default_scenario.npz"), path=path, ref_global=ref_global, start=np.array(start, float),
             goal=np.array(goal, float), occupancy=occupancy.astype(np.uint8))
    print("default scenario: path points", len(path), "ref_global", ref_global.shape, "yaw[0:2]", ref_global[:2, 2])


def golden_optima():
    """Certified optima (KKT-Newton) for the reference's unit-test input and seeded synthetic problems."""
    out = {}
    p5 = O.Params(horizon=5, wheelbase_px=2.8 / 0.2)          # tests/test_mpc_controller.py:8-13
    x0 = np.array([0.0, 0.0, 0.0, 5.0]); ref = np.tile(np.array([1.0, 0.0, 0.0, 5.0]), (6, 1))
    u0, X, U, _ = O.solve_kkt_newton(x0, ref, None, p5)
    out.update(unit_x0=x0, unit_ref=ref, unit_u0=u0, unit_X=X, unit_U=U)
    for name, N, seed, du in (("n20", 20, 2, 0.15), ("n50", 50, 3, 0.02)):
        p = dataclasses.replace(O.Params(horizon=N), du_bounds=((-12.0, 12.0), (-du, du)))
        x0, ref, up = make_batch(32, N, seed)
        U0 = np.zeros((32, 2)); XS = np.zeros((32, 4, N + 1)); US = np.zeros((32, 2, N)); NS = np.zeros(32, int)
        for b in range(32):
            U0[b], XS[b], US[b], sl = O.solve_kkt_newton(x0[b], ref[b], up[b], p)
            NS[b] = int((sl > 1e-9).sum())
        out.update({f"{name}_x0": x0, f"{name}_ref": ref, f"{name}_up": up, f"{name}_u0": U0, f"{name}_X": XS, f"{name}_U": US,
                    f"{name}_active_slacks": NS})
    np.savez(os.path.join(HERE, "optima.npz"), **out)


if __name__ == "__main__":
    golden_vehicle_model()
    golden_linearize_window()
    golden_ref_builder()
    golden_default_scenario()
    golden_optima()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
