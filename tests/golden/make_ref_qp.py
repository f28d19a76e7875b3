"""Generate tests/golden/ref_qp.npz by EXECUTING the reference's own solver wrapper and closed loop.

Run in the build container (where /root/reference exists):   python tests/golden/make_ref_qp.py

What runs here is the UNMODIFIED reference source:
  * ``src.control.mpc_controller.MPCController.solve``  (/root/reference/src/control/mpc_controller.py:39-141)
  * ``src.pipeline.control_stage.TrajectoryTracker.track`` (/root/reference/src/pipeline/control_stage.py:58-157)
with two absent third-party packages replaced in ``sys.modules``:
  * ``cvxpy``  -> tests/golden/ref_qp_stub.py, which records the QP the reference states (objective terms and
                  constraints exactly as written) and solves it with a dense interior-point + active-set solver that
                  shares no code with oracle/ or the CUDA path;
  * ``matplotlib`` -> inert mocks (track() is called with visualize=False; only the imports must succeed).
OSQP itself (the ADMM iteration, its status codes, its iteration counts) is NOT executed: with a strictly convex QP the
optimum is unique, so the recorded QP plus its exact optimum is what pins parity.

Fixture layout (all arrays in one compressed npz):
  <set>_x0, <set>_ref, <set>_up            inputs handed to MPCController.solve
  <set>_u0, <set>_X, <set>_U              what the reference returned (U.value[:,0], X.value, U.value)
  <set>_z                                  the stub solver's full primal vector (stub variable layout)
  <set>_kkt                                (cases, 4) KKT residuals of that solution: dual, eq, ineq, -min multiplier
  <set>_qp_idx                             which cases carry the recorded QP
  <set>_qp<i>_{H,Aeq,Ain}_{r,c,v}, _g, _beq, _bin, _c0      recorded QP of case i (COO triplets, stub layout)
  sets: unit (the reference's own unit-test input), roll (every solve of the default-config closed loop),
        n20 / n50 (seeded synthetic configs 2 / 3), nd20 (non-diagonal Q, R, Q_N)
  roll_states, roll_path_idx               the closed-loop states TrajectoryTracker.track returned
"""
import dataclasses
import os
import sys
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_qp_stub as stub                                            # noqa: E402

sys.modules["cvxpy"] = stub
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.axes", "matplotlib.figure", "matplotlib.patches", "matplotlib.lines",
             "matplotlib.colors", "matplotlib.rcsetup", "matplotlib.animation", "matplotlib.collections", "matplotlib.transforms",
             "imageio", "imageio.v2"):
    sys.modules[name] = mock.MagicMock()
for pkg in ("src", "src.pipeline"):                                   # skip the eager __init__ import chains
    shim = types.ModuleType(pkg)
    shim.__path__ = [os.path.join(REF, *pkg.split("."))]
    sys.modules[pkg] = shim

from src.config import MPCConfig, VizConfig                           # noqa: E402
from src.control.mpc_controller import MPCController                  # noqa: E402
from src.pipeline.artifacts import MapArtifacts, PlanningArtifacts    # noqa: E402
from src.pipeline.control_stage import TrajectoryTracker              # noqa: E402
from src.planning.plan_result import PlanResult                       # noqa: E402

from rrt_mpc_b200.synthetic import make_batch                         # noqa: E402

CALLS = []
_orig_solve = MPCController.solve


def _recording_solve(self, x0, ref_traj, *, u_init=None, u_prev=None):
    """Pass-through wrapper: notes the arguments of each call, then runs the reference's own solve()."""
    n0 = len(stub.RECORDS)
    out = _orig_solve(self, x0, ref_traj, u_init=u_init, u_prev=u_prev)
    assert len(stub.RECORDS) == n0 + 1
    CALLS.append(dict(x0=np.array(x0, float), ref=np.array(ref_traj, float), up=np.zeros(2) if u_prev is None else np.array(u_prev, float),
                      out=out, rec=stub.RECORDS[-1], params=self.params))
    return out


MPCController.solve = _recording_solve


def coo(M):
    r, c = np.nonzero(M)
    return r.astype(np.int32), c.astype(np.int32), M[r, c]


def pack(name, calls, out, qp_every=1):
    out[f"{name}_x0"] = np.stack([c["x0"] for c in calls])
    out[f"{name}_ref"] = np.stack([c["ref"] for c in calls])
    out[f"{name}_up"] = np.stack([c["up"] for c in calls])
    out[f"{name}_u0"] = np.stack([c["out"][0] for c in calls])
    out[f"{name}_X"] = np.stack([c["out"][1] for c in calls])
    out[f"{name}_U"] = np.stack([c["out"][2] for c in calls])
    out[f"{name}_z"] = np.stack([c["rec"]["z"] for c in calls])
    out[f"{name}_kkt"] = np.array([[c["rec"]["info"]["dual"], c["rec"]["info"]["eq"], c["rec"]["info"]["ineq"], -c["rec"]["info"]["lam_min"]] for c in calls])
    for c in calls:
        assert c["rec"]["info"]["method"] == "active-set vertex", c["rec"]["info"]
        assert c["rec"]["solver"] == "OSQP" and c["rec"]["kwargs"]["polish"] is True and c["rec"]["kwargs"]["eps_abs"] == 1e-3
    idx = list(range(0, len(calls), qp_every))
    out[f"{name}_qp_idx"] = np.array(idx, np.int32)
    for i in idx:
        qp = calls[i]["rec"]["qp"]
        for key in ("H", "Aeq", "Ain"):
            r, c, v = coo(qp[key])
            out[f"{name}_qp{i}_{key}_r"], out[f"{name}_qp{i}_{key}_c"], out[f"{name}_qp{i}_{key}_v"] = r, c, v
        out[f"{name}_qp{i}_g"], out[f"{name}_qp{i}_beq"], out[f"{name}_qp{i}_bin"] = qp["g"], qp["beq"], qp["bin"]
        out[f"{name}_qp{i}_c0"] = np.array(qp["c0"])
    print(f"{name}: {len(calls)} solves, n = {calls[0]['rec']['qp']['H'].shape[0]}, rows = {calls[0]['rec']['qp']['Aeq'].shape[0]} eq + "
          f"{calls[0]['rec']['qp']['Ain'].shape[0]} ineq, worst KKT residual {out[f'{name}_kkt'].max():.2e}")


def nondiag_params(N):
    """PSD, non-diagonal Q / R / Q_N (cp.quad_form accepts any PSD matrix, mpc_controller.py:74-75,112)."""
    base = MPCConfig(horizon=N).to_parameters(0.8)
    rng = np.random.default_rng(77)

    def perturb(M, scale):
        G = rng.normal(size=M.shape) * scale
        S = M + 0.5 * (G + G.T) * np.sqrt(np.outer(np.diag(M), np.diag(M)))
        w = np.linalg.eigvalsh(S)
        assert w.min() > 0.2 * np.diag(M).min(), w
        return S
    return dataclasses.replace(base, q=perturb(base.q, 0.25), r=perturb(base.r, 0.25), q_terminal=perturb(base.q_terminal, 0.25))


def main():
    out = {}
    # (i) the reference's own unit-test input (tests/test_mpc_controller.py:7-17)
    CALLS.clear()
    p5 = MPCConfig(horizon=5).to_parameters(0.2)
    u0, Xp, Up = MPCController(p5).solve(np.array([0.0, 0.0, 0.0, 5.0]), np.tile(np.array([1.0, 0.0, 0.0, 5.0]), (6, 1)))
    assert u0 is not None and Xp[0, 1] > 0.0                                  # the reference test's own assertions
    pack("unit", list(CALLS), out)

    # (ii) the default-config closed loop through the reference's TrajectoryTracker.track
    sc = np.load(os.path.join(HERE, "default_scenario.npz"))
    plan = PlanResult(success=True, path=[tuple(p) for p in sc["path"]], nodes=[], iterations=0, goal_index=None)
    maps = MapArtifacts(occupancy=sc["occupancy"], raw_occupancy=sc["occupancy"], inflation_mask=sc["occupancy"],
                        start=tuple(sc["start"]), goal=tuple(sc["goal"]))
    CALLS.clear()
    tr = TrajectoryTracker(MPCConfig(), VizConfig()).track(PlanningArtifacts(plan=plan), maps, map_resolution=0.8, visualize=False)
    pack("roll", list(CALLS), out, qp_every=4)
    out["roll_states"] = np.stack(tr.states)
    print("closed loop: steps", len(tr.states), "final", tr.states[-1])

    # (iii) seeded synthetic problems of BASELINE configs 2 and 3
    for name, N, seed, du, cnt in (("n20", 20, 2, 0.15, 32), ("n50", 50, 3, 0.02, 8)):
        p = dataclasses.replace(MPCConfig(horizon=N).to_parameters(0.8), du_bounds=((-12.0, 12.0), (-du, du)))
        x0, ref, up = make_batch(32, N, seed)
        CALLS.clear()
        for b in range(cnt):
            MPCController(p).solve(x0[b], ref[b], u_prev=up[b])
        pack(name, list(CALLS), out, qp_every=4)

    # (iv) non-diagonal weights
    N = 20
    pnd = nondiag_params(N)
    x0, ref, up = make_batch(32, N, 2)
    CALLS.clear()
    for b in range(16):
        MPCController(pnd).solve(x0[b], ref[b], u_prev=up[b])
    pack("nd20", list(CALLS), out, qp_every=8)
    out["nd20_q"], out["nd20_r"], out["nd20_qn"] = pnd.q, pnd.r, pnd.q_terminal

    path = os.path.join(HERE, "ref_qp.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
