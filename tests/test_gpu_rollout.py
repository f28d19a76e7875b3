"""GPU (B200): closed loop (TrajectoryTracker.track / track_batch) against the oracle's restatement of
control_stage.py:74-157 on the default-config scenario reproduced from the real reference modules."""
import dataclasses
from types import SimpleNamespace as NS

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
TIGHT = dict(eps_abs=1e-6, eps_rel=1e-6)


def scenario():
    d = load_golden("default_scenario.npz")
    path = [tuple(p) for p in d["path"]]
    return d, path


def test_default_rollout_matches_oracle_within_1e3():
    from oracle import c_oracle as CO, mpc_numpy as O
    from rrt_mpc_b200 import MPCConfig, SolverSettings, TrajectoryTracker
    from rrt_mpc_b200.control_stage import initial_state
    d, path = scenario()
    ora = CO.track(O.Params(horizon=15), d["ref_global"], initial_state(path, d["start"]), d["goal"], 300, polish_passes=3, **TIGHT)
    tr = TrajectoryTracker(MPCConfig(), None, settings=SolverSettings(polish_passes=3, polish_retry=2, **TIGHT))
    # (1) batched, device-resident loop; cold start each step like the reference (a new Problem per call)
    res = tr.track_batch([path], [d["start"]], [d["goal"]], map_resolution=0.8, warm_start=False)
    n = int(res.n_steps[0])
    assert n == ora["n_steps"] and bool(res.goal_reached[0]) and not bool(res.aborted[0])
    assert np.abs(res.states[0, :n, :2] - ora["states"][:n, :2]).max() < 1e-3          # tracked positions, full rollout
    assert np.abs(res.controls[0, :n] - ora["controls"][:n]).max() < 1e-5             # every first control
    assert np.isnan(res.states[0, n:]).all()
    # (2) warm-started loop converges to the same optima
    resw = tr.track_batch([path], [d["start"]], [d["goal"]], map_resolution=0.8, warm_start=True)
    assert int(resw.n_steps[0]) == n
    assert np.abs(resw.states[0, :n, :2] - ora["states"][:n, :2]).max() < 1e-3
    assert (resw.step_status[0, :n] == 1).all()
    # (3) the reference signature: track(planning, maps, ...) -> TrackingResult(states=[...])
    planning = NS(plan=NS(success=True, path=path))
    maps = NS(start=tuple(d["start"]), goal=tuple(d["goal"]))
    out = tr.track(planning, maps, map_resolution=0.8, visualize=False)
    assert len(out.states) == n
    assert np.abs(np.array(out.states)[:, :2] - ora["states"][:n, :2]).max() < 1e-3


def test_many_perturbed_vehicles():
    """config 4 shape in small: perturbed copies of the default path, each tracked independently on the device."""
    from oracle import c_oracle as CO, mpc_numpy as O
    from rrt_mpc_b200 import MPCConfig, SolverSettings, TrajectoryTracker
    from rrt_mpc_b200.control_stage import initial_state
    from rrt_mpc_b200.ref_builder import build_reference
    d, path = scenario()
    rng = np.random.default_rng(4)
    B, T = 48, 80
    paths, starts = [], []
    for b in range(B):
        pts = np.array(path) + rng.normal(size=(len(path), 2)) * 0.15
        pts[0] = path[0]
        paths.append([tuple(p) for p in pts]); starts.append(pts[0] + rng.normal(size=2) * 0.5)
    goals = np.tile(d["goal"], (B, 1))
    tr = TrajectoryTracker(MPCConfig(sim_steps=T), None, settings=SolverSettings(polish_passes=3, polish_retry=2, **TIGHT))
    res = tr.track_batch(paths, starts, goals, map_resolution=0.8, warm_start=True)                       # references built on the device (K_ref)
    res_h = tr.track_batch(paths, starts, goals, map_resolution=0.8, warm_start=True, build_on_device=False)  # ... and by the NumPy mirror
    # 1e-12 differences in the reference rows can change which solves end polished; closed-loop bar is 1e-3 px
    assert np.array_equal(res.n_steps, res_h.n_steps) and np.nanmax(np.abs(res.states[:, :, :2] - res_h.states[:, :, :2])) < 1e-3
    assert not res.aborted.any() and res.goal_reached.mean() > 0.9
    for b in (0, 7, 23, 47):
        rg = build_reference(paths[b], 15.0, 15, 0.1)
        ora = CO.track(O.Params(horizon=15), rg, initial_state(paths[b], starts[b]), goals[b], T, polish_passes=3, **TIGHT)
        n = int(res.n_steps[b])
        assert n == ora["n_steps"]
        assert np.abs(res.states[b, :n, :2] - ora["states"][:n, :2]).max() < 1e-3


def test_reference_closed_loop_states_within_1e3():
    """The 65 states the reference's own TrajectoryTracker.track returned for the default scenario (tests/golden/ref_qp.npz,
    executed from /root/reference/src/pipeline/control_stage.py:58-157 by tests/golden/make_ref_qp.py)."""
    from rrt_mpc_b200 import MPCConfig, SolverSettings, TrajectoryTracker
    g = load_golden("ref_qp.npz")
    d, path = scenario()
    want = g["roll_states"]
    tr = TrajectoryTracker(MPCConfig(), None, settings=SolverSettings(polish_passes=3, polish_retry=2, **TIGHT))
    for warm in (False, True):
        res = tr.track_batch([path], [d["start"]], [d["goal"]], map_resolution=0.8, warm_start=warm)
        n = int(res.n_steps[0])
        assert n == len(want) and bool(res.goal_reached[0]) and not bool(res.relaxed[0])
        assert np.abs(res.states[0, :n, :2] - want[:, :2]).max() < 1e-3                # north star: positions within 1e-3
        assert np.abs(res.states[0, :n] - want).max() < 1e-5                           # what we reach (all four states)
        assert np.abs(res.controls[0, :n] - g["roll_u0"]).max() < 1e-5


def test_relaxation_branch_is_entered_and_matches_the_oracle():
    """control_stage.py:45-56.  Every limit is soft, so the retry is reachable only through the iteration limit: with
    max_iter = 550 at eps 1e-6 the nominal solves of steps 0, 1, 4 and 5 end `max_iter reached` while the relaxed problems
    (v_ref * 0.6, du_bounds widened) converge; the roll-out then completes.  With max_iter = 300 the retry fails too and the
    vehicle aborts at step 0.  Oracle in mirror mode (unscaled, z0 = clip(0)): same statuses, iterations and states."""
    from oracle import c_oracle as CO, mpc_numpy as O
    from rrt_mpc_b200 import MPCConfig, SolverSettings, TrajectoryTracker
    from rrt_mpc_b200.control_stage import initial_state
    d, path = scenario()
    s0 = initial_state(path, d["start"])
    for max_iter, n_want, aborted in ((550, 63, False), (300, 0, True)):
        ora = CO.track(O.Params(horizon=15), d["ref_global"], s0, d["goal"], 300, polish_passes=1, scaling=0, z0_projected=1,
                       max_iter=max_iter, **TIGHT)
        assert ora["n_steps"] == n_want and bool(ora["flags"] & 4) and bool(ora["flags"] & 2) == aborted
        tr = TrajectoryTracker(MPCConfig(), None, settings=SolverSettings(polish_passes=1, max_iter=max_iter, **TIGHT))
        res = tr.track_batch([path, path], [d["start"]] * 2, [d["goal"]] * 2, map_resolution=0.8, warm_start=False)
        for b in range(2):
            n = int(res.n_steps[b])
            assert n == n_want and bool(res.relaxed[b]) and bool(res.aborted[b]) == aborted and bool(res.goal_reached[b]) != aborted
            k = n + (1 if aborted else 0)                                                # an aborted vehicle still reports its failed step
            assert np.array_equal(res.step_status[b, :k], ora["step_status"][:k])
            assert np.array_equal(res.step_iters[b, :k], ora["step_iters"][:k])
            if n:
                assert np.abs(res.states[b, :n] - ora["states"][:n]).max() < 1e-7
                assert (res.step_iters[b, [0, 1, 4, 5]] <= max_iter).all() and set(np.unique(res.step_status[b, :n])) <= {1, 2}
        # without the retry the same roll-out aborts at step 0
        res0 = tr.track_batch([path], [d["start"]], [d["goal"]], map_resolution=0.8, warm_start=False, relax_on_failure=False)
        assert bool(res0.aborted[0]) and int(res0.n_steps[0]) == 0 and res0.step_status[0, 0] == -2
    # the reference signature on the host: _solve_with_relaxation returns the relaxed solve's triple
    tr = TrajectoryTracker(MPCConfig(), None, settings=SolverSettings(polish_passes=1, max_iter=550, **TIGHT))
    base = MPCConfig().to_parameters(0.8)
    window = d["ref_global"][:16]
    u0, Xp, Up = tr._solve_with_relaxation(s0, window, np.zeros(2), base)
    relaxed = CO.solve_batch(dataclasses.replace(O.Params(horizon=15), du_bounds=((-17.0, 17.0), (-0.2, 0.2))), s0[None],
                             (window * np.array([1, 1, 1, 0.6]))[None], np.zeros((1, 2)), polish_passes=1, scaling=0, z0_projected=1, max_iter=550, **TIGHT)
    assert u0 is not None and np.abs(u0 - relaxed["u0"][0]).max() < 1e-8
    assert tr._controller(base).solve(s0, window, u_prev=np.zeros(2)) == (None, None, None)
