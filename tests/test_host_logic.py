"""CPU: host-side mirror of the reference interface (config, parameters, synthetic inputs, tracker guards)."""
import dataclasses

import numpy as np
import pytest

from conftest import load_golden


def test_mpc_config_defaults_match_reference():
    """src/config.py:66-92"""
    from rrt_mpc_b200 import MPCConfig
    c = MPCConfig()
    assert (c.wheelbase_m, c.dt, c.horizon, c.v_px_s, c.sim_steps) == (2.8, 0.1, 15, 15.0, 300)
    p = c.to_parameters(0.8)
    assert p.wheelbase_px == 2.8 / 0.8 and p.horizon == 15
    assert np.array_equal(np.diag(p.q), [4.0, 4.0, 0.6, 0.1]) and np.array_equal(np.diag(p.r), [0.03, 0.25])
    assert np.array_equal(np.diag(p.q_terminal), [8.0, 8.0, 1.0, 0.2])
    assert p.u_bounds == ((-35.0, 35.0), (-0.6, 0.6)) and p.v_bounds == (0.0, 90.0) and p.du_bounds == ((-12.0, 12.0), (-0.15, 0.15))
    assert (p.slack_velocity, p.slack_input, p.slack_rate) == (1e3, 5e2, 5e2)
    assert MPCConfig(horizon=5).to_parameters(0.2).wheelbase_px == pytest.approx(14.0)   # tests/test_mpc_controller.py:8-9


def test_params_to_c_roundtrip():
    from rrt_mpc_b200 import MPCConfig
    from rrt_mpc_b200.mpc_controller import params_to_c, SolverSettings
    p = MPCConfig(horizon=20).to_parameters(0.8)
    c = params_to_c(p)
    assert c.horizon == 20 and list(c.u_bounds) == [-35.0, 35.0, -0.6, 0.6] and list(c.du_bounds) == [-12.0, 12.0, -0.15, 0.15]
    assert c.q[0] == 4.0 and c.q[5] == 4.0 and c.q[10] == 0.6 and c.q[15] == 0.1 and c.q[1] == 0.0
    s = SolverSettings(eps_abs=1e-6, eps_rel=1e-6, polish=False).to_c()
    assert s.polish_passes == 0 and s.eps_abs == 1e-6 and s.max_iter == 60000
    d = SolverSettings().to_c()                       # the reference's call: OSQP defaults for everything it does not pass
    assert (d.check_termination, d.adaptive_rho_interval, d.early_polish, d.polish_retry, d.warm_start) == (25, 50, 0, 0, 0)
    e = SolverSettings.early_certified().to_c()       # what bench.py's headline runs
    assert (e.eps_abs, e.polish_passes, e.polish_retry, e.early_polish, e.check_termination, e.warm_start) == (1e-6, 5, 4, 1, 50, -1)
    assert SolverSettings.early_certified(1e-4, polish_retry=2).to_c().polish_retry == 2


def test_synthetic_batches_are_deterministic_and_shardable():
    from rrt_mpc_b200.synthetic import make_batch
    a = make_batch(10000, 20, seed=2)
    b = make_batch(10000, 20, seed=2)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    s0 = make_batch(10000, 20, seed=2, start=0, count=5000)
    s1 = make_batch(10000, 20, seed=2, start=5000, count=5000)
    for full, lo, hi in zip(a, s0, s1):
        assert np.array_equal(full[:5000], lo) and np.array_equal(full[5000:], hi)
    x0, ref, up = make_batch(64, 50, seed=3)
    assert x0.shape == (64, 4) and ref.shape == (64, 51, 4) and up.shape == (64, 2)
    assert (np.abs(np.diff(ref[:, :, :2], axis=1)).max(axis=(1, 2)) > 0).all()          # windows are not all tail padding
    assert np.abs(up[:, 0]).max() <= 5.0 and np.abs(up[:, 1]).max() <= 0.2


def test_tracker_guards_like_reference():
    """control_stage.py:69-72"""
    from types import SimpleNamespace as NS
    from rrt_mpc_b200 import MPCConfig, TrajectoryTracker
    tr = TrajectoryTracker(MPCConfig(), None)
    maps = NS(start=(0, 0), goal=(1, 1))
    with pytest.raises(RuntimeError, match="did not succeed"):
        tr.track(NS(plan=NS(success=False, path=[])), maps, map_resolution=0.8, visualize=False)
    with pytest.raises(RuntimeError, match="empty path"):
        tr.track(NS(plan=NS(success=True, path=[])), maps, map_resolution=0.8, visualize=False)


def test_initial_state_matches_reference_rule():
    from rrt_mpc_b200.control_stage import initial_state
    d = load_golden("default_scenario.npz")
    s = initial_state([tuple(p) for p in d["path"]], d["start"])
    assert s[0] == 70.0 and s[1] == 70.0 and s[3] == 5.0
    assert s[2] == pytest.approx(np.arctan2(d["path"][1, 1] - d["path"][0, 1], d["path"][1, 0] - d["path"][0, 0]))
    assert initial_state([(1.0, 2.0)], (1.0, 2.0))[2] == 0.0


def test_batch_tracking_result_roundtrip(tmp_path):
    from rrt_mpc_b200 import BatchTrackingResult
    T = 5
    states = np.full((2, T, 4), np.nan); states[0, :3] = np.arange(12).reshape(3, 4); states[1, :5] = 1.0
    r = BatchTrackingResult(states=states, controls=np.zeros((2, T, 2)), n_steps=np.array([3, 5], np.int32), goal_reached=np.array([True, False]),
                            aborted=np.array([False, False]), step_status=np.ones((2, T), np.int32), step_iters=np.full((2, T), 25, np.int32))
    f = str(tmp_path / "roll.npz")
    r.save_npz(f)
    q = BatchTrackingResult.load_npz(f)
    assert np.array_equal(q.n_steps, r.n_steps) and np.array_equal(np.isnan(q.states), np.isnan(r.states))
    tr = q.result(0)                                       # what TrackingResult.states holds for vehicle 0 (artifacts.py:34-38)
    assert len(tr.states) == 3 and np.array_equal(tr.states[2], [8.0, 9.0, 10.0, 11.0])


def test_solve_with_relaxation_retries_once_with_relaxed_problem():
    """control_stage.py:33-56 on the host, with the controllers replaced by recorders (no GPU needed): a nominal failure
    triggers exactly one retry with v_ref * 0.6 and du_bounds widened by (5.0, 0.05); the retry's triple is returned as is."""
    from rrt_mpc_b200 import MPCConfig, TrajectoryTracker
    calls = []

    class Fake:
        def __init__(self, params, outcome):
            self.params, self.outcome = params, outcome

        def solve(self, state, reference, *, u_prev=None):
            calls.append((self.params, np.array(reference, copy=True), np.array(u_prev, copy=True)))
            return self.outcome

    tr = TrajectoryTracker(MPCConfig(), None)
    base = MPCConfig().to_parameters(0.8)
    ok = (np.array([1.0, 0.1]), np.zeros((4, 16)), np.zeros((2, 15)))
    ref = np.tile(np.array([1.0, 2.0, 0.3, 10.0]), (16, 1))
    # nominal success: one call, no retry
    tr._controller = lambda params, key="base": Fake(params, ok)
    out = tr._solve_with_relaxation(np.zeros(4), ref, np.array([0.5, 0.0]), base)
    assert all(a is b for a, b in zip(out, ok)) and len(calls) == 1 and calls[0][0] is base
    # nominal failure: retry with the relaxed problem, and its result (here again a failure) is returned unchanged
    calls.clear()
    outcomes = iter([(None, None, None), (None, None, None)])
    tr._controller = lambda params, key="base": Fake(params, next(outcomes))
    out = tr._solve_with_relaxation(np.zeros(4), ref, np.array([0.5, 0.0]), base)
    assert out == (None, None, None) and len(calls) == 2
    relaxed_params, relaxed_ref, up = calls[1]
    assert relaxed_params.du_bounds == ((-17.0, 17.0), (-0.2, 0.2))                 # +-12 -> +-17, +-0.15 -> +-0.2
    assert np.array_equal(relaxed_ref[:, 3], ref[:, 3] * 0.6) and np.array_equal(relaxed_ref[:, :3], ref[:, :3])
    assert np.array_equal(ref[:, 3], np.full(16, 10.0))                             # the caller's window is not mutated
    assert np.array_equal(up, [0.5, 0.0]) and relaxed_params.u_bounds == base.u_bounds
