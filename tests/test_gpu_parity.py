"""GPU (B200): the CUDA path, called through the C ABI (ctypes wrapper), against the oracle and the golden vectors."""
import dataclasses

import numpy as np
import pytest

from conftest import load_golden, oracle_params, product_params

pytestmark = pytest.mark.gpu
TIGHT = dict(eps_abs=1e-6, eps_rel=1e-6)


def controller(N, du=0.15, passes=3, max_batch=64, retry=2, **kw):
    from rrt_mpc_b200 import MPCController, SolverSettings
    return MPCController(product_params(N, du), SolverSettings(polish_passes=passes, polish_retry=retry, **{**TIGHT, **kw}), max_batch=max_batch)


def test_extension_is_loaded_and_counts_launches():
    import os
    ctl = controller(20)
    g = load_golden("optima.npz")
    ctl.solve_batch(g["n20_x0"][:4], g["n20_ref"][:4], u_prev=g["n20_up"][:4])
    assert ctl.launch_count() == 1
    maps = open(f"/proc/{os.getpid()}/maps").read()
    assert "libcudampc.so" in maps


@pytest.mark.parametrize("name,N", [("n20", 20), ("n50", 50)])
def test_linearize_batch_within_1e12_of_reference(name, N):
    g = load_golden("linearize_window.npz")          # produced by the REAL vehicle_model.linearize
    ctl = controller(N)
    A, B, c = ctl.linearize_batch(g[f"{name}_ref"])
    sa = np.abs(g[f"{name}_A"]).max()
    assert np.abs(A - g[f"{name}_A"]).max() <= 1e-12 * sa
    assert np.abs(B - g[f"{name}_B"]).max() <= 1e-12 * np.abs(g[f"{name}_B"]).max()
    assert np.abs(c - g[f"{name}_c"]).max() <= 1e-12 * max(sa, np.abs(g[f"{name}_ref"][:, :, :2]).max())
    assert np.array_equal(A[:, :, 2, 3], np.zeros_like(A[:, :, 2, 3]))       # ulin = 0 => A[2,3] = 0


@pytest.mark.parametrize("name,N,du", [("n20", 20, 0.15), ("n50", 50, 0.02)])
def test_u0_within_1e5_of_certified_optimum(name, N, du):
    g = load_golden("optima.npz")
    ctl = controller(N, du)
    r = ctl.solve_batch(g[f"{name}_x0"], g[f"{name}_ref"], u_prev=g[f"{name}_up"])
    assert (r.status == 1).all()
    assert np.abs(r.u0 - g[f"{name}_u0"]).max() < 1e-5                     # the north-star bar
    assert np.abs(r.u0 - g[f"{name}_u0"]).max() < 1e-8                     # what we actually reach
    assert np.abs(r.Xp - g[f"{name}_X"]).max() < 1e-6 and np.abs(r.Up - g[f"{name}_U"]).max() < 1e-6
    assert np.abs(r.Xp[:, :, 0] - g[f"{name}_x0"]).max() < 1e-9            # X_0 = x0


@pytest.mark.parametrize("name,N,du", [("n20", 20, 0.15), ("n50", 50, 0.02)])
def test_status_and_iterations_identical_to_oracle(name, N, du):
    from oracle import c_oracle as CO
    g = load_golden("optima.npz")
    nb = 12
    ctl = controller(N, du, passes=1, retry=0)
    r = ctl.solve_batch(g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], u_prev=g[f"{name}_up"][:nb])
    # mirror mode: the same ADMM on the same QP, solved through a generic sparse KKT LDL' on the CPU
    c = CO.solve_batch(oracle_params(N, du), g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], scaling=0, z0_projected=1, **TIGHT)
    assert np.array_equal(r.status, c["status"]) and np.array_equal(r.iters, c["iters"])
    assert np.array_equal(r.info[:, 0], c["info"][:, 0]) and np.array_equal(r.info[:, 2], c["info"][:, 2])
    assert np.abs(r.u0 - c["u0"]).max() < 1e-9
    # literal OSQP defaults (Ruiz scaling 10): same status, same optimum within the bar where the oracle itself is converged
    o = CO.solve_batch(oracle_params(N, du), g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], polish_passes=3, **TIGHT)
    assert np.array_equal(r.status, o["status"])
    r3 = controller(N, du, passes=3).solve_batch(g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], u_prev=g[f"{name}_up"][:nb])
    assert np.abs(r3.u0 - o["u0"]).max() < 1e-5


@pytest.mark.parametrize("check", [25, 50])      # 50 = the rho-adaptation interval: the setting bench.py runs early polish with
@pytest.mark.parametrize("name,N,du", [("n20", 20, 0.15), ("n50", 50, 0.02)])
def test_early_polish_same_optimum_fewer_iterations(name, N, du, check):
    g = load_golden("optima.npz")
    lit = controller(N, du, passes=5).solve_batch(g[f"{name}_x0"], g[f"{name}_ref"], u_prev=g[f"{name}_up"])
    ear = controller(N, du, passes=5, early_polish=True, check_termination=check).solve_batch(g[f"{name}_x0"], g[f"{name}_ref"], u_prev=g[f"{name}_up"])
    assert (ear.status == 1).all() and (ear.info[:, 2] > 0).all()
    assert np.abs(ear.u0 - g[f"{name}_u0"]).max() < 1e-8 and np.abs(ear.Xp - g[f"{name}_X"]).max() < 1e-6
    assert np.abs(ear.u0 - lit.u0).max() < 1e-9
    assert ear.iters.sum() < 0.75 * lit.iters.sum()


def test_matches_host_emulation_of_the_same_source():
    import emu_driver as E
    g = load_golden("optima.npz")
    ctl = controller(20)
    r = ctl.solve_batch(g["n20_x0"][:8], g["n20_ref"][:8], u_prev=g["n20_up"][:8])
    e = E.solve(oracle_params(20), g["n20_x0"][:8], g["n20_ref"][:8], g["n20_up"][:8], polish_passes=3, polish_retry=2, **TIGHT)
    assert np.array_equal(r.iters, e["iters"]) and np.array_equal(r.status, e["status"])
    assert np.abs(r.u0 - e["u0"]).max() < 1e-10 and np.abs(r.Xp - e["Xp"]).max() < 1e-9


def test_reference_signature_single_problem():
    """tests/test_mpc_controller.py:7-17 of the reference, verbatim semantics."""
    from rrt_mpc_b200 import MPCConfig, MPCController
    cfg = MPCConfig(horizon=5)
    controller_ = MPCController(cfg.to_parameters(map_resolution=0.2))
    x0 = np.array([0.0, 0.0, 0.0, 5.0])
    ref = np.tile(np.array([1.0, 0.0, 0.0, 5.0]), (cfg.horizon + 1, 1))
    ref_copy = ref.copy()
    u0, Xp, Up = controller_.solve(x0, ref)
    assert u0 is not None and Xp is not None and Up is not None
    assert Xp[0, 1] > x0[0]
    assert u0.shape == (2,) and Xp.shape == (4, 6) and Up.shape == (2, 5)
    assert np.array_equal(ref, ref_copy)                                   # inputs are not mutated
    g = load_golden("optima.npz")
    assert np.abs(u0 - g["unit_u0"]).max() < 1e-5
    u0b, _, _ = controller_.solve(x0, ref, u_init=np.ones((5, 2)), u_prev=np.zeros(2))     # u_init is dead upstream
    assert np.array_equal(u0, u0b)


def test_failed_solve_maps_to_none_triple():
    """mpc_controller.py:137-139: non-optimal status -> (None, None, None)."""
    from rrt_mpc_b200 import MPCController, SolverSettings
    g = load_golden("optima.npz")
    ctl = MPCController(product_params(20), SolverSettings(max_iter=5, **TIGHT))
    assert ctl.solve(g["n20_x0"][0], g["n20_ref"][0], u_prev=g["n20_up"][0]) == (None, None, None)
    r = ctl.solve_batch(g["n20_x0"][:3], g["n20_ref"][:3], u_prev=g["n20_up"][:3])
    assert (r.status == -2).all() and (r.iters == 5).all()


@pytest.mark.parametrize("B", [0, 1, 3, 31, 257])
def test_ragged_batch_sizes_and_device_path(B):
    import torch
    from rrt_mpc_b200.synthetic import make_batch
    x0, ref, up = make_batch(max(B, 1), 20, seed=7)
    x0, ref, up = x0[:B], ref[:B], up[:B]
    ctl = controller(20, max_batch=300)
    r = ctl.solve_batch(x0, ref, u_prev=up)
    assert r.u0.shape == (B, 2) and r.Xp.shape == (B, 4, 21) and r.Up.shape == (B, 2, 20)
    if B == 0:
        return
    d = lambda a: torch.as_tensor(a).cuda()
    rd = ctl.solve_batch(d(x0), d(ref), u_prev=d(up))
    torch.cuda.synchronize()
    assert np.array_equal(rd.status.cpu().numpy(), r.status) and np.array_equal(rd.iters.cpu().numpy(), r.iters)
    assert np.array_equal(rd.u0.cpu().numpy(), r.u0) and np.array_equal(rd.Xp.cpu().numpy(), r.Xp)
    assert (r.status == 1).all()
    r0 = ctl.solve_batch(x0, ref)                                          # u_prev = None -> zeros
    rz = ctl.solve_batch(x0, ref, u_prev=np.zeros((B, 2)))
    assert np.array_equal(r0.u0, rz.u0)


# 100 / 150: only two problems / one problem fit an SM -> the two-warps-per-problem instantiation of K_solve
@pytest.mark.parametrize("N", [1, 2, 7, 33, 64, 100, 150])
def test_horizon_edge_cases(N):
    from oracle import mpc_numpy as O
    p = oracle_params(N)
    rng = np.random.default_rng(N)
    ref = np.zeros((4, N + 1, 4))
    for b in range(4):
        yaw = 0.4 * b + np.cumsum(rng.normal(size=N + 1) * 0.05)
        ref[b, :, 2] = yaw; ref[b, :, 3] = 12.0 + rng.normal(size=N + 1)
        ref[b, :, 0] = 80 + np.cumsum(1.5 * np.cos(yaw)); ref[b, :, 1] = 60 + np.cumsum(1.5 * np.sin(yaw))
    x0 = ref[:, 0] + rng.normal(size=(4, 4)) * [0.5, 0.5, 0.05, 1.0]
    up = rng.uniform(-1, 1, size=(4, 2)) * [3.0, 0.1]
    r = controller(N).solve_batch(x0, ref, u_prev=up)
    assert (r.status == 1).all()
    for b in range(4):
        u0, X, U, _ = O.solve_kkt_newton(x0[b], ref[b], up[b], p)
        assert np.abs(r.u0[b] - u0).max() < 1e-6 and np.abs(r.Xp[b] - X).max() < 1e-6


def test_warm_start_same_optimum_fewer_iterations():
    from rrt_mpc_b200 import SolverSettings
    g = load_golden("optima.npz")
    ctl = controller(20)
    x0, ref, up = g["n20_x0"], g["n20_ref"], g["n20_up"]
    cold = ctl.solve_batch(x0, ref, u_prev=up)
    x1 = x0 + 0.01
    hot = ctl.solve_batch(x1, ref, u_prev=up, settings=SolverSettings(polish_passes=3, warm_start=True, **TIGHT))
    fresh = controller(20).solve_batch(x1, ref, u_prev=up)
    assert (hot.status == 1).all() and hot.iters.sum() < 0.6 * fresh.iters.sum()
    assert np.abs(hot.u0 - fresh.u0).max() < 1e-6
    assert cold.iters.sum() > 0


def test_stateless_solve_is_the_same_solve_and_leaves_the_slots_alone():
    """keep_iterate=False (C ABI: warm_start = -1): the iterate back-ups live in per-resident-problem slots instead of the
    problem's own warm-start slot.  Same arithmetic, so the results are identical; a later warm start finds no stored iterate
    and starts cold."""
    from rrt_mpc_b200 import SolverSettings
    g = load_golden("optima.npz")
    x0, ref, up = g["n20_x0"], g["n20_ref"], g["n20_up"]
    kw = dict(polish_passes=5, polish_retry=2, early_polish=True, **TIGHT)
    a = controller(20).solve_batch(x0, ref, u_prev=up, settings=SolverSettings(**kw))
    ctl = controller(20)
    b = ctl.solve_batch(x0, ref, u_prev=up, settings=SolverSettings(keep_iterate=False, **kw))
    for k in ("u0", "Xp", "Up", "iters", "status"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    hot = ctl.solve_batch(x0, ref, u_prev=up, settings=SolverSettings(warm_start=True, **kw))     # nothing was stored: cold
    assert np.array_equal(hot.iters, a.iters)


def _same_solves(a, b, what):
    """Two forms of the iteration on the device.  In the host emulation they are bit-identical (tests/test_emulation.py); on the
    device nvcc contracts a * b + c into fma per expression shape, so iterates may differ in the last bits and, rarely, a polish
    attempt is accepted one check earlier or later: same status, same result to round-off, same iteration count almost always."""
    assert np.array_equal(a.status, b.status), what
    assert (a.iters == b.iters).mean() >= 0.9, what
    assert np.abs(a.iters - b.iters).max() <= 50, what
    assert np.abs(a.u0 - b.u0).max() < 1e-8 and np.abs(a.Xp - b.Xp).max() < 1e-7 and np.abs(a.Up - b.Up).max() < 1e-7, what


@pytest.mark.parametrize("form", ["reg", "pair", "general"])
@pytest.mark.parametrize("name,N,du", [("n50", 50, 0.02), ("n20", 20, 0.15)])
def test_forms_of_the_iteration_agree_on_the_device(form, name, N, du, monkeypatch):
    """The library picks the form of the ADMM iteration from the horizon (short: a lane per stage; general: one parity of stages
    at a time); the pair form and the two-warp register form (mpc_solve_reg_kernel: stage records in registers over a block of
    iterations) are selected with CUDAMPC_FORM.  Every sum has the same operands and order in all of them (bit-identical in the
    host emulation); see _same_solves for what that means on the device."""
    import dataclasses
    from rrt_mpc_b200 import MPCConfig, MPCController, SolverSettings
    g = load_golden("optima.npz")
    par = MPCConfig(horizon=N).to_parameters(0.8)
    par = dataclasses.replace(par, du_bounds=((-12.0, 12.0), (-du, du)))
    x0, ref, up = g[f"{name}_x0"], g[f"{name}_ref"], g[f"{name}_up"]
    st = SolverSettings(polish_passes=5, polish_retry=2, early_polish=True, **TIGHT)
    monkeypatch.delenv("CUDAMPC_FORM", raising=False)
    a = MPCController(par, st, max_batch=len(x0)).solve_batch(x0, ref, u_prev=up)
    monkeypatch.setenv("CUDAMPC_FORM", form)
    b = MPCController(par, st, max_batch=len(x0)).solve_batch(x0, ref, u_prev=up)
    assert (a.status == 1).all()
    _same_solves(a, b, form)
    assert np.abs(a.u0 - g[f"{name}_u0"]).max() < 1e-8


@pytest.mark.parametrize("kw", [dict(check_termination=10, adaptive_rho_interval=35), dict(max_iter=73, polish_passes=1),
                                dict(adaptive_rho=False, max_iter=400)])
def test_register_form_kernel_block_schedule_on_the_device(kw, monkeypatch):
    """mpc_solve_reg_kernel runs the iterations in blocks up to the next event (termination check, rho adaptation, iteration
    limit) and keeps the driver's state in shared memory between the blocks: whatever the schedule of events, it must return what
    the default kernel returns."""
    import dataclasses
    from rrt_mpc_b200 import MPCConfig, MPCController, SolverSettings
    g = load_golden("optima.npz")
    par = dataclasses.replace(MPCConfig(horizon=50).to_parameters(0.8), du_bounds=((-12.0, 12.0), (-0.02, 0.02)))
    x0, ref, up = g["n50_x0"], g["n50_ref"], g["n50_up"]
    base = dict(polish_passes=3, polish_retry=1, early_polish=True, **TIGHT)
    base.update(kw)
    st = SolverSettings(**base)
    monkeypatch.delenv("CUDAMPC_FORM", raising=False)
    a = MPCController(par, st, max_batch=len(x0)).solve_batch(x0, ref, u_prev=up)
    monkeypatch.setenv("CUDAMPC_FORM", "reg")
    b = MPCController(par, st, max_batch=len(x0)).solve_batch(x0, ref, u_prev=up)
    _same_solves(a, b, kw)


def test_invalid_arguments_fail_loudly():
    ctl = controller(20)
    with pytest.raises(ValueError):
        ctl.solve_batch(np.zeros((2, 4)), np.zeros((2, 20, 4)))            # window one row short
    with pytest.raises(ValueError):
        ctl.solve_batch(np.zeros((2, 4)), np.zeros((2, 21, 4)), u_prev=np.zeros((3, 2)))
    from rrt_mpc_b200 import MPCController
    q = np.diag([4.0, 4.0, 0.6, 0.1]); q[0, 1] = q[1, 0] = 5.0            # indefinite
    with pytest.raises(RuntimeError, match="positive semidefinite"):
        MPCController(dataclasses.replace(product_params(20), q=q)).solve(np.zeros(4), np.zeros((21, 4)))


def test_build_reference_batch_matches_reference_producer():
    """K_ref against build_reference outputs of the REAL reference (tests/golden/ref_builder.npz, default scenario)."""
    g = load_golden("ref_builder.npz")
    d = load_golden("default_scenario.npz")
    by_cfg = {}
    for i in range(12):
        v, N, dt = g[f"args{i}"]
        by_cfg.setdefault((float(v), int(N), float(dt)), []).append(i)
    from rrt_mpc_b200 import MPCConfig, MPCController
    for (v, N, dt), idx in by_cfg.items():
        ctl = MPCController(MPCConfig(horizon=N, dt=dt).to_parameters(0.8))
        ref, ref_len = ctl.build_reference_batch([g[f"path{i}"] for i in idx], v)
        ref, ref_len = ref.cpu().numpy(), ref_len.cpu().numpy()
        for k, i in enumerate(idx):
            want = g[f"ref{i}"]
            assert ref_len[k] == len(want), (i, ref_len[k], len(want))
            got = ref[k, :len(want)]
            assert np.abs(got[:, :2] - want[:, :2]).max() <= 1e-12 * np.abs(want[:, :2]).max()
            assert np.abs(got[:, 2] - want[:, 2]).max() <= 1e-12 * max(1.0, np.abs(want[:, 2]).max())
            assert np.abs(got[:, 3] - want[:, 3]).max() <= 1e-12 * v
    ctl = MPCController(MPCConfig().to_parameters(0.8))
    ref, ref_len = ctl.build_reference_batch([d["path"]] * 3, 15.0)
    assert int(ref_len[0]) == 45
    assert np.abs(ref[1, :45].cpu().numpy() - d["ref_global"]).max() <= 1e-12 * 100


def test_f_discrete_hook_within_1e12_of_reference():
    """K_f (the closed loop's integrator) against outputs of the REAL vehicle_model.f_discrete (vehicle_model.py:11-21)."""
    g = load_golden("vehicle_model.npz")
    ctl = controller(20)
    out = ctl.f_discrete_batch(g["x"], g["u"], g["dt_L"])
    scale = np.maximum(1.0, np.abs(g["f"]))
    assert (np.abs(out - g["f"]) <= 1e-12 * scale).all()
    # default (dt, L) of the handle: config.py:79-81 wheelbase 2.8 m / 0.8 m per px, dt 0.1
    from rrt_mpc_b200.vehicle_model import f_discrete
    out = ctl.f_discrete_batch(g["x"][:16], g["u"][:16])
    want = np.stack([f_discrete(g["x"][i], g["u"][i], 0.1, 2.8 / 0.8) for i in range(16)])
    assert (np.abs(out - want) <= 1e-12 * np.maximum(1.0, np.abs(want))).all()


@pytest.mark.parametrize("name,N,du,res", [("unit", 5, 0.15, 0.2), ("roll", 15, 0.15, 0.8), ("n20", 20, 0.15, 0.8), ("n50", 50, 0.02, 0.8)])
def test_matches_what_the_reference_itself_returned(name, N, du, res):
    """tests/golden/ref_qp.npz: return values of the UNMODIFIED reference MPCController.solve (executed through a recording
    cvxpy stand-in and an exact QP solver, tests/golden/make_ref_qp.py).  Bar: u0 within 1e-5 at eps 1e-6."""
    from rrt_mpc_b200 import MPCController, SolverSettings
    g = load_golden("ref_qp.npz")
    ctl = MPCController(product_params(N, du, res), SolverSettings(polish_passes=3, polish_retry=2, **TIGHT), max_batch=128)
    for early in (False, True):
        r = ctl.solve_batch(g[f"{name}_x0"], g[f"{name}_ref"], u_prev=g[f"{name}_up"],
                            settings=SolverSettings(polish_passes=5 if early else 3, polish_retry=2, early_polish=early, **TIGHT))
        assert (r.status == 1).all()
        assert np.abs(r.u0 - g[f"{name}_u0"]).max() < 1e-5
        assert np.abs(r.u0 - g[f"{name}_u0"]).max() < 1e-8
        assert np.abs(r.Xp - g[f"{name}_X"]).max() < 1e-6 and np.abs(r.Up - g[f"{name}_U"]).max() < 1e-6


def test_non_diagonal_weights_match_the_reference():
    """Any PSD q / r / q_terminal (cp.quad_form, mpc_controller.py:74-75,112): the nd20 set of ref_qp.npz."""
    from rrt_mpc_b200 import MPCController, SolverSettings
    g = load_golden("ref_qp.npz")
    p = dataclasses.replace(product_params(20), q=g["nd20_q"], r=g["nd20_r"], q_terminal=g["nd20_qn"])
    ctl = MPCController(p, SolverSettings(polish_passes=3, polish_retry=2, **TIGHT), max_batch=64)
    r = ctl.solve_batch(g["nd20_x0"], g["nd20_ref"], u_prev=g["nd20_up"])
    assert (r.status == 1).all()
    assert np.abs(r.u0 - g["nd20_u0"]).max() < 1e-8
    assert np.abs(r.Xp - g["nd20_X"]).max() < 1e-6 and np.abs(r.Up - g["nd20_U"]).max() < 1e-6
    u0, Xp, Up = ctl.solve(g["nd20_x0"][0], g["nd20_ref"][0], u_prev=g["nd20_up"][0])       # reference signature
    assert np.abs(u0 - g["nd20_u0"][0]).max() < 1e-8


def test_default_settings_regression_bound():
    """MPCController(params).solve() at the reference's hard-coded settings (eps 1e-3, one polish): same loose contract as
    the restated CPU path (tests/test_reference_pinned.py::test_reference_default_settings_regression_bound)."""
    from rrt_mpc_b200 import MPCController, MPCConfig
    g = load_golden("ref_qp.npz")
    ctl = MPCController(MPCConfig().to_parameters(0.8), max_batch=128)                  # default SolverSettings
    r = ctl.solve_batch(g["roll_x0"], g["roll_ref"], u_prev=g["roll_up"])
    assert set(np.unique(r.status)) <= {1}
    err = np.abs(r.u0 - g["roll_u0"])
    assert np.median(err.max(axis=1)) <= 1e-6 and (err.max(axis=1) <= 1e-3).mean() >= 0.6
    assert err[:, 0].max() <= 15.0 and err[:, 1].max() <= 0.1
