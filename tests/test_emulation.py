"""CPU: the kernel's per-problem driver (rrt_mpc_b200/csrc/mpc_solve.h), compiled for the host and run lane by
lane (tests/emu), against the oracle.  This is the same source the CUDA kernel compiles; it checks the index
logic, the slack-eliminated banded solve, polish, warm start and edge cases where no GPU is available."""
import dataclasses

import numpy as np
import pytest

import emu_driver as E
from conftest import load_golden, oracle_params
from oracle import c_oracle as CO
from oracle import mpc_numpy as O

TIGHT = dict(eps_abs=1e-6, eps_rel=1e-6)


@pytest.mark.parametrize("name,N,du", [("n20", 20, 0.15), ("n50", 50, 0.02)])
def test_matches_certified_optimum(name, N, du):
    g = load_golden("optima.npz")
    p = oracle_params(N, du)
    nb = 16
    r = E.solve(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], polish_passes=3, **TIGHT)
    assert (r["status"] == 1).all()
    assert np.abs(r["u0"] - g[f"{name}_u0"][:nb]).max() < 1e-8          # bar: 1e-5
    assert np.abs(r["Xp"] - g[f"{name}_X"][:nb]).max() < 1e-7
    assert np.abs(r["Up"] - g[f"{name}_U"][:nb]).max() < 1e-7


def test_iteration_identical_to_oracle_in_mirror_mode():
    """Same ADMM (unscaled, z0 = clip(0)) solved two unrelated ways (sparse KKT LDL' vs slack-eliminated band)."""
    g = load_golden("optima.npz")
    for name, N, du in (("n20", 20, 0.15), ("n50", 50, 0.02)):
        p = oracle_params(N, du)
        nb = 6
        r = E.solve(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], polish_passes=1, **TIGHT)
        c = CO.solve_batch(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], scaling=0, z0_projected=1, **TIGHT)
        assert np.array_equal(r["iters"], c["iters"])
        assert np.array_equal(r["status"], c["status"])
        assert np.array_equal(r["info"][:, 0], c["info"][:, 0])          # rho updates
        assert np.array_equal(r["info"][:, 2], c["info"][:, 2])          # polish accepted / rejected
        assert np.abs(r["u0"] - c["u0"]).max() < 1e-9


@pytest.mark.parametrize("check", [25, 50])      # 50 = the rho-adaptation interval: the setting bench.py runs early polish with
def test_early_polish_certifies_the_same_optimum(check):
    g = load_golden("optima.npz")
    for name, N, du in (("n20", 20, 0.15), ("n50", 50, 0.02)):
        p = oracle_params(N, du)
        nb = 12
        lit = E.solve(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], polish_passes=5, polish_retry=2, **TIGHT)
        ear = E.solve(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], polish_passes=5, polish_retry=2, early_polish=1,
                      check_termination=check, **TIGHT)
        assert (ear["status"] == 1).all() and (ear["info"][:, 2] > 0).all()
        assert np.abs(ear["u0"] - g[f"{name}_u0"][:nb]).max() < 1e-8
        assert np.abs(ear["u0"] - lit["u0"]).max() < 1e-9
        assert ear["iters"].sum() < 0.75 * lit["iters"].sum()


def test_stage_order_hazard_check():
    """Stage-parallel phases must not depend on the order stages are visited in (what a warp does concurrently)."""
    g = load_golden("optima.npz")
    p = oracle_params(20)
    a = E.solve(p, g["n20_x0"][:4], g["n20_ref"][:4], g["n20_up"][:4], polish_passes=3, **TIGHT)
    b = E.solve(p, g["n20_x0"][:4], g["n20_ref"][:4], g["n20_up"][:4], polish_passes=3, reverse=1, **TIGHT)
    for k in ("u0", "Xp", "Up", "iters", "status"):
        assert np.array_equal(a[k], b[k])


def test_reference_unit_case_and_defaults():
    g = load_golden("optima.npz")
    p = O.Params(horizon=5, wheelbase_px=2.8 / 0.2)
    r = E.solve(p, g["unit_x0"], g["unit_ref"][None], None)             # eps 1e-3, polish, as the reference calls OSQP
    assert r["status"][0] == 1 and r["Xp"][0][0, 1] > 0.0
    assert np.abs(r["u0"][0] - g["unit_u0"]).max() < 1e-5
    assert r["Xp"][0][:, 0] == pytest.approx(g["unit_x0"], abs=1e-9)     # X_0 = x0


@pytest.mark.parametrize("N", [1, 2, 3, 7, 33, 64])
def test_horizon_edge_cases(N):
    p = oracle_params(N)
    rng = np.random.default_rng(N)
    ref = np.zeros((3, N + 1, 4))
    for b in range(3):
        yaw = 0.4 * b + np.cumsum(rng.normal(size=N + 1) * 0.05)
        ref[b, :, 2] = yaw
        ref[b, :, 3] = 12.0 + rng.normal(size=N + 1)
        ref[b, :, 0] = 80 + np.cumsum(1.5 * np.cos(yaw)); ref[b, :, 1] = 60 + np.cumsum(1.5 * np.sin(yaw))
    x0 = ref[:, 0] + rng.normal(size=(3, 4)) * [0.5, 0.5, 0.05, 1.0]
    up = rng.uniform(-1, 1, size=(3, 2)) * [3.0, 0.1]
    r = E.solve(p, x0, ref, up, polish_passes=3, **TIGHT)
    assert (r["status"] == 1).all()
    for b in range(3):
        u0, X, U, _ = O.solve_kkt_newton(x0[b], ref[b], up[b], p)
        assert np.abs(r["u0"][b] - u0).max() < 1e-6
        assert np.abs(r["Xp"][b] - X).max() < 1e-6


def test_active_limits_everywhere():
    """Saturated problem: speed far above v_hi, reference demanding a hard turn -> v, u and du slacks all active."""
    N = 12
    p = dataclasses.replace(oracle_params(N, 0.02), v_bounds=(0.0, 10.0), u_bounds=((-2.0, 2.0), (-0.1, 0.1)))
    ref = np.zeros((N + 1, 4)); ref[:, 0] = np.linspace(0, 30, N + 1); ref[:, 1] = np.linspace(0, 25, N + 1) ** 1.2
    ref[:, 2] = np.linspace(0, 2.5, N + 1); ref[:, 3] = 25.0
    x0 = np.array([0.0, 1.0, -0.3, 20.0]); up = np.array([1.5, 0.09])
    r = E.solve(p, x0, ref[None], up[None], polish_passes=3, **TIGHT)
    u0, X, U, sl = O.solve_kkt_newton(x0, ref, up, p)
    assert (sl > 1e-6).sum() >= 10
    assert r["status"][0] == 1
    assert np.abs(r["u0"][0] - u0).max() < 1e-6 and np.abs(r["Xp"][0] - X).max() < 1e-5


def test_yaw_wrap_in_window():
    g = load_golden("optima.npz")
    p = oracle_params(20)
    ref = g["n20_ref"][:3].copy()
    ref[:, :, 2] = ((ref[:, :, 2] + 2.9 + np.pi) % (2 * np.pi)) - np.pi          # wrapped representation of shifted headings
    x0 = ref[:, 0].copy(); x0[:, 2] = np.unwrap(ref[:, :, 2], axis=1)[:, 0] + 0.02
    r = E.solve(p, x0, ref, None, polish_passes=3, **TIGHT)
    for b in range(3):
        u0, X, U, _ = O.solve_kkt_newton(x0[b], ref[b], None, p)
        assert np.abs(r["u0"][b] - u0).max() < 1e-6


def test_status_max_iter_and_inaccurate():
    g = load_golden("optima.npz")
    p = oracle_params(20)
    r = E.solve(p, g["n20_x0"][:2], g["n20_ref"][:2], g["n20_up"][:2], max_iter=10, **TIGHT)
    assert (r["status"] == -2).all() and (r["iters"] == 10).all()          # OSQP_MAX_ITER_REACHED
    c = CO.solve_batch(p, g["n20_x0"][:2], g["n20_ref"][:2], g["n20_up"][:2], scaling=0, z0_projected=1, max_iter=10, **TIGHT)
    assert np.array_equal(c["status"], r["status"])


def test_warm_start_converges_faster_to_same_optimum():
    g = load_golden("optima.npz")
    p = oracle_params(20)
    nb = 4
    warm = np.zeros((nb, E.warm_size(20)))
    cold = E.solve(p, g["n20_x0"][:nb], g["n20_ref"][:nb], g["n20_up"][:nb], warm=warm, warm_start=0, polish_passes=3, **TIGHT)
    x0b = g["n20_x0"][:nb] + 0.01                                          # next closed-loop step: slightly moved state
    hot = E.solve(p, x0b, g["n20_ref"][:nb], g["n20_up"][:nb], warm=warm, warm_start=1, polish_passes=3, **TIGHT)
    ref = E.solve(p, x0b, g["n20_ref"][:nb], g["n20_up"][:nb], polish_passes=3, **TIGHT)
    assert (hot["status"] == 1).all()
    assert hot["iters"].sum() < 0.6 * ref["iters"].sum()
    assert np.abs(hot["u0"] - ref["u0"]).max() < 1e-7
    assert cold["iters"].sum() == E.solve(p, g["n20_x0"][:nb], g["n20_ref"][:nb], g["n20_up"][:nb], polish_passes=3, **TIGHT)["iters"].sum()


def test_non_diagonal_weights_against_reference_outputs():
    """cp.quad_form accepts any PSD Q / R / Q_N (mpc_controller.py:74-75,112); fixture nd20 of tests/golden/ref_qp.npz is
    what the reference's own solve() returned for such weights."""
    g = load_golden("ref_qp.npz")
    p = dataclasses.replace(O.Params(horizon=20), q=g["nd20_q"], r=g["nd20_r"], q_terminal=g["nd20_qn"])
    r = E.solve(p, g["nd20_x0"], g["nd20_ref"], g["nd20_up"], polish_passes=3, **TIGHT)
    assert (r["status"] == 1).all()
    assert np.abs(r["u0"] - g["nd20_u0"]).max() < 1e-8
    assert np.abs(r["Xp"] - g["nd20_X"]).max() < 1e-7 and np.abs(r["Up"] - g["nd20_U"]).max() < 1e-7
    c = CO.solve_batch(p, g["nd20_x0"], g["nd20_ref"], g["nd20_up"], scaling=0, z0_projected=1, **TIGHT)
    assert np.array_equal(r["iters"], c["iters"]) and np.array_equal(r["status"], c["status"])


def test_warm_start_from_a_never_solved_slot_starts_cold():
    """A zeroed warm buffer (what cudampc_create allocates) must not be read as an iterate with rho = 0."""
    g = load_golden("optima.npz")
    p = oracle_params(20)
    nb = 3
    warm = np.zeros((nb, E.warm_size(20)))
    for adaptive in (1, 0):
        warm[:] = 0.0
        hot = E.solve(p, g["n20_x0"][:nb], g["n20_ref"][:nb], g["n20_up"][:nb], warm=warm, warm_start=1, polish_passes=3, adaptive_rho=adaptive, **TIGHT)
        cold = E.solve(p, g["n20_x0"][:nb], g["n20_ref"][:nb], g["n20_up"][:nb], polish_passes=3, adaptive_rho=adaptive, **TIGHT)
        assert np.array_equal(hot["iters"], cold["iters"]) and np.array_equal(hot["status"], cold["status"])
        assert np.array_equal(hot["u0"], cold["u0"])
    warm[:, -1] = np.nan                                     # a poisoned slot is ignored as well
    hot = E.solve(p, g["n20_x0"][:nb], g["n20_ref"][:nb], g["n20_up"][:nb], warm=warm, warm_start=1, polish_passes=3, **TIGHT)
    cold = E.solve(p, g["n20_x0"][:nb], g["n20_ref"][:nb], g["n20_up"][:nb], polish_passes=3, **TIGHT)
    assert np.array_equal(hot["iters"], cold["iters"])


def test_short_and_general_form_of_the_phases_agree():
    """Horizons with N+1 <= 32 run the ADMM phases as one pass over all stages plus two light parity steps
    (solve_problem<true>), longer ones one parity of stages at a time (solve_problem<false>): the same arithmetic per stage,
    so iterates agree to the last bit."""
    g = load_golden("optima.npz")
    p = oracle_params(20)
    nb = 6
    kw = dict(polish_passes=5, polish_retry=2, early_polish=1, **TIGHT)
    try:
        E.set_form(1)
        a = E.solve(p, g["n20_x0"][:nb], g["n20_ref"][:nb], g["n20_up"][:nb], **kw)
        ar = E.solve(p, g["n20_x0"][:nb], g["n20_ref"][:nb], g["n20_up"][:nb], reverse=1, **kw)
        E.set_form(0)
        b = E.solve(p, g["n20_x0"][:nb], g["n20_ref"][:nb], g["n20_up"][:nb], **kw)
    finally:
        E.set_form(-1)
    for k in ("u0", "Xp", "Up", "iters", "status"):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], ar[k])
    assert np.abs(a["u0"] - g["n20_u0"][:nb]).max() < 1e-8


@pytest.mark.parametrize("N,du", [(50, 0.02), (20, 0.15), (15, 0.15), (33, 0.05), (63, 0.05), (2, 0.15), (1, 0.15)])
@pytest.mark.parametrize("form", [2, 3, 7])
def test_pair_and_general_form_of_the_phases_agree(N, du, form):
    """Horizons with N+1 <= 64 run an iteration as ONE pass with a lane per pair of stages (mpc_pair.h: update of iteration
    i and right-hand side of iteration i+1 fused, neighbours' values exchanged between lanes).  Every sum has the operands
    and the order of the general parity form, so the iterates agree to the last bit; the lanes of a part may run in either
    order (what a warp does concurrently).  Even / odd N: the terminal stage is the even / the odd stage of the last lane.
    Form 3 (mpc_reg.h) keeps every stage record in the registers of one lane for a block of iterations, form 7 is the same
    two-warp schedule with the state left in the records: same arithmetic again."""
    from rrt_mpc_b200.synthetic import make_batch
    p = oracle_params(N, du)
    nb = 4
    x0, ref, up = make_batch(nb, N, seed=11)
    kw = dict(polish_passes=5, polish_retry=2, early_polish=1, **TIGHT)
    try:
        E.set_form(form)
        a = E.solve(p, x0, ref, up, **kw)
        ar = E.solve(p, x0, ref, up, reverse=1, **kw)
        E.set_form(0)
        b = E.solve(p, x0, ref, up, **kw)
    finally:
        E.set_form(-1)
    assert (a["status"] == 1).all()
    for k in ("u0", "Xp", "Up", "iters", "status", "info"):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], ar[k]), k


def test_saved_factor_equals_refactorisation_on_resume():
    """A polish factorises in the place of the ADMM factor.  When ADMM resumes after a rejected / unsettled polish the factor
    comes back from a side buffer (rho is unchanged) instead of being recomputed: the same numbers, so the solves agree to the
    last bit; only the factorisation count drops."""
    g = load_golden("optima.npz")
    for name, N, du, form in (("n20", 20, 0.15, -1), ("n50", 50, 0.02, 0), ("n50", 50, 0.02, 3)):
        p = oracle_params(N, du)
        nb = 16
        kw = dict(polish_passes=5, polish_retry=2, early_polish=1, **TIGHT)
        try:
            E.set_form(form)
            E.set_fsave(1)
            a = E.solve(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], **kw)
            E.set_fsave(0)
            b = E.solve(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], **kw)
        finally:
            E.set_fsave(1)
            E.set_form(-1)
        for k in ("u0", "Xp", "Up", "iters", "status"):
            assert np.array_equal(a[k], b[k]), k
        assert a["info"][:, 1].sum() < b["info"][:, 1].sum()          # fewer factorisations


@pytest.mark.parametrize("kw", [
    dict(check_termination=10, adaptive_rho_interval=35),            # events of the two kinds interleave: 10, 20, 30, 35, 40, ...
    dict(max_iter=73, polish_passes=1),                              # iteration limit inside a block, status from the 10x looser test
    dict(check_termination=25, adaptive_rho_interval=0),             # 0 = the library's fixed default interval
    dict(adaptive_rho=0, max_iter=400),                              # no rho adaptation at all
    dict(polish_passes=0),                                           # no polish: the raw ADMM iterate is the answer
])
def test_register_form_block_schedule_matches_the_iteration_loop(kw):
    """The register form runs the iterations in blocks that end at the next event (termination check, rho adaptation, iteration
    limit; mpc_drv.h: drv_prepare).  Whatever the schedule of events, status, iteration counts, rho updates and results must be
    those of the form that tests after every iteration."""
    g = load_golden("optima.npz")
    for name, N, du in (("n20", 20, 0.15), ("n50", 50, 0.02)):
        p = oracle_params(N, du)
        nb = 5
        s = dict(polish_passes=3, polish_retry=1, early_polish=1, **TIGHT)
        s.update(kw)
        try:
            E.set_form(3)
            a = E.solve(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], **s)
            E.set_form(0)
            b = E.solve(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], **s)
        finally:
            E.set_form(-1)
        for k in ("status", "iters", "info", "u0", "Xp", "Up", "pri", "dua"):
            assert np.array_equal(a[k], b[k]), (kw, name, k)


def test_register_form_warm_start_and_slot_state():
    """Warm start through the register form: the iterate stored by one solve is the starting point of the next (same slot), and
    the stored iterate itself equals the one the general form stores."""
    g = load_golden("optima.npz")
    p = oracle_params(20)
    nb = 6
    kw = dict(polish_passes=3, **TIGHT)
    outs = {}
    for form in (3, 0):
        try:
            E.set_form(form)
            warm = np.zeros((nb, E.warm_size(20)))
            cold = E.solve(p, g["n20_x0"][:nb], g["n20_ref"][:nb], g["n20_up"][:nb], warm=warm, warm_start=0, **kw)
            hot = E.solve(p, g["n20_x0"][:nb] + 0.01, g["n20_ref"][:nb], g["n20_up"][:nb], warm=warm, warm_start=1, **kw)
        finally:
            E.set_form(-1)
        outs[form] = (cold, hot, warm.copy())
    for i in (0, 1):
        for k in ("status", "iters", "u0", "Xp", "Up"):
            assert np.array_equal(outs[3][i][k], outs[0][i][k]), (i, k)
    assert np.array_equal(outs[3][2], outs[0][2])
    assert outs[3][1]["iters"].sum() < outs[3][0]["iters"].sum()
