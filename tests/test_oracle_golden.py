"""CPU: the oracle (C restatement + NumPy twin) and the host-side mirrors against the golden fixtures that
tests/golden/make_golden.py produced from the REAL reference modules."""
import numpy as np
import pytest

from conftest import load_golden, oracle_params
from oracle import c_oracle as CO
from oracle import mpc_numpy as O


def test_vehicle_model_against_reference_vectors():
    g = load_golden("vehicle_model.npz")
    from rrt_mpc_b200 import vehicle_model as VM
    import ctypes as C
    L = CO.lib()
    for i in range(len(g["x"])):
        dt, Lw = g["dt_L"][i]
        A, B, fx = O.linearize(g["x"][i], g["u"][i], dt, Lw)
        A2, B2, fx2 = VM.linearize(g["x"][i], g["u"][i], dt, Lw)
        for got in ((A, B, fx), (A2, B2, fx2)):
            assert np.array_equal(got[0], g["A"][i]) and np.array_equal(got[1], g["B"][i]) and np.array_equal(got[2], g["fx"][i])
        assert np.array_equal(VM.f_discrete(g["x"][i], g["u"][i], dt, Lw), g["f"][i])
        Ac, Bc, fc = np.zeros(16), np.zeros(8), np.zeros(4)
        x, u = np.ascontiguousarray(g["x"][i]), np.ascontiguousarray(g["u"][i])
        L.oracle_linearize(x.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), C.c_double(dt), C.c_double(Lw),
                           Ac.ctypes.data_as(C.c_void_p), Bc.ctypes.data_as(C.c_void_p), fc.ctypes.data_as(C.c_void_p))
        scale = max(1.0, np.abs(g["A"][i]).max())
        assert np.abs(Ac.reshape(4, 4) - g["A"][i]).max() <= 1e-12 * scale
        assert np.abs(Bc.reshape(4, 2) - g["B"][i]).max() <= 1e-12 * max(1.0, np.abs(g["B"][i]).max())
        assert np.abs(fc - g["fx"][i]).max() <= 1e-12 * max(1.0, np.abs(g["fx"][i]).max())


def test_survey_known_vector():
    A, B, fx = O.linearize(np.array([1.0, 2.0, 0.2, 8.0]), np.array([1.0, 0.1]), 0.1, 5.0)
    assert A[0, 2] == pytest.approx(-0.158935464636049, rel=1e-14)
    assert B[2, 1] == pytest.approx(0.16161072726436151, rel=1e-14)
    assert fx == pytest.approx([1.7840532622729932, 2.158935464636049, 0.2160535475336721, 8.1], rel=1e-14)


@pytest.mark.parametrize("name,N", [("n20", 20), ("n50", 50)])
def test_linearize_window_against_reference(name, N):
    g = load_golden("linearize_window.npz")
    p = oracle_params(N)
    import emu_driver as E
    Ae, Be, ce = E.linearize(p, g[f"{name}_ref"])
    for b in range(len(g[f"{name}_ref"])):
        _, As, Bs, cs = O.linearize_window(g[f"{name}_ref"][b], p)
        _, Ac, Bc, cc = CO.linearize_window(p, g[f"{name}_ref"][b])
        GA, GB, Gc = g[f"{name}_A"][b], g[f"{name}_B"][b], g[f"{name}_c"][b]
        assert np.array_equal(As, GA) and np.array_equal(Bs, GB) and np.array_equal(cs, Gc)    # twin == reference
        sa = np.abs(GA).max()
        for A_, B_, c_ in ((Ac, Bc, cc), (Ae[b], Be[b], ce[b])):                                # C oracle, kernel code on host
            assert np.abs(A_ - GA).max() <= 1e-12 * sa
            assert np.abs(B_ - GB).max() <= 1e-12 * np.abs(GB).max()
            assert np.abs(c_ - Gc).max() <= 1e-12 * max(sa, np.abs(g[f"{name}_ref"][b][:, :2]).max())   # c cancels |X| ~ 1e2


def test_ref_builder_against_reference():
    from rrt_mpc_b200.ref_builder import build_reference
    g = load_golden("ref_builder.npz")
    for i in range(12):
        v, N, dt = g[f"args{i}"]
        got = build_reference([tuple(p) for p in g[f"path{i}"]], float(v), int(N), float(dt))
        assert got.shape == g[f"ref{i}"].shape
        assert np.array_equal(got, g[f"ref{i}"])
    d = load_golden("default_scenario.npz")
    rg = build_reference([tuple(p) for p in d["path"]], 15.0, 15, 0.1)
    assert np.array_equal(rg, d["ref_global"])
    assert rg.shape == (45, 4) and rg[0, 2] == 0.0 and rg[1, 2] == pytest.approx(-2.0839441, abs=1e-6)


@pytest.mark.parametrize("N", [5, 15, 20, 50])
def test_qp_shape_and_twin_agreement(N):
    p = oracle_params(N)
    rng = np.random.default_rng(N)
    ref = np.cumsum(rng.normal(size=(N + 1, 4)) * [1, 1, 0.05, 0.1], axis=0) + [50, 50, 0.3, 12]
    x0, up = ref[0] + rng.normal(size=4) * 0.3, rng.normal(size=2) * 0.1
    P, q, A, l, u = CO.qp_dense(p, x0, ref, up)
    Pn, qn, An, ln, un, lay = O.build_qp(x0, ref, up, p)
    assert P.shape == (11 * N + 5,) * 2 and A.shape == (19 * N + 7, 11 * N + 5)          # n, m
    assert np.count_nonzero(A) == 43 * N + 5 and np.count_nonzero(P) == 11 * N + 5        # nnz(A), nnz(P) (diagonal weights)
    assert np.array_equal(P, Pn.toarray()) and np.array_equal(A, An.toarray())
    assert np.array_equal(q, qn)
    # c_k = fx - A @ xlin cancels |X| ~ 1e2: BLAS vs sequential summation differ by an ulp of |X|
    assert np.abs(l - ln).max() <= 1e-12 * 100 and np.abs(u - un).max() <= 1e-12 * 100
    assert int((l == u).sum()) == 4 + 4 * N                                               # equalities


def test_unit_test_case_of_the_reference():
    """tests/test_mpc_controller.py:7-17 input: outputs exist and Xp[0,1] > x0[0]; plus the certified optimum."""
    g = load_golden("optima.npz")
    p = O.Params(horizon=5, wheelbase_px=2.8 / 0.2)
    r = CO.solve_batch(p, g["unit_x0"], g["unit_ref"][None], None)                       # reference settings (eps 1e-3)
    assert r["status"][0] == 1 and r["Xp"][0][0, 1] > g["unit_x0"][0]
    assert r["Xp"][0][0, 1] == pytest.approx(0.5, abs=1e-9)
    assert np.abs(r["u0"][0] - g["unit_u0"]).max() < 1e-5
    assert g["unit_u0"] == pytest.approx([-9.10462245, 0.0], abs=1e-7)


@pytest.mark.parametrize("name,N,du", [("n20", 20, 0.15), ("n50", 50, 0.02)])
def test_oracle_reaches_certified_optimum(name, N, du):
    """Restated OSQP (Ruiz scaling 10, eps 1e-6, polish) vs the independent KKT-Newton optimum."""
    g = load_golden("optima.npz")
    p = oracle_params(N, du)
    nb = 12
    r = CO.solve_batch(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], eps_abs=1e-6, eps_rel=1e-6, polish_passes=3)
    assert (r["status"] == 1).all()
    assert np.abs(r["u0"] - g[f"{name}_u0"][:nb]).max() < 1e-5
    assert np.abs(r["Xp"] - g[f"{name}_X"][:nb]).max() < 1e-4
    # single polish pass (literal OSQP): still a solved status; most problems already within 1e-5
    r1 = CO.solve_batch(p, g[f"{name}_x0"][:nb], g[f"{name}_ref"][:nb], g[f"{name}_up"][:nb], eps_abs=1e-6, eps_rel=1e-6)
    assert (r1["status"] == 1).all()
    assert np.median(np.abs(r1["u0"] - g[f"{name}_u0"][:nb]).max(axis=1)) < 1e-6


def test_c_oracle_matches_numpy_twin_iteration_for_iteration():
    g = load_golden("optima.npz")
    p = oracle_params(20)
    for b in range(3):
        for kw_c, kw_n in ((dict(scaling=10), dict(scaling=10)), (dict(scaling=0, z0_projected=1), dict(scaling=0, z0_projected=True))):
            r = CO.solve_batch(p, g["n20_x0"][b], g["n20_ref"][b][None], g["n20_up"][b][None], eps_abs=1e-6, eps_rel=1e-6, **kw_c)
            u0, X, U, res = O.mpc_solve(g["n20_x0"][b], g["n20_ref"][b], g["n20_up"][b], p, O.Settings(eps_abs=1e-6, eps_rel=1e-6, **kw_n), info=True)
            assert r["iters"][0] == res.iters and r["info"][0, 0] == res.rho_updates
            assert np.abs(r["u0"][0] - u0).max() < 1e-9


def test_oracle_track_default_scenario_reaches_goal():
    """Closed loop on the default config inputs (control_stage.py:74-157): goal radius reached in ~63 steps."""
    d = load_golden("default_scenario.npz")
    from rrt_mpc_b200.control_stage import initial_state
    p = O.Params(horizon=15)
    s0 = initial_state([tuple(q) for q in d["path"]], d["start"])
    assert s0[2] == pytest.approx(-2.0793, abs=1e-3) and s0[3] == 5.0
    r = CO.track(p, d["ref_global"], s0, d["goal"], 300, eps_abs=1e-6, eps_rel=1e-6, polish_passes=3)
    assert r["flags"] == 1 and 50 <= r["n_steps"] <= 80
    last = r["states"][r["n_steps"] - 1]
    assert np.hypot(last[0] - d["goal"][0], last[1] - d["goal"][1]) < 8.0


def test_exact_optimum_certificate_recovers_the_reference_optima():
    """tests/certificate.py (run on EVERY problem of the full-size GPU batches): from a perturbed candidate it must return
    the optimum the reference itself returned (tests/golden/ref_qp.npz), for diagonal and non-diagonal weights."""
    from certificate import exact_optimum, lin_from_hook
    g = load_golden("ref_qp.npz")
    import dataclasses
    for name, N, du in (("n20", 20, 0.15), ("n50", 50, 0.02), ("roll", 15, 0.15), ("nd20", 20, 0.15)):
        p = oracle_params(N, du)
        if name == "nd20":
            p = dataclasses.replace(p, q=g["nd20_q"], r=g["nd20_r"], q_terminal=g["nd20_qn"])
        x0, ref, up, U = g[f"{name}_x0"], g[f"{name}_ref"], g[f"{name}_up"], g[f"{name}_U"]
        lins = [O.linearize_window(ref[b], p) for b in range(len(ref))]
        refu = np.stack([l[0] for l in lins]); A = np.stack([l[1] for l in lins]); Bm = np.stack([l[2] for l in lins])
        rng = np.random.default_rng(0)
        for scale in (0.0, 1e-6, 1e-2):
            c = exact_optimum(p, x0, refu, up, U + rng.normal(size=U.shape) * scale, lin_from_hook(A, Bm))
            assert c["settled"].all()
            assert np.abs(c["U"] - U).max() < 1e-8 and np.abs(c["X"] - g[f"{name}_X"]).max() < 1e-7
            assert c["newton_steps"].max() <= (1 if scale == 0.0 else 2 if scale == 1e-6 else 40)
