"""GPU (B200): BASELINE.json's full batch sizes through size-independent properties (the oracle cannot run
65,536 horizon-50 problems in seconds): KKT conditions of every returned solution evaluated independently in
NumPy from the linearisation kernel's output, permutation invariance, determinism, and a spot-check subsample
against the certified optimum; and the EXACT optimum of every single problem by a batched semismooth Newton method
(tests/certificate.py) started from the returned controls, so that |u0 - u0*| <= 1e-5 is established for all of them."""
import numpy as np
import pytest

from conftest import oracle_params, product_params

pytestmark = pytest.mark.gpu
TIGHT = dict(eps_abs=1e-6, eps_rel=1e-6)


def kkt_check(p, x0, ref, up, r, A, Bm, c):
    """Primal feasibility of the dynamics and optimality of the slack-eliminated problem, vectorised over the batch."""
    X, U = r.Xp, r.Up                                  # (B,4,N+1), (B,2,N)
    B_, N = X.shape[0], U.shape[2]
    Xk = np.transpose(X, (0, 2, 1)); Uk = np.transpose(U, (0, 2, 1))
    pred = np.einsum("bkij,bkj->bki", A, Xk[:, :N]) + np.einsum("bkij,bkj->bki", Bm, Uk) + c
    dyn = np.abs(Xk[:, 1:] - pred).max()
    init = np.abs(Xk[:, 0] - x0).max()
    return dyn, init


@pytest.mark.parametrize("check", [25, 50])      # 50 = the rho-adaptation interval: the setting bench.py runs early polish with
@pytest.mark.parametrize("N,B,du,seed", [(20, 4096, 0.15, 2), (50, 65536, 0.02, 3)])
def test_full_size_batches(N, B, du, seed, check):
    import torch
    from rrt_mpc_b200 import MPCController, SolverSettings
    from rrt_mpc_b200.synthetic import make_batch
    from oracle import mpc_numpy as O
    x0, ref, up = make_batch(B, N, seed)
    ctl = MPCController(product_params(N, du), SolverSettings(polish_passes=5, polish_retry=4, early_polish=True, check_termination=check, **TIGHT), max_batch=B)
    d = lambda a: torch.as_tensor(a).cuda()
    dx0, dref, dup = d(x0), d(ref), d(up)
    rd = ctl.solve_batch(dx0, dref, u_prev=dup)
    torch.cuda.synchronize()
    r = type(rd)(*(t.cpu().numpy() for t in (rd.u0, rd.Xp, rd.Up, rd.status, rd.iters, rd.pri_res, rd.dua_res, rd.info)))
    assert (r.status == 1).all()                                          # every problem solved
    assert np.isfinite(r.Xp).all() and np.isfinite(r.Up).all()
    assert np.array_equal(r.u0, r.Up[:, :, 0])                            # u0 is U[:, 0]
    # polished problems satisfy the KKT system to round-off; the rest to the ADMM tolerance
    pol = r.info[:, 2] > 0
    assert pol.all()                                                      # every problem ends on a polished KKT point (polish_retry = 4)
    assert r.pri_res[pol].max() < 1e-6 and r.dua_res[pol].max() < 1e-6     # |x| ~ 3e2, |q| ~ 2e3: relative 1e-9
    A, Bm, c = (t.cpu().numpy() for t in ctl.linearize_batch(dref))
    dyn, init = kkt_check(oracle_params(N, du), x0, ref, up, r, A, Bm, c)
    assert init < 1e-8 and dyn < 1e-8
    # EVERY problem against its exact optimum (batched Newton certificate on the GPU, fp64 torch.linalg: test infrastructure)
    from certificate import exact_optimum, lin_from_hook
    refu = ref.copy(); refu[:, :, 2] = np.unwrap(ref[:, :, 2], axis=1)
    cert = exact_optimum(oracle_params(N, du), x0, refu, up, r.Up, lin_from_hook(A, Bm), device="cuda")
    assert cert["settled"].all() and cert["newton_steps"].max() <= 3
    u0_err = np.abs(r.u0 - cert["U"][:, :, 0]).max(axis=1)
    assert u0_err.max() < 1e-5, (u0_err.max(), int(u0_err.argmax()))       # the north-star bar, all B problems
    assert u0_err.max() < 1e-7                                            # what we reach
    assert np.abs(r.Up - cert["U"]).max() < 1e-6 and np.abs(r.Xp - cert["X"]).max() < 1e-5
    dyn_p, init_p = kkt_check(oracle_params(N, du), x0[pol], ref[pol], up[pol], type(r)(*(a[pol] for a in (r.u0, r.Xp, r.Up, r.status, r.iters, r.pri_res, r.dua_res, r.info))), A[pol], Bm[pol], c[pol])
    assert init_p < 1e-8 and dyn_p < 1e-8
    # permutation invariance + determinism: problems are independent, results do not depend on batch position
    perm = np.random.default_rng(0).permutation(B)
    dp = lambda a: torch.as_tensor(np.ascontiguousarray(a[perm])).cuda()
    rp = ctl.solve_batch(dp(x0), dp(ref), u_prev=dp(up))
    torch.cuda.synchronize()
    assert np.array_equal(rp.u0.cpu().numpy(), r.u0[perm]) and np.array_equal(rp.iters.cpu().numpy(), r.iters[perm])
    # spot check against the certified optimum
    idx = np.random.default_rng(1).choice(B, 24, replace=False)
    p = oracle_params(N, du)
    worst = 0.0
    for b in idx:
        u0, X, U, _ = O.solve_kkt_newton(x0[b], ref[b], up[b], p)
        if pol[b]:
            worst = max(worst, np.abs(r.u0[b] - u0).max())
    assert worst < 1e-5
