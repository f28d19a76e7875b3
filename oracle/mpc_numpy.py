"""NumPy twin of the CPU oracle (TEST INFRASTRUCTURE — never imported by the product path).

Restates, in plain NumPy/SciPy, the reference's MPC tracking step:

* ``f_discrete`` / ``linearize``      -> /root/reference/src/control/vehicle_model.py:11-45
* horizon QP assembly               -> /root/reference/src/control/mpc_controller.py:47-117
* solver call + status handling     -> /root/reference/src/control/mpc_controller.py:119-141
* closed loop                       -> /root/reference/src/pipeline/control_stage.py:74-157

The solver itself (cvxpy ``>=1.4,<2.0`` -> osqp ``>=0.6.5``, requirements.txt:6-7) is a third-party
dependency that is NOT vendored under /root/reference and is not installable offline, so the ADMM
below restates the published OSQP algorithm (Stellato et al., Math. Prog. Comp. 2020) with the
settings of mpc_controller.py:121-131.  PARITY UNPINNED for the solver part: the reference's own
tests hold no solver-dependent number (tests/test_mpc_controller.py:7-17).  What is pinned: the
linearisation (against the real reference function, tests/golden/) and the optimum itself (the QP is
strictly convex; ``solve_kkt_newton`` below finds its unique minimiser by an independent method).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Optional, Tuple

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

OSQP_INFTY = 1e30
STATUS_SOLVED = 1
STATUS_SOLVED_INACCURATE = 2
STATUS_MAX_ITER = -2
STATUS_UNSOLVED = -10


# --------------------------------------------------------------------------------------
# Parameters (mirror of MPCParameters, mpc_controller.py:17-30; defaults from config.py:66-92)
# --------------------------------------------------------------------------------------
@dataclass
class Params:
    wheelbase_px: float = 2.8 / 0.8
    dt: float = 0.1
    horizon: int = 15
    q: np.ndarray = field(default_factory=lambda: np.diag([4.0, 4.0, 0.6, 0.1]))
    r: np.ndarray = field(default_factory=lambda: np.diag([0.03, 0.25]))
    q_terminal: np.ndarray = field(default_factory=lambda: np.diag([8.0, 8.0, 1.0, 0.2]))
    u_bounds: Tuple[Tuple[float, float], Tuple[float, float]] = ((-35.0, 35.0), (-0.6, 0.6))
    v_bounds: Tuple[float, float] = (0.0, 90.0)
    du_bounds: Tuple[Tuple[float, float], Tuple[float, float]] = ((-12.0, 12.0), (-0.15, 0.15))
    slack_velocity: float = 1e3
    slack_input: float = 5e2
    slack_rate: float = 5e2


@dataclass
class Settings:
    """OSQP settings; first block is what mpc_controller.py:121-131 passes, rest are OSQP defaults."""
    eps_abs: float = 1e-3
    eps_rel: float = 1e-3
    max_iter: int = 60000
    polish: bool = True
    adaptive_rho: bool = True
    rho: float = 0.1
    alpha: float = 1.6
    sigma: float = 1e-6
    scaling: int = 10               # Ruiz passes (OSQP default); 0 = unscaled, the CUDA path's mode
    check_termination: int = 25
    adaptive_rho_interval: int = 50  # fixed (upstream 0.6.x derives it from wall-clock time)
    adaptive_rho_tolerance: float = 5.0
    rho_eq_factor: float = 1e3
    rho_min: float = 1e-6
    rho_max: float = 1e6
    delta: float = 1e-6
    polish_refine_iter: int = 3
    z0_projected: bool = False      # True: start from z0 = clip(0,l,u) (what the merged z/y state does)


# --------------------------------------------------------------------------------------
# Vehicle model (vehicle_model.py:11-45)
# --------------------------------------------------------------------------------------
def f_discrete(x, u, dt, L):
    xk, yk, yaw, v = x
    a, delta = u
    return np.array([xk + dt * v * np.cos(yaw + 0.0), yk + dt * v * np.sin(yaw + 0.0),
                     yaw + dt * (v / L) * np.tan(delta), v + dt * a], dtype=float)


def linearize(x, u, dt, L):
    _, _, yaw, v = x
    _, delta = u
    c, s = np.cos(yaw), np.sin(yaw)
    tan_d = np.tan(delta)
    sec2_d = 1.0 / (np.cos(delta) ** 2 + 1e-9)          # vehicle_model.py:31 (the 1e-9 is contractual)
    A = np.eye(4)
    A[0, 2] = -dt * v * s
    A[0, 3] = dt * c
    A[1, 2] = dt * v * c
    A[1, 3] = dt * s
    A[2, 3] = dt * (1.0 / L) * tan_d
    B = np.zeros((4, 2))
    B[3, 0] = dt
    B[2, 1] = dt * (v / L) * sec2_d
    return A, B, f_discrete(x, u, dt, L)


def linearize_window(ref_traj, p: Params):
    """(A_k, B_k, c_k) for k=0..N-1 as mpc_controller.py:59-70,108-109 computes them.

    Stage k is linearised at ref[max(k-1,0)] of the *unwrapped copy* of the window, with ulin = 0.
    """
    N = p.horizon
    ref = np.array(ref_traj, dtype=float, copy=True)
    ref[:, 2] = np.unwrap(ref[:, 2])
    As, Bs, cs = np.zeros((N, 4, 4)), np.zeros((N, 4, 2)), np.zeros((N, 4))
    xlin = ref[0]
    ulin = np.zeros(2)
    for k in range(N):
        A, B, fx = linearize(xlin, ulin, p.dt, p.wheelbase_px)
        As[k], Bs[k], cs[k] = A, B, fx - A @ xlin - B @ ulin
        xlin = ref[k]
    return ref, As, Bs, cs


# --------------------------------------------------------------------------------------
# QP assembly, stage-interleaved ordering
#   z = [x_0 u_0 sv_0 su_0 sdu_0 | ... | x_{N-1} .. sdu_{N-1} | x_N sv_N],  n = 11N+5
#   rows = [X0=x0 (4)] + per stage k<N: [dyn(4), v(hi,lo,s>=0), u(hi0,lo0,s0,hi1,lo1,s1), du(...6)] + v_N(3)
#   m = 19N+7
# --------------------------------------------------------------------------------------
class Layout:
    def __init__(self, N):
        self.N = N
        self.n = 11 * N + 5
        self.m = 19 * N + 7

    def x(self, k, i):   return 11 * k + i
    def u(self, k, i):   return 11 * k + 4 + i
    def sv(self, k):     return 11 * k + 6 if k < self.N else 11 * self.N + 4
    def su(self, k, i):  return 11 * k + 7 + i
    def sdu(self, k, i): return 11 * k + 9 + i
    # rows
    def r_init(self, i):      return i
    def r_dyn(self, k, i):    return 4 + 19 * k + i
    def r_v(self, k, j):      return 4 + 19 * k + (4 if k < self.N else 0) + j   # j: 0 hi, 1 lo, 2 s>=0
    def r_u(self, k, i, j):   return 4 + 19 * k + 7 + 3 * i + j
    def r_du(self, k, i, j):  return 4 + 19 * k + 13 + 3 * i + j


def build_qp(x0, ref_traj, u_prev, p: Params):
    """Literal standard form of mpc_controller.py:47-117: min 1/2 z'Pz + q'z, l <= Az <= u."""
    N = p.horizon
    lay = Layout(N)
    x0 = np.asarray(x0, dtype=float)
    u_prev = np.zeros(2) if u_prev is None else np.asarray(u_prev, dtype=float)
    ref, As, Bs, cs = linearize_window(ref_traj, p)
    n, m = lay.n, lay.m
    P = sp.lil_matrix((n, n))
    q = np.zeros(n)
    Q, R, QN = np.asarray(p.q, float), np.asarray(p.r, float), np.asarray(p.q_terminal, float)
    Qs, Rs, QNs = Q + Q.T, R + R.T, QN + QN.T            # quad_form(e, Q) = e'Qe = 1/2 e'(Q+Q')e
    for k in range(N + 1):
        W = Qs if k < N else QNs
        for i in range(4):
            for j in range(4):
                if W[i, j] != 0.0:
                    P[lay.x(k, i), lay.x(k, j)] = W[i, j]
        q[[lay.x(k, i) for i in range(4)]] = -W @ ref[k]
        P[lay.sv(k), lay.sv(k)] = 2.0 * p.slack_velocity
    for k in range(N):
        for i in range(2):
            for j in range(2):
                if Rs[i, j] != 0.0:
                    P[lay.u(k, i), lay.u(k, j)] = Rs[i, j]
            P[lay.su(k, i), lay.su(k, i)] = 2.0 * p.slack_input
            P[lay.sdu(k, i), lay.sdu(k, i)] = 2.0 * p.slack_rate
    A = sp.lil_matrix((m, n))
    l = np.full(m, -OSQP_INFTY)
    u = np.full(m, OSQP_INFTY)
    for i in range(4):
        A[lay.r_init(i), lay.x(0, i)] = 1.0
        l[lay.r_init(i)] = u[lay.r_init(i)] = x0[i]
    for k in range(N):
        for i in range(4):
            r = lay.r_dyn(k, i)
            A[r, lay.x(k + 1, i)] = 1.0
            for j in range(4):
                if As[k, i, j] != 0.0:
                    A[r, lay.x(k, j)] = -As[k, i, j]
            for j in range(2):
                if Bs[k, i, j] != 0.0:
                    A[r, lay.u(k, j)] = -Bs[k, i, j]
            l[r] = u[r] = cs[k, i]
    for k in range(N + 1):
        r = lay.r_v(k, 0); A[r, lay.x(k, 3)] = 1.0; A[r, lay.sv(k)] = -1.0; u[r] = p.v_bounds[1]
        r = lay.r_v(k, 1); A[r, lay.x(k, 3)] = 1.0; A[r, lay.sv(k)] = 1.0;  l[r] = p.v_bounds[0]
        r = lay.r_v(k, 2); A[r, lay.sv(k)] = 1.0; l[r] = 0.0
    for k in range(N):
        for i in range(2):
            r = lay.r_u(k, i, 0); A[r, lay.u(k, i)] = 1.0; A[r, lay.su(k, i)] = -1.0; u[r] = p.u_bounds[i][1]
            r = lay.r_u(k, i, 1); A[r, lay.u(k, i)] = 1.0; A[r, lay.su(k, i)] = 1.0;  l[r] = p.u_bounds[i][0]
            r = lay.r_u(k, i, 2); A[r, lay.su(k, i)] = 1.0; l[r] = 0.0
            off = u_prev[i] if k == 0 else 0.0
            r = lay.r_du(k, i, 0); A[r, lay.u(k, i)] = 1.0; A[r, lay.sdu(k, i)] = -1.0; u[r] = p.du_bounds[i][1] + off
            if k > 0: A[r, lay.u(k - 1, i)] = -1.0
            r = lay.r_du(k, i, 1); A[r, lay.u(k, i)] = 1.0; A[r, lay.sdu(k, i)] = 1.0;  l[r] = p.du_bounds[i][0] + off
            if k > 0: A[r, lay.u(k - 1, i)] = -1.0
            r = lay.r_du(k, i, 2); A[r, lay.sdu(k, i)] = 1.0; l[r] = 0.0
    return sp.csc_matrix(P), q, sp.csc_matrix(A), l, u, lay


def extract(z, lay):
    N = lay.N
    X = np.array([[z[lay.x(k, i)] for k in range(N + 1)] for i in range(4)])
    U = np.array([[z[lay.u(k, i)] for k in range(N)] for i in range(2)])
    return X, U


# --------------------------------------------------------------------------------------
# Scaling
# --------------------------------------------------------------------------------------
def ruiz_scaling(P, q, A, iters):
    """OSQP's modified Ruiz equilibration of [[P, A'], [A, 0]] plus cost scaling."""
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P, A, q = P.copy().tocsc(), A.copy().tocsc(), q.copy()
    MINS, MAXS = 1e-4, 1e4
    for _ in range(iters):
        colP = np.asarray(abs(P).max(axis=0).todense()).ravel() if P.nnz else np.zeros(n)
        colA = np.asarray(abs(A).max(axis=0).todense()).ravel()
        rowA = np.asarray(abs(A).max(axis=1).todense()).ravel()
        dn = np.maximum(colP, colA)
        dn = np.where(dn < MINS, 1.0, np.minimum(dn, MAXS))
        en = np.where(rowA < MINS, 1.0, np.minimum(rowA, MAXS))
        dt, et = 1.0 / np.sqrt(dn), 1.0 / np.sqrt(en)
        Dm, Em = sp.diags(dt), sp.diags(et)
        P = (Dm @ P @ Dm).tocsc(); A = (Em @ A @ Dm).tocsc(); q = dt * q
        D *= dt; E *= et
        colP = np.asarray(abs(P).max(axis=0).todense()).ravel()
        cn = colP.mean()
        qn = np.abs(q).max()
        cn = max(cn, qn)
        cn = 1.0 if cn < MINS else min(cn, MAXS)
        ct = 1.0 / cn
        P = P * ct; q = q * ct; c *= ct
    return D, E, c


# --------------------------------------------------------------------------------------
# OSQP restatement
# --------------------------------------------------------------------------------------
@dataclass
class Result:
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    status: int
    iters: int
    pri_res: float
    dua_res: float
    rho_updates: int
    polished: int      # 1 accepted, -1 rejected, 0 not attempted
    rho: float = 0.0


def _ninf(v):
    return np.abs(v).max() if v.size else 0.0


def osqp_solve(P, q, A, l, u, s: Settings, *, scaling=None, warm=None) -> Result:
    n, m = P.shape[0], A.shape[0]
    if scaling is None:
        D, E, c = ruiz_scaling(P, q, A, s.scaling) if s.scaling > 0 else (np.ones(n), np.ones(m), 1.0)
    else:
        D, E, c = scaling
    Dm, Em = sp.diags(D), sp.diags(E)
    Ps = (c * (Dm @ P @ Dm)).tocsc()
    qs = c * D * q
    As = (Em @ A @ Dm).tocsc()
    ls = np.where(l <= -OSQP_INFTY, -OSQP_INFTY, E * l)
    us = np.where(u >= OSQP_INFTY, OSQP_INFTY, E * u)
    Dinv, Einv, cinv = 1.0 / D, 1.0 / E, 1.0 / c
    is_eq = (us - ls) < 1e-4                       # OSQP RHO_TOL
    is_free = (ls <= -OSQP_INFTY) & (us >= OSQP_INFTY)

    def rho_vector(rho):
        rv = np.full(m, rho)
        rv[is_eq] = s.rho_eq_factor * rho
        rv[is_free] = s.rho_min
        return rv

    def factor(rv):
        M = (Ps + s.sigma * sp.eye(n) + As.T @ sp.diags(rv) @ As).tocsc()
        return spla.splu(M)

    rho = s.rho
    rv = rho_vector(rho)
    lu = factor(rv)
    if warm is not None:
        x = warm[0] / D
        y = c * warm[1] / E
        z = As @ x
    else:
        x, y = np.zeros(n), np.zeros(m)
        z = np.clip(np.zeros(m), ls, us) if s.z0_projected else np.zeros(m)
    status, it, n_rho = STATUS_UNSOLVED, 0, 0
    pri = dua = np.inf

    def residuals(x, z, y):
        Ax = As @ x
        Px = Ps @ x
        Aty = As.T @ y
        pri = _ninf(Einv * (Ax - z))
        dua = cinv * _ninf(Dinv * (Px + qs + Aty))
        eps_p = s.eps_abs + s.eps_rel * max(_ninf(Einv * Ax), _ninf(Einv * z))
        eps_d = s.eps_abs + s.eps_rel * cinv * max(_ninf(Dinv * Px), _ninf(Dinv * Aty), _ninf(Dinv * qs))
        # scaled quantities for the rho estimate
        sp_ = _ninf(Ax - z) / (max(_ninf(Ax), _ninf(z)) + 1e-10)
        sd_ = _ninf(Px + qs + Aty) / (max(_ninf(Px), _ninf(Aty), _ninf(qs)) + 1e-10)
        return pri, dua, eps_p, eps_d, sp_, sd_

    for it in range(1, s.max_iter + 1):
        rhs = s.sigma * x - qs + As.T @ (rv * z - y)
        xt = lu.solve(rhs)
        zt = As @ xt
        x = s.alpha * xt + (1 - s.alpha) * x
        w = s.alpha * zt + (1 - s.alpha) * z
        znew = np.clip(w + y / rv, ls, us)
        y = y + rv * (w - znew)
        z = znew
        check = (it % s.check_termination == 0)
        adapt = s.adaptive_rho and (it % s.adaptive_rho_interval == 0)
        if check or adapt:
            pri, dua, eps_p, eps_d, sp_, sd_ = residuals(x, z, y)
            if check and pri <= eps_p and dua <= eps_d:
                status = STATUS_SOLVED
                break
            if adapt:
                rho_new = float(np.clip(rho * np.sqrt(sp_ / (sd_ + 1e-10)), s.rho_min, s.rho_max))
                if rho_new > rho * s.adaptive_rho_tolerance or rho_new < rho / s.adaptive_rho_tolerance:
                    rho = rho_new
                    rv = rho_vector(rho)
                    lu = factor(rv)
                    n_rho += 1
    if status != STATUS_SOLVED:
        pri, dua, eps_p, eps_d, _, _ = residuals(x, z, y)
        s10 = replace(s, eps_abs=10 * s.eps_abs, eps_rel=10 * s.eps_rel)
        eps_p10 = s10.eps_abs + (eps_p - s.eps_abs) * 10
        eps_d10 = s10.eps_abs + (eps_d - s.eps_abs) * 10
        status = STATUS_SOLVED_INACCURATE if (pri <= eps_p10 and dua <= eps_d10) else STATUS_MAX_ITER

    polished = 0
    if s.polish and status == STATUS_SOLVED:
        low = (z - ls) < -y
        upp = (us - z) < y
        act = low | upp
        ia = np.flatnonzero(act)
        Aa = As[ia, :]
        b = np.where(low, ls, us)[ia]
        na = len(ia)
        K = sp.bmat([[Ps, Aa.T], [Aa, None]]).tocsc() if na else Ps
        Kr = sp.bmat([[Ps + s.delta * sp.eye(n), Aa.T], [Aa, -s.delta * sp.eye(na)]]).tocsc() if na else (Ps + s.delta * sp.eye(n)).tocsc()
        klu = spla.splu(Kr)
        rhs = np.concatenate([-qs, b])
        sol = klu.solve(rhs)
        for _ in range(s.polish_refine_iter):
            sol = sol + klu.solve(rhs - K @ sol)
        xp = sol[:n]
        yp = np.zeros(m); yp[ia] = sol[n:]
        zp_ = As @ xp
        zp = np.clip(zp_, ls, us)
        pri_p = _ninf(Einv * (zp_ - zp))
        dua_p = cinv * _ninf(Dinv * (Ps @ xp + qs + As.T @ yp))
        ok = (pri_p < pri and dua_p < dua) or (pri_p < pri and dua < 1e-10) or (dua_p < dua and pri < 1e-10)
        if ok:
            x, y, z, pri, dua, polished = xp, yp, zp, pri_p, dua_p, 1
        else:
            polished = -1
    return Result(x=D * x, y=cinv * E * y, z=Einv * z, status=status, iters=it, pri_res=pri, dua_res=dua,
                  rho_updates=n_rho, polished=polished, rho=rho)


# --------------------------------------------------------------------------------------
# Independent certificate: semismooth Newton on the slack-eliminated piecewise-quadratic problem
# --------------------------------------------------------------------------------------
def solve_kkt_newton(x0, ref_traj, u_prev, p: Params, max_iter=100):
    """Unique minimiser of the reference QP by an active-set/semismooth-Newton method (no ADMM).

    Eliminating each slack analytically (min_s>=0 w s^2 s.t. lo - s <= g <= hi + s) turns every soft
    limit into the C^1 penalty w*dist(g,[lo,hi])^2; with the equality-constrained dynamics this is a
    piecewise-quadratic strictly convex problem whose Newton iteration terminates finitely.
    Returns (u0, X, U, slack dict).
    """
    N = p.horizon
    x0 = np.asarray(x0, float)
    u_prev = np.zeros(2) if u_prev is None else np.asarray(u_prev, float)
    ref, As, Bs, cs = linearize_window(ref_traj, p)
    nv = 4 * (N + 1) + 2 * N
    ix = lambda k, i: 6 * k + i if k < N else 6 * N + i
    iu = lambda k, i: 6 * k + 4 + i
    Q, R, QN = np.asarray(p.q, float), np.asarray(p.r, float), np.asarray(p.q_terminal, float)
    H0 = np.zeros((nv, nv)); g0 = np.zeros(nv)
    for k in range(N + 1):
        W = (Q + Q.T) if k < N else (QN + QN.T)
        idx = [ix(k, i) for i in range(4)]
        H0[np.ix_(idx, idx)] += W
        g0[idx] += -W @ ref[k]
    for k in range(N):
        idx = [iu(k, i) for i in range(2)]
        H0[np.ix_(idx, idx)] += (R + R.T)
    # equalities
    ne = 4 * (N + 1)
    G = np.zeros((ne, nv)); h = np.zeros(ne)
    for i in range(4):
        G[i, ix(0, i)] = 1.0; h[i] = x0[i]
    for k in range(N):
        for i in range(4):
            r = 4 + 4 * k + i
            G[r, ix(k + 1, i)] = 1.0
            G[r, [ix(k, j) for j in range(4)]] -= As[k, i]
            G[r, [iu(k, j) for j in range(2)]] -= Bs[k, i]
            h[r] = cs[k, i]
    # soft rows g = C v in [lo, hi] with weight w
    rows, lo, hi, w = [], [], [], []
    for k in range(N + 1):
        e = np.zeros(nv); e[ix(k, 3)] = 1.0
        rows.append(e); lo.append(p.v_bounds[0]); hi.append(p.v_bounds[1]); w.append(p.slack_velocity)
    for k in range(N):
        for i in range(2):
            e = np.zeros(nv); e[iu(k, i)] = 1.0
            rows.append(e); lo.append(p.u_bounds[i][0]); hi.append(p.u_bounds[i][1]); w.append(p.slack_input)
    for k in range(N):
        for i in range(2):
            e = np.zeros(nv); e[iu(k, i)] = 1.0
            off = u_prev[i] if k == 0 else 0.0
            if k > 0: e[iu(k - 1, i)] = -1.0
            rows.append(e); lo.append(p.du_bounds[i][0] + off); hi.append(p.du_bounds[i][1] + off); w.append(p.slack_rate)
    C = np.array(rows); lo = np.array(lo); hi = np.array(hi); w = np.array(w)
    def objective(v):
        gv = C @ v
        d = np.maximum(0.0, np.maximum(gv - hi, lo - gv))
        return 0.5 * v @ (H0 @ v) + g0 @ v + np.sum(w * d * d)

    def newton_point(act_hi, act_lo):
        Wd = 2.0 * w * (act_hi | act_lo)
        t = np.where(act_hi, hi, np.where(act_lo, lo, 0.0))
        H = H0 + C.T @ (Wd[:, None] * C)
        g = g0 - C.T @ (Wd * t)
        K = np.block([[H, G.T], [G, np.zeros((ne, ne))]])
        rhs = np.concatenate([-g, h])
        sol = np.linalg.solve(K, rhs)
        sol = sol + np.linalg.solve(K, rhs - K @ sol)
        return sol[:nv]

    # feasible start: minimiser with no limit active; then damped (Armijo) semismooth Newton, which
    # terminates finitely on a C^1 piecewise-quadratic strictly convex objective.
    none = np.zeros(len(w), bool)
    v = newton_point(none, none)
    for _ in range(max_iter):
        gv = C @ v
        act_hi, act_lo = gv > hi, gv < lo
        vh = newton_point(act_hi, act_lo)
        gh = C @ vh
        if np.array_equal(gh > hi, act_hi) and np.array_equal(gh < lo, act_lo):
            v = vh
            break
        f0, d, t = objective(v), vh - v, 1.0
        while t > 1e-8 and objective(v + t * d) > f0 - 1e-4 * t * abs(f0 - objective(vh) if objective(vh) < f0 else 0.0):
            t *= 0.5
        v = v + t * d
    else:
        raise RuntimeError("kkt newton did not settle")
    X = np.array([[v[ix(k, i)] for k in range(N + 1)] for i in range(4)])
    U = np.array([[v[iu(k, i)] for k in range(N)] for i in range(2)])
    gv = C @ v
    slack = np.maximum(0.0, np.maximum(gv - hi, lo - gv))
    return U[:, 0].copy(), X, U, slack


# --------------------------------------------------------------------------------------
# MPCController.solve restatement (mpc_controller.py:39-141)
# --------------------------------------------------------------------------------------
def mpc_solve(x0, ref_traj, u_prev, p: Params, s: Optional[Settings] = None, *, scaling=None, info=False):
    s = s or Settings()
    P, q, A, l, u, lay = build_qp(x0, ref_traj, u_prev, p)
    res = osqp_solve(P, q, A, l, u, s, scaling=scaling)
    if res.status not in (STATUS_SOLVED, STATUS_SOLVED_INACCURATE):      # mpc_controller.py:137-139
        return (None, None, None, res) if info else (None, None, None)
    X, U = extract(res.x, lay)
    return (U[:, 0].copy(), X, U, res) if info else (U[:, 0].copy(), X, U)
