"""A recording stand-in for ``cvxpy`` — just enough of its API to EXECUTE the reference's own QP statement.

TEST INFRASTRUCTURE (used by tests/golden/make_golden.py in the build container, where /root/reference exists).
cvxpy and osqp are not installable offline, so the reference's ``MPCController.solve``
(/root/reference/src/control/mpc_controller.py:39-141) cannot run as shipped.  Installed as ``sys.modules["cvxpy"]``
this module lets the UNMODIFIED reference source run: every ``cp.Variable``, index, ``@``, ``+``/``-``, comparison,
``cp.quad_form`` / ``cp.square`` / ``cp.sum_squares``, ``cp.Minimize`` and ``cp.Problem(...).solve(...)`` the
reference executes is recorded as an affine/quadratic expression over the problem's variables, and ``solve`` assembles

    min 1/2 z'Hz + g'z + c0   s.t.  Aeq z = beq,  Ain z <= bin

from exactly what the reference wrote.  It then solves that QP with a dense primal-dual interior-point method
followed by an active-set vertex solve (both below; they share no code with oracle/ or with the CUDA path) and
fills ``Variable.value`` / ``Problem.status`` so that the reference returns ``U.value[:, 0], X.value, U.value``.

Only the API surface mpc_controller.py:53-132 touches is covered; anything else raises.
Variable layout of the recorded QP: variables in creation order (X, U, s_v, s_du, s_u), each flattened row-major.
"""
from __future__ import annotations

import numpy as np

OSQP = "OSQP"
OPTIMAL = "optimal"
OPTIMAL_INACCURATE = "optimal_inaccurate"


class SolverError(Exception):
    pass


RECORDS = []          # one dict per Problem.solve call (make_golden.py drains it)
_COUNTER = [0]


def _as_vec(c):
    return np.atleast_1d(np.asarray(c, dtype=float)).ravel()


class Expr:
    """Affine expression  sum_v M_v vec(v) + c  with a 1-D value (scalars have one row)."""
    __array_ufunc__ = None        # make ndarray (+,-,@,<=,...) Expr defer to the reflected methods below

    def __init__(self, coef, const):
        self.coef = coef          # {Variable: (rows, var.size) ndarray}
        self.const = _as_vec(const)

    @property
    def rows(self):
        return self.const.shape[0]

    @staticmethod
    def wrap(x, rows=None):
        if isinstance(x, Expr):
            return x
        if isinstance(x, Variable):
            return x.expr()
        c = _as_vec(x)
        if rows is not None and c.shape[0] == 1 and rows != 1:
            c = np.full(rows, c[0])
        return Expr({}, c)

    def _broadcast(self, rows):
        if self.rows == rows:
            return self
        if self.rows != 1:
            raise ValueError(f"shape mismatch: {self.rows} vs {rows}")
        return Expr({v: np.repeat(m, rows, axis=0) for v, m in self.coef.items()}, np.repeat(self.const, rows))

    def __add__(self, other):
        o = Expr.wrap(other)
        rows = max(self.rows, o.rows)
        a, b = self._broadcast(rows), o._broadcast(rows)
        coef = {v: m.copy() for v, m in a.coef.items()}
        for v, m in b.coef.items():
            coef[v] = coef[v] + m if v in coef else m.copy()
        return Expr(coef, a.const + b.const)

    __radd__ = __add__

    def __neg__(self):
        return Expr({v: -m for v, m in self.coef.items()}, -self.const)

    def __sub__(self, other):
        return self + (-Expr.wrap(other))

    def __rsub__(self, other):
        return Expr.wrap(other) + (-self)

    def __mul__(self, k):
        k = float(k)
        return Expr({v: k * m for v, m in self.coef.items()}, k * self.const)

    __rmul__ = __mul__

    def __rmatmul__(self, A):                   # ndarray @ Expr
        A = np.atleast_2d(np.asarray(A, dtype=float))
        return Expr({v: A @ m for v, m in self.coef.items()}, A @ self.const)

    def __le__(self, other):
        return Constraint(self - other, "<=")

    def __ge__(self, other):
        return Constraint(Expr.wrap(other) - self, "<=")

    def __eq__(self, other):                    # noqa: D105 - cvxpy semantics: builds a constraint
        return Constraint(self - other, "==")

    __hash__ = None


class Variable:
    __array_ufunc__ = None

    def __init__(self, shape=()):
        self.shape = (shape,) if isinstance(shape, int) else tuple(shape)
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.id = _COUNTER[0]
        _COUNTER[0] += 1
        self.value = None

    def __hash__(self):
        return self.id

    def expr(self):
        return Expr({self: np.eye(self.size)}, np.zeros(self.size))

    def __getitem__(self, idx):
        flat = np.atleast_1d(np.arange(self.size).reshape(self.shape)[idx]).ravel()
        m = np.zeros((flat.shape[0], self.size))
        m[np.arange(flat.shape[0]), flat] = 1.0
        return Expr({self: m}, np.zeros(flat.shape[0]))

    def __add__(self, o): return self.expr() + o
    __radd__ = __add__
    def __sub__(self, o): return self.expr() - o
    def __rsub__(self, o): return o - self.expr()
    def __neg__(self): return -self.expr()
    def __le__(self, o): return self.expr() <= o
    def __ge__(self, o): return self.expr() >= o
    def __eq__(self, o): return self.expr() == o          # noqa: D105
    def __rmatmul__(self, A): return A @ self.expr()


class Constraint:
    def __init__(self, lhs: Expr, kind: str):      # lhs <= 0  or  lhs == 0
        self.lhs, self.kind = lhs, kind


class Quad:
    """sum_i w_i * e_i' Q_i e_i  (+ constant)."""
    __array_ufunc__ = None

    def __init__(self, terms, const=0.0):
        self.terms, self.const = terms, float(const)

    def __add__(self, other):
        if isinstance(other, Quad):
            return Quad(self.terms + other.terms, self.const + other.const)
        if np.isscalar(other):
            return Quad(list(self.terms), self.const + float(other))
        return NotImplemented

    __radd__ = __add__

    def __mul__(self, k):
        return Quad([(float(k) * w, e, Q) for w, e, Q in self.terms], float(k) * self.const)

    __rmul__ = __mul__


def quad_form(x, P):
    e = Expr.wrap(x)
    P = np.asarray(P, dtype=float)
    if P.shape != (e.rows, e.rows):
        raise ValueError("quad_form: shape mismatch")
    return Quad([(1.0, e, P)])


def sum_squares(x):
    e = Expr.wrap(x)
    return Quad([(1.0, e, np.eye(e.rows))])


def square(x):
    e = Expr.wrap(x)
    if e.rows != 1:
        raise NotImplementedError("square(): only scalars are summed into the cost by the reference")
    return Quad([(1.0, e, np.eye(1))])


class Minimize:
    def __init__(self, cost):
        if not isinstance(cost, Quad):
            raise TypeError("objective must be quadratic")
        self.cost = cost


class Problem:
    def __init__(self, objective: Minimize, constraints):
        self.objective, self.constraints = objective, list(constraints)
        self.status = None
        self.value = None

    def _variables(self):
        seen = {}
        for _, e, _ in self.objective.cost.terms:
            for v in e.coef:
                seen[v.id] = v
        for c in self.constraints:
            for v in c.lhs.coef:
                seen[v.id] = v
        return [seen[i] for i in sorted(seen)]

    def assemble(self):
        vs = self._variables()
        off, n = {}, 0
        for v in vs:
            off[v] = n
            n += v.size

        def dense(e: Expr):
            M = np.zeros((e.rows, n))
            for v, m in e.coef.items():
                M[:, off[v]:off[v] + v.size] += m
            return M

        H, g, c0 = np.zeros((n, n)), np.zeros(n), self.objective.cost.const
        for w, e, Q in self.objective.cost.terms:
            M, c = dense(e), e.const
            Qs = 0.5 * (Q + Q.T)
            H += 2.0 * w * (M.T @ Qs @ M)
            g += 2.0 * w * (M.T @ (Qs @ c))
            c0 += w * float(c @ Qs @ c)
        Aeq, beq, Ain, bin_ = [], [], [], []
        for c in self.constraints:
            M = dense(c.lhs)
            (Aeq if c.kind == "==" else Ain).append(M)
            (beq if c.kind == "==" else bin_).append(-c.lhs.const)
        cat = lambda rows, width: np.vstack(rows) if rows else np.zeros((0, width))
        vec = lambda rows: np.concatenate(rows) if rows else np.zeros(0)
        return vs, off, dict(H=H, g=g, c0=c0, Aeq=cat(Aeq, n), beq=vec(beq), Ain=cat(Ain, n), bin=vec(bin_))

    def solve(self, solver=None, **kwargs):
        vs, off, qp = self.assemble()
        z, info = solve_qp(qp["H"], qp["g"], qp["Aeq"], qp["beq"], qp["Ain"], qp["bin"])
        for v in vs:
            v.value = z[off[v]:off[v] + v.size].reshape(v.shape)
        self.status = OPTIMAL
        self.value = float(0.5 * z @ qp["H"] @ z + qp["g"] @ z + qp["c0"])
        RECORDS.append(dict(qp=qp, z=z, solver=solver, kwargs=dict(kwargs), info=info,
                            layout=[(v.shape, off[v]) for v in vs]))
        return self.value


# ------------------------------------------------------------------------------------------------------
# Independent dense QP solver: Mehrotra predictor-corrector interior point, then an active-set vertex solve.
# ------------------------------------------------------------------------------------------------------
def _kkt_residuals(H, g, Aeq, beq, Ain, bin_, z, nu, lam):
    r_d = H @ z + g + Aeq.T @ nu + Ain.T @ lam
    r_e = Aeq @ z - beq
    r_i = np.maximum(Ain @ z - bin_, 0.0)
    comp = lam * (bin_ - Ain @ z)
    return dict(dual=float(np.abs(r_d).max(initial=0.0)), eq=float(np.abs(r_e).max(initial=0.0)),
                ineq=float(r_i.max(initial=0.0)), comp=float(np.abs(comp).max(initial=0.0)),
                lam_min=float(lam.min(initial=0.0)))


def _interior_point(H, g, Aeq, beq, Ain, bin_, tol=1e-11, max_iter=200):
    n, me, mi = H.shape[0], Aeq.shape[0], Ain.shape[0]
    z, nu = np.zeros(n), np.zeros(me)
    s = np.maximum(bin_ - Ain @ z, 1.0)
    lam = np.ones(mi)
    for _ in range(max_iter):
        r_d = H @ z + g + Aeq.T @ nu + Ain.T @ lam
        r_e = Aeq @ z - beq
        r_i = Ain @ z + s - bin_
        mu = float(s @ lam) / max(mi, 1)
        if max(np.abs(r_d).max(initial=0), np.abs(r_e).max(initial=0), np.abs(r_i).max(initial=0), mu) < tol:
            break
        d = lam / s
        K = np.zeros((n + me, n + me))
        K[:n, :n] = H + Ain.T @ (d[:, None] * Ain)
        K[:n, n:] = Aeq.T
        K[n:, :n] = Aeq
        K[n:, n:] = -1e-14 * np.eye(me)

        def step(r_c):
            # eliminate ds = -r_i - Ain dz, dlam = (-r_c - lam ds) / s
            rhs = np.concatenate((-r_d - Ain.T @ ((-r_c + lam * r_i) / s), -r_e))
            sol = np.linalg.solve(K, rhs)
            dz, dnu = sol[:n], sol[n:]
            ds = -r_i - Ain @ dz
            dlam = (-r_c - lam * ds) / s
            return dz, dnu, ds, dlam

        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, float((-v[neg] / dv[neg]).min())) if neg.any() else 1.0

        dz, dnu, ds, dlam = step(s * lam)                               # predictor
        a = min(max_step(s, ds), max_step(lam, dlam))
        mu_aff = float((s + a * ds) @ (lam + a * dlam)) / max(mi, 1)
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        dz, dnu, ds, dlam = step(s * lam + ds * dlam - sigma * mu)      # corrector
        a = 0.995 * min(max_step(s, ds), max_step(lam, dlam))
        a = min(a, 1.0)
        z, nu, s, lam = z + a * dz, nu + a * dnu, s + a * ds, lam + a * dlam
    return z, nu, lam, s


def _vertex(H, g, Aeq, beq, Ain, bin_, active):
    """Equality-constrained QP with the rows `active` of Ain held at their bound (minimum-norm multipliers)."""
    n = H.shape[0]
    A = np.vstack((Aeq, Ain[active]))
    b = np.concatenate((beq, bin_[active]))
    K = np.zeros((n + A.shape[0], n + A.shape[0]))
    K[:n, :n] = H
    K[:n, n:] = A.T
    K[n:, :n] = A
    rhs = np.concatenate((-g, b))
    sol = np.linalg.lstsq(K, rhs, rcond=None)[0]
    z = sol[:n]
    # one step of iterative refinement
    sol = sol + np.linalg.lstsq(K, rhs - K @ sol, rcond=None)[0]
    z = sol[:n]
    nu = sol[n:n + Aeq.shape[0]]
    lam = np.zeros(Ain.shape[0])
    lam[active] = sol[n + Aeq.shape[0]:]
    return z, nu, lam


def solve_qp(H, g, Aeq, beq, Ain, bin_):
    z, nu, lam, s = _interior_point(H, g, Aeq, beq, Ain, bin_)
    best = (z, _kkt_residuals(H, g, Aeq, beq, Ain, bin_, z, nu, lam), "interior-point")
    active = lam > s
    for _ in range(20):
        zv, nuv, lamv = _vertex(H, g, Aeq, beq, Ain, bin_, active)
        slack = bin_ - Ain @ zv
        drop = active & (lamv < -1e-9)
        add = (~active) & (slack < -1e-10)
        if not drop.any() and not add.any():
            res = _kkt_residuals(H, g, Aeq, beq, Ain, bin_, zv, nuv, np.maximum(lamv, 0.0))
            if max(res["dual"], res["eq"], res["ineq"]) <= 1e-8 * max(1.0, np.abs(g).max()):
                best = (zv, res, "active-set vertex")
            break
        active = (active & ~drop) | add
    z, res, how = best
    res["method"] = how
    res["n_active"] = int(active.sum())
    return z, res
