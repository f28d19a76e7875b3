"""Exact optimum of EVERY problem of a batch, started from a candidate solution (test helper; torch fp64, CPU or CUDA).

The reference QP (/root/reference/src/control/mpc_controller.py:47-117) with each slack minimised out analytically and the
states eliminated through the linearised dynamics is
    min_U  phi(U) = sum_k (x_k - r_k)'Q(x_k - r_k) + u_k'R u_k + sum_rows w * dist(g_row(X(U), U), [lo, hi])^2 ,
a C^1, strictly convex, piecewise-quadratic function of the 2N controls.  On the piece (active set) that contains U the
Newton step U+ = U - H^-1 grad is that piece's exact minimiser; if U+ lies on the same piece it is THE minimiser (strict
convexity), otherwise the step is repeated from U+ (semismooth Newton, finite termination).  Everything is batched, so a
65,536-problem launch is certified problem by problem: the returned U* gives |u0 - u0*| exactly, not a bound.
Inputs are the (A_k, B_k) entries of the linearisation hook, itself checked at 1e-12 against vehicle_model.linearize.
This shares no code with the CUDA path, the oracle's ADMM or its KKT-Newton (which works in (X, U) space with multipliers).
"""
import numpy as np
import torch


def _t(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)


def _dist(g, lo, hi):
    return torch.clamp(g - hi, min=0.0) - torch.clamp(lo - g, min=0.0)


class _Problem:
    def __init__(self, p, x0, refu, u_prev, lin, dev):
        self.p, self.dev = p, dev
        self.x0, self.r, self.up = _t(x0, dev), _t(refu, dev), _t(u_prev, dev)
        self.a02, self.a03, self.a12, self.a13, self.b21 = (_t(a, dev) for a in lin)
        self.B, self.N = self.a02.shape
        Q, R, QN = (np.asarray(m, float) for m in (p.q, p.r, p.q_terminal))
        self.Qs, self.Rs, self.QNs = _t(Q + Q.T, dev), _t(R + R.T, dev), _t(QN + QN.T, dev)
        N, dt, r = self.N, p.dt, self.r
        kl = torch.clamp(torch.arange(N, device=dev) - 1, min=0)
        xl = r[:, kl]                                       # linearisation points ref[max(k-1,0)], ulin = 0
        fx0 = xl[..., 0] + dt * xl[..., 3] * torch.cos(xl[..., 2]); fx1 = xl[..., 1] + dt * xl[..., 3] * torch.sin(xl[..., 2])
        self.c0 = fx0 - (xl[..., 0] + self.a02 * xl[..., 2] + self.a03 * xl[..., 3])
        self.c1 = fx1 - (xl[..., 1] + self.a12 * xl[..., 2] + self.a13 * xl[..., 3])
        D = torch.zeros((2 * N, 2 * N), dtype=torch.float64, device=dev)          # rate rows: u_k - u_{k-1}
        for k in range(N):
            for i in range(2):
                D[2 * k + i, 2 * k + i] = 1.0
                if k:
                    D[2 * k + i, 2 * (k - 1) + i] = -1.0
        self.D = D

    def rollout(self, u, sens):
        """u (B,2N) ordered (k,i) -> X (B,N+1,4) and, if sens, S = dX/du (B,N+1,4,2N)."""
        B, N, dt = self.B, self.N, self.p.dt
        X = torch.zeros((B, N + 1, 4), dtype=torch.float64, device=self.dev); X[:, 0] = self.x0
        S = torch.zeros((B, N + 1, 4, 2 * N), dtype=torch.float64, device=self.dev) if sens else None
        for k in range(N):
            x = X[:, k]
            X[:, k + 1, 0] = x[:, 0] + self.a02[:, k] * x[:, 2] + self.a03[:, k] * x[:, 3] + self.c0[:, k]
            X[:, k + 1, 1] = x[:, 1] + self.a12[:, k] * x[:, 2] + self.a13[:, k] * x[:, 3] + self.c1[:, k]
            X[:, k + 1, 2] = x[:, 2] + self.b21[:, k] * u[:, 2 * k + 1]
            X[:, k + 1, 3] = x[:, 3] + dt * u[:, 2 * k]
            if sens:
                s = S[:, k]
                S[:, k + 1, 0] = s[:, 0] + self.a02[:, k, None] * s[:, 2] + self.a03[:, k, None] * s[:, 3]
                S[:, k + 1, 1] = s[:, 1] + self.a12[:, k, None] * s[:, 2] + self.a13[:, k, None] * s[:, 3]
                S[:, k + 1, 2] = s[:, 2]; S[:, k + 1, 2, 2 * k + 1] += self.b21[:, k]
                S[:, k + 1, 3] = s[:, 3]; S[:, k + 1, 3, 2 * k] += dt
        return X, S

    def pieces(self, u, X):
        """signed distances of every soft row to its interval: v (B,N+1), inputs (B,2N), rates (B,2N)."""
        p, N = self.p, self.N
        dv = _dist(X[:, :, 3], p.v_bounds[0], p.v_bounds[1])
        lo_u = _t(np.tile([p.u_bounds[0][0], p.u_bounds[1][0]], N), self.dev); hi_u = _t(np.tile([p.u_bounds[0][1], p.u_bounds[1][1]], N), self.dev)
        du_ = _dist(u, lo_u, hi_u)
        rate = u @ self.D.T
        off = torch.zeros_like(u); off[:, :2] = self.up                                   # k = 0 rows: u_0 - u_prev
        lo_d = _t(np.tile([p.du_bounds[0][0], p.du_bounds[1][0]], N), self.dev); hi_d = _t(np.tile([p.du_bounds[0][1], p.du_bounds[1][1]], N), self.dev)
        dd = _dist(rate - off, lo_d, hi_d)
        return dv, du_, dd

    def objective(self, u):
        p, N = self.p, self.N
        X, _ = self.rollout(u, False)
        dv, du_, dd = self.pieces(u, X)
        e = X - self.r
        f = 0.5 * torch.einsum("bki,ij,bkj->b", e[:, :N], self.Qs, e[:, :N]) + 0.5 * torch.einsum("bi,ij,bj->b", e[:, N], self.QNs, e[:, N])
        uk = u.reshape(self.B, N, 2)
        f = f + 0.5 * torch.einsum("bki,ij,bkj->b", uk, self.Rs, uk)
        return f + p.slack_velocity * (dv ** 2).sum(1) + p.slack_input * (du_ ** 2).sum(1) + p.slack_rate * (dd ** 2).sum(1)

    def grad_hess(self, u):
        p, B, N = self.p, self.B, self.N
        X, S = self.rollout(u, True)
        dv, du_, dd = self.pieces(u, X)
        e = X - self.r
        gx = torch.einsum("ij,bkj->bki", self.Qs, e)
        gx[:, N] = torch.einsum("ij,bj->bi", self.QNs, e[:, N])
        gx[:, :, 3] += 2.0 * p.slack_velocity * dv
        Sf = S.reshape(B, 4 * (N + 1), 2 * N)
        uk = u.reshape(B, N, 2)
        g = torch.bmm(Sf.transpose(1, 2), gx.reshape(B, -1, 1)).squeeze(2) + torch.einsum("ij,bkj->bki", self.Rs, uk).reshape(B, -1) \
            + 2.0 * p.slack_input * du_ + (2.0 * p.slack_rate * dd) @ self.D
        Sw = torch.einsum("ij,bkjn->bkin", self.Qs, S)
        Sw[:, N] = torch.einsum("ij,bjn->bin", self.QNs, S[:, N])
        Sw[:, :, 3] += (2.0 * p.slack_velocity * (dv != 0))[:, :, None] * S[:, :, 3]
        H = torch.bmm(Sf.transpose(1, 2), Sw.reshape(B, 4 * (N + 1), 2 * N))
        H = H + torch.block_diag(*([self.Rs] * N))[None]
        H = H + torch.diag_embed(2.0 * p.slack_input * (du_ != 0).double())
        H = H + torch.einsum("rn,br,rm->bnm", self.D, 2.0 * p.slack_rate * (dd != 0).double(), self.D)
        return g, H, (dv != 0, du_ != 0, dd != 0)


def exact_optimum(p, x0, ref_unwrapped, u_prev, U, lin, device="cpu", max_iter=40, chunk=2048):
    """U (B,2,N): candidate controls.  Returns dict(U (B,2,N) exact minimisers, X (B,4,N+1), settled (B,) bool,
    newton_steps (B,), grad_norm (B,) at the returned point)."""
    B_all = len(U)
    out = dict(U=np.empty_like(U), X=np.empty((B_all, 4, U.shape[2] + 1)), settled=np.zeros(B_all, bool),
               newton_steps=np.zeros(B_all, np.int32), grad_norm=np.zeros(B_all))
    for lo in range(0, B_all, chunk):
        sl = slice(lo, min(lo + chunk, B_all))
        pr = _Problem(p, x0[sl], ref_unwrapped[sl], u_prev[sl], tuple(a[sl] for a in lin), device)
        u = _t(np.transpose(U[sl], (0, 2, 1)).reshape(pr.B, -1), device)
        settled = torch.zeros(pr.B, dtype=torch.bool, device=device)
        steps = torch.zeros(pr.B, dtype=torch.int32, device=device)
        for _ in range(max_iter):
            g, H, act = pr.grad_hess(u)
            step = -torch.linalg.solve(H, g.unsqueeze(2)).squeeze(2)
            # damped (Armijo) semismooth Newton: a full step whenever it does not increase phi (always, near the optimum)
            f0, slope = pr.objective(u), (g * step).sum(1)
            t = torch.ones(pr.B, dtype=torch.float64, device=device)
            for _h in range(30):
                bad = pr.objective(u + t[:, None] * step) > f0 + 1e-4 * t * slope + 1e-12 * f0.abs()
                if not bool(bad.any()):
                    break
                t = torch.where(bad, 0.5 * t, t)
            un = u + t[:, None] * step
            Xn, _ = pr.rollout(un, False)
            actn = tuple(d != 0 for d in pr.pieces(un, Xn))
            same = torch.ones(pr.B, dtype=torch.bool, device=device)
            for a, b in zip(act, actn):
                same &= (a == b).all(dim=1)
            move = ~settled
            u = torch.where(move[:, None], un, u)
            steps += move.int()
            settled |= same & (t == 1.0)
            if bool(settled.all()):
                break
        g, _, _ = pr.grad_hess(u)
        X, _ = pr.rollout(u, False)
        out["U"][sl] = u.reshape(pr.B, -1, 2).permute(0, 2, 1).cpu().numpy()
        out["X"][sl] = X.permute(0, 2, 1).cpu().numpy()
        out["settled"][sl] = settled.cpu().numpy(); out["newton_steps"][sl] = steps.cpu().numpy()
        out["grad_norm"][sl] = g.norm(dim=1).cpu().numpy()
    return out


def lin_from_hook(A, Bm):
    """(a02, a03, a12, a13, b21) from the linearisation hook's A (B,N,4,4), B (B,N,4,2)."""
    return A[:, :, 0, 2], A[:, :, 0, 3], A[:, :, 1, 2], A[:, :, 1, 3], Bm[:, :, 2, 1]
