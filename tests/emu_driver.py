"""ctypes driver of tests/emu/libemu.so: the kernel's per-problem driver (mpc_solve.h) run lane by lane on the
host.  Test infrastructure only."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(os.path.join(_HERE, "emu", "libemu.so"))
    return _lib


def _P(a):
    return a.ctypes.data_as(dp) if a is not None else None


def pack(p, **s):
    sym = lambda m: (np.asarray(m, float) + np.asarray(m, float).T).ravel()
    par = np.array([p.wheelbase_px, p.dt, *sym(p.q), *sym(p.r), *sym(p.q_terminal),
                    p.u_bounds[0][0], p.u_bounds[1][0], p.u_bounds[0][1], p.u_bounds[1][1], p.v_bounds[0], p.v_bounds[1],
                    p.du_bounds[0][0], p.du_bounds[1][0], p.du_bounds[0][1], p.du_bounds[1][1],
                    p.slack_velocity, p.slack_input, p.slack_rate], float)
    st = np.array([s.get("eps_abs", 1e-3), s.get("eps_rel", 1e-3), s.get("rho", 0.1), s.get("alpha", 1.6), s.get("sigma", 1e-6),
                   5.0, 1e3, 1e-6, 1e6, s.get("delta", 1e-6), s.get("max_iter", 60000), s.get("check_termination", 25),
                   s.get("adaptive_rho", 1), s.get("adaptive_rho_interval", 50), s.get("polish_passes", 1), s.get("polish_refine_iter", 3),
                   s.get("warm_start", 0), s.get("polish_retry", 0), s.get("early_polish", 0), s.get("early_polish_start", 50)], float)
    return par, st


def solve(p, x0, ref, up, reverse=0, warm=None, **s):
    N = p.horizon
    x0 = np.ascontiguousarray(x0, float).reshape(-1, 4); B = len(x0)
    ref = np.ascontiguousarray(ref, float).reshape(B, N + 1, 4)
    up = None if up is None else np.ascontiguousarray(up, float).reshape(B, 2)
    par, st = pack(p, **s)
    out = dict(u0=np.zeros((B, 2)), Xp=np.zeros((B, 4, N + 1)), Up=np.zeros((B, 2, N)), status=np.zeros(B, np.int32),
               iters=np.zeros(B, np.int32), pri=np.zeros(B), dua=np.zeros(B), info=np.zeros((B, 4), np.int32))
    lib().emu_solve_batch(_P(par), _P(st), N, B, reverse, _P(x0), _P(ref), _P(up), _P(warm), _P(out["u0"]), _P(out["Xp"]),
                          _P(out["Up"]), out["status"].ctypes.data_as(ip), out["iters"].ctypes.data_as(ip), _P(out["pri"]),
                          _P(out["dua"]), out["info"].ctypes.data_as(ip))
    return out


def set_form(form):
    """-1: what the library picks (short form of the phases for N+1 <= 32), 0: general (parity) form, 1: short form."""
    lib().emu_set_form(int(form))


def set_fsave(on):
    """1: the ADMM factor is kept in a side buffer while a polish uses its place; 0: it is recomputed when ADMM resumes."""
    lib().emu_set_fsave(int(on))


def warm_size(N):
    return lib().emu_warm_size(N)


def linearize(p, ref):
    N = p.horizon
    ref = np.ascontiguousarray(ref, float); B = ref.shape[0]
    par, st = pack(p)
    A, Bm, c = np.zeros((B, N, 4, 4)), np.zeros((B, N, 4, 2)), np.zeros((B, N, 4))
    lib().emu_linearize_batch(_P(par), _P(st), N, B, _P(ref), _P(A), _P(Bm), _P(c))
    return A, Bm, c
