"""ctypes wrapper of oracle/liboracle.so (the C restatement, oracle/mpc_oracle.c).  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "mpc_oracle.c")
_LIB = os.path.join(_HERE, "liboracle.so")


class OParams(C.Structure):
    _fields_ = [("wheelbase_px", C.c_double), ("dt", C.c_double), ("horizon", C.c_int32), ("_pad", C.c_int32),
                ("q", C.c_double * 16), ("r", C.c_double * 4), ("q_terminal", C.c_double * 16),
                ("u_bounds", C.c_double * 4), ("v_bounds", C.c_double * 2), ("du_bounds", C.c_double * 4),
                ("slack_velocity", C.c_double), ("slack_input", C.c_double), ("slack_rate", C.c_double)]


class OSettings(C.Structure):
    _fields_ = [("eps_abs", C.c_double), ("eps_rel", C.c_double), ("rho", C.c_double), ("alpha", C.c_double),
                ("sigma", C.c_double), ("adaptive_rho_tolerance", C.c_double), ("rho_eq_factor", C.c_double),
                ("rho_min", C.c_double), ("rho_max", C.c_double), ("delta", C.c_double),
                ("max_iter", C.c_int32), ("check_termination", C.c_int32), ("adaptive_rho", C.c_int32),
                ("adaptive_rho_interval", C.c_int32), ("polish_passes", C.c_int32), ("polish_refine_iter", C.c_int32),
                ("scaling", C.c_int32), ("z0_projected", C.c_int32)]


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-o", _LIB, _SRC, "-lm"], check=True)
    return _LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.oracle_default_settings.argtypes = [C.POINTER(OSettings)]
        L.oracle_solve_batch.restype = C.c_int
        L.oracle_track.restype = C.c_int
        _lib = L
    return _lib


def make_params(p) -> OParams:
    """p: any object with the MPCParameters fields (oracle.mpc_numpy.Params or rrt_mpc_b200.MPCParameters)."""
    o = OParams()
    o.wheelbase_px, o.dt, o.horizon = float(p.wheelbase_px), float(p.dt), int(p.horizon)
    o.q[:] = np.asarray(p.q, float).reshape(16).tolist()
    o.r[:] = np.asarray(p.r, float).reshape(4).tolist()
    o.q_terminal[:] = np.asarray(p.q_terminal, float).reshape(16).tolist()
    o.u_bounds[:] = [p.u_bounds[0][0], p.u_bounds[0][1], p.u_bounds[1][0], p.u_bounds[1][1]]
    o.v_bounds[:] = [p.v_bounds[0], p.v_bounds[1]]
    o.du_bounds[:] = [p.du_bounds[0][0], p.du_bounds[0][1], p.du_bounds[1][0], p.du_bounds[1][1]]
    o.slack_velocity, o.slack_input, o.slack_rate = float(p.slack_velocity), float(p.slack_input), float(p.slack_rate)
    return o


def make_settings(**kw) -> OSettings:
    s = OSettings()
    lib().oracle_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def solve_batch(p, x0, ref, u_prev=None, **settings):
    """Returns dict(u0, Xp, Up, status, iters, pri, dua, info)."""
    L = lib()
    op, s = make_params(p), make_settings(**settings)
    N = int(p.horizon)
    x0 = np.ascontiguousarray(x0, float).reshape(-1, 4)
    B = x0.shape[0]
    ref = np.ascontiguousarray(ref, float).reshape(B, N + 1, 4)
    up = None if u_prev is None else np.ascontiguousarray(u_prev, float).reshape(B, 2)
    out = dict(u0=np.zeros((B, 2)), Xp=np.zeros((B, 4, N + 1)), Up=np.zeros((B, 2, N)), status=np.zeros(B, np.int32),
               iters=np.zeros(B, np.int32), pri=np.zeros(B), dua=np.zeros(B), info=np.zeros((B, 4), np.int32))
    L.oracle_solve_batch(C.byref(op), C.byref(s), B, _p(x0), _p(ref), _p(up), _p(out["u0"]), _p(out["Xp"]), _p(out["Up"]),
                         _p(out["status"]), _p(out["iters"]), _p(out["pri"]), _p(out["dua"]), _p(out["info"]))
    return out


def qp_dense(p, x0, ref, u_prev=None):
    L = lib()
    op = make_params(p)
    N = int(p.horizon); n, m = 11 * N + 5, 19 * N + 7
    P, q, A, l, u = np.zeros((n, n)), np.zeros(n), np.zeros((m, n)), np.zeros(m), np.zeros(m)
    x0 = np.ascontiguousarray(x0, float); ref = np.ascontiguousarray(ref, float)
    up = None if u_prev is None else np.ascontiguousarray(u_prev, float)
    L.oracle_qp_dense(C.byref(op), _p(x0), _p(ref), _p(up), _p(P), _p(q), _p(A), _p(l), _p(u))
    return P, q, A, l, u


def linearize_window(p, ref):
    L = lib()
    op = make_params(p)
    N = int(p.horizon)
    ref = np.ascontiguousarray(ref, float)
    refu, As, Bs, cs = np.zeros((N + 1, 4)), np.zeros((N, 4, 4)), np.zeros((N, 4, 2)), np.zeros((N, 4))
    L.oracle_linearize_window(C.byref(op), _p(ref), _p(refu), _p(As), _p(Bs), _p(cs))
    return refu, As, Bs, cs


def track(p, ref_global, state0, goal, sim_steps, **settings):
    L = lib()
    op, s = make_params(p), make_settings(**settings)
    rg = np.ascontiguousarray(ref_global, float)
    st0 = np.ascontiguousarray(state0, float); g = np.ascontiguousarray(goal, float)
    states = np.full((sim_steps, 4), np.nan); ctr = np.full((sim_steps, 2), np.nan)
    sst = np.zeros(sim_steps, np.int32); sit = np.zeros(sim_steps, np.int32); fl = C.c_int(0)
    n = L.oracle_track(C.byref(op), C.byref(s), _p(rg), len(rg), _p(st0), _p(g), int(sim_steps), _p(states), _p(ctr), _p(sst), _p(sit), C.byref(fl))
    return dict(states=states, controls=ctr, n_steps=n, flags=fl.value, step_status=sst, step_iters=sit)
