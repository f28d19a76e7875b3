"""CPU: libcudampc.so loads and exports every symbol include/cudampc.h declares; struct layouts agree; calls that
need a device fail loudly (no fallback).  No compute happens here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, product_params


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cudampc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cudampc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rrt_mpc_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 14
    for n in names:
        assert getattr(lib, n) is not None
    assert set(names) == set(_lib.SYMBOLS)
    assert lib.cudampc_version() == 100


def test_struct_layouts_and_defaults():
    from rrt_mpc_b200 import _lib
    lib = _lib.load()
    assert C.sizeof(_lib.Params) == 8 * 2 + 8 + 8 * (16 + 4 + 16 + 4 + 2 + 4 + 3)
    assert C.sizeof(_lib.Settings) == 8 * 10 + 4 * 10
    assert C.sizeof(_lib.RolloutCfg) == 8 + 8 * 5 + 8
    s = _lib.Settings()
    lib.cudampc_default_settings(C.byref(s))
    # the reference's OSQP call, src/control/mpc_controller.py:121-131
    assert (s.eps_abs, s.eps_rel, s.max_iter, s.rho, s.alpha, s.adaptive_rho) == (1e-3, 1e-3, 60000, 0.1, 1.6, 1)
    assert s.polish_passes == 1 and s.sigma == 1e-6 and s.check_termination == 25
    c = _lib.RolloutCfg()
    lib.cudampc_default_rollout_cfg(C.byref(c))
    # control_stage.py:141-150 and :45-56
    assert (c.advance_dist2, c.goal_radius, c.relax_v_scale, c.relax_da, c.relax_ddelta) == (25.0, 8.0, 0.6, 5.0, 0.05)


def test_create_rejects_bad_arguments_without_touching_the_device():
    from rrt_mpc_b200 import _lib
    from rrt_mpc_b200.mpc_controller import params_to_c
    import dataclasses, numpy as np
    lib = _lib.load()
    h = C.c_void_p()
    p = params_to_c(product_params(20))
    assert lib.cudampc_create(C.byref(p), 0, 0, C.byref(h)) == -1                    # max_batch < 1
    assert b"max_batch" in lib.cudampc_last_error(None)
    q = np.diag([4.0, 4.0, 0.6, 0.1]); q[0, 1] = q[1, 0] = 5.0                       # indefinite: cvxpy's quad_form would refuse it too
    p2 = params_to_c(dataclasses.replace(product_params(20), q=q))
    assert lib.cudampc_create(C.byref(p2), 8, 0, C.byref(h)) == -1
    assert b"positive semidefinite" in lib.cudampc_last_error(None)
    p3 = params_to_c(dataclasses.replace(product_params(20), horizon=0))
    assert lib.cudampc_create(C.byref(p3), 8, 0, C.byref(h)) == -1
    assert lib.cudampc_solve_batch(None, 1, None, None, None, None, None, None, None, None, None, None, None, None, None) == -1


def test_no_silent_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from rrt_mpc_b200 import MPCController
    ctl = MPCController(product_params(5))
    with pytest.raises(RuntimeError, match="cudampc_create failed"):
        ctl.solve(np.zeros(4), np.zeros((6, 4)))


def test_missing_library_is_an_import_error(monkeypatch):
    from rrt_mpc_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcudampc.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_product_path_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under rrt_mpc_b200/ may reference it."""
    pkg = os.path.join(ROOT, "rrt_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
                assert "libemu" not in txt, f
