import dataclasses
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not errored) where no CUDA device is visible."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Compile the native artefacts once (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def oracle_params(N, du_delta=0.15, **kw):
    from oracle import mpc_numpy as O
    return dataclasses.replace(O.Params(horizon=N), du_bounds=((-12.0, 12.0), (-du_delta, du_delta)), **kw)


def product_params(N, du_delta=0.15, map_resolution=0.8):
    from rrt_mpc_b200 import MPCConfig
    p = MPCConfig(horizon=N).to_parameters(map_resolution)
    return dataclasses.replace(p, du_bounds=((-12.0, 12.0), (-du_delta, du_delta)))
