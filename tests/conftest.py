import hashlib
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def nsb():
    """The product package (ctypes binding of libnsb.so).  Built on demand; never falls back to the oracle."""
    mod = importlib.import_module("nice-slam-cpp_b200")
    if not os.path.exists(mod.lib_path()):
        build = importlib.import_module("nice-slam-cpp_b200.build")
        build.build()
    return mod


@pytest.fixture(scope="session")
def syn():
    return importlib.import_module("nice-slam-cpp_b200.synthetic")


@pytest.fixture(scope="session")
def model_inputs(syn):
    grids = syn.make_grids(0)
    decs = syn.make_decoders(0, bias_scale=0.05)
    h = hashlib.sha256()
    for k in syn.LEVELS:
        h.update(np.ascontiguousarray(grids[k]).tobytes()); h.update(np.ascontiguousarray(decs[k]).tobytes())
    return grids, decs, h.hexdigest()


@pytest.fixture(scope="session")
def frames(syn):
    return syn.make_frames(5, 0)


def load_golden(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
