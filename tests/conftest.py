import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def small_test_params():
    from tiger_hlm_gpu_b200.hostio import load_spatial_params
    return load_spatial_params(os.path.join(GOLDEN, "small_test.csv"))


@pytest.fixture(scope="session")
def golden204():
    return np.load(os.path.join(GOLDEN, "model204_example.npz"))


@pytest.fixture(scope="session")
def golden_dummy():
    return np.load(os.path.join(GOLDEN, "dummy_example.npz"))


@pytest.fixture(scope="session")
def solver():
    """One device context for the GPU tests; raises (never skips) when the CUDA path is unusable."""
    from tiger_hlm_gpu_b200 import Solver
    s = Solver(0)
    yield s
    s.close()
