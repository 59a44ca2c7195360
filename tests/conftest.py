import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
        return cache[name]

    return load


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the C-ABI library; GPU tests must go through it."""
    import torch
    from ideal_ballooning_solver_b200 import _lib
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return _lib.load(build_if_missing=True)
