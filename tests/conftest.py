import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN_DIR, "reference_golden.npz"))


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, built if necessary (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    from sspslam_b200 import cabi
    return cabi.load()


def has_reference():
    return os.path.isfile("/root/reference/sspslam/sspspace.py")
