import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kat():
    return np.load(os.path.join(GOLDEN, "kat_small.npz"))


@pytest.fixture(scope="session")
def golden_c1():
    return np.load(os.path.join(GOLDEN, "c1.npz"))


@pytest.fixture(scope="session")
def golden_c2():
    return np.load(os.path.join(GOLDEN, "c2.npz"))


@pytest.fixture(scope="session")
def ctx():
    """The CUDA context of the product library; gpu tests only.  No skip: a missing device or extension must fail."""
    import chan_vese_b200 as cv
    c = cv.Context(0)
    yield c
    c.close()


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-300))


SMALL = ["rgb", "gray", "thin", "tall", "two"]
