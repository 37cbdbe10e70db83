import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: full BASELINE.json sizes (tens of seconds each on the GPU box)")


@pytest.fixture(scope="session")
def box2():
    return np.load(os.path.join(GOLDEN, "box2_tris.npy"))


@pytest.fixture(scope="session")
def box():
    return np.load(os.path.join(GOLDEN, "box_tris.npy"))


@pytest.fixture(scope="session")
def bunny():
    return np.load(os.path.join(GOLDEN, "bunny_tris.npz"))["tris"]


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def bs():
    """The product package with a live context; fails loudly when the CUDA library or the GPU is missing."""
    import baby_shark_b200 as B
    B.load_library()
    B.Context.default()
    return B
