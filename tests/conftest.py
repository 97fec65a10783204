import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bsplinekit.jl_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def bk():
    """The product package (loads libbsplinekit_b200.so; fails loudly if not built)."""
    import bsplinekit_b200
    return bsplinekit_b200


def dec_range(a, b, n):
    """Julia range a:step:b with decimal steps -> correctly rounded decimals (TwicePrecision)."""
    import numpy as np
    return np.array([float("%.10g" % (a + (b - a) * i / (n - 1))) for i in range(n)])
