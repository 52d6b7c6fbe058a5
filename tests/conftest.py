import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The C oracle (builds oracle/liboracle.so on first use)."""
    from oracle import coracle
    coracle.lib()
    return coracle


@pytest.fixture(scope="session")
def pyref():
    from oracle import pyref
    return pyref


@pytest.fixture(scope="session")
def b200():
    """The product library, initialised on cuda:0.  GPU tests only."""
    import nova_b200
    from nova_b200.native import check, lib
    check(lib().b200_init(0))
    return nova_b200
