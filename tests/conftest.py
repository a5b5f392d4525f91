import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def b2d():
    """The product package, bound to cuda:0.  GPU tests only: fails loudly when the library or the GPU is missing."""
    import b2d_loader
    m = b2d_loader.load()
    m.init(0)
    return m


@pytest.fixture(scope="session")
def b2d_nogpu():
    import b2d_loader
    return b2d_loader.load()
