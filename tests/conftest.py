import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import lt_oracle
    lt_oracle.build()
    return lt_oracle


@pytest.fixture(scope="session")
def ctx():
    """The product library on cuda:0.  No fallback: a missing .so or device is a hard error."""
    from bonej2_b200 import _native
    c = _native.Context(0)
    yield c
    c.close()
