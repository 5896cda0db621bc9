import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    """The reference's own C files compiled by oracle/Makefile; None if oracle/_ref was never built."""
    from oracle import oracle
    return oracle.ref()


@pytest.fixture(scope="session")
def checker():
    from oracle import oracle
    return oracle.best()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
