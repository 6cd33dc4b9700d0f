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
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def K():
    """The product package (ctypes over libcck_b200.so)."""
    import cck_b200
    return cck_b200


@pytest.fixture(scope="session")
def W():
    import cck_b200  # noqa: F401
    from cck_b200 import workloads
    return workloads


@pytest.fixture()
def ctx(K):
    c = K.Context(0)
    yield c
    c.close()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
