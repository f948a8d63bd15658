import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def pcr():
    """The product package; session-scoped default context on cuda:0 (GPU tests only)."""
    import pointclouds_rs_b200 as p

    p.default_context()
    return p


def n_threads():
    return max(1, min(16, os.cpu_count() or 1))
