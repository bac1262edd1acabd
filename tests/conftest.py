import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: headline-size cases whose CPU oracle takes tens of seconds (still part of -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Builds libaces_b200.so and the oracle if they are stale (nvcc cross-compiles on CPU)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    import aces_b200

    c = aces_b200.Context(0)  # raises without a GPU: there is no CPU fallback
    yield c
    c.close()
