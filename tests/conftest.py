import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/efg_oracle.py): the checker, built on demand with gcc."""
    import efg_oracle
    efg_oracle.build()
    return efg_oracle


@pytest.fixture(scope="session")
def fg():
    import flexigal_b200
    return flexigal_b200
