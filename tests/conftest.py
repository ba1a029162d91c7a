import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/liboracle.so); built on demand with gcc. Test infrastructure only."""
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def engine():
    """The product: libcdp_b200.so on cuda:0.  No fallback -- a missing extension or GPU fails the test."""
    from curdleproofs_b200 import Engine
    e = Engine(0)
    yield e
    e.close()
