import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ci_workflow: marker of the reference's own tests (loaded by path in test_runner_dropin.py)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    oracle.build()
    return oracle.lib()


@pytest.fixture(scope="session", autouse=True)
def _engine_library():
    """The C-ABI library is built in-tree (it travels with the repository snapshot); a checkout without it builds it
    once per session. Nothing is skipped when the build fails: the product path has no fallback."""
    from limap_b200 import _build
    if not os.path.exists(_build.LIB):
        _build.build_native()
    yield
