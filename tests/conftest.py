import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Every test session needs the in-tree shared library (CPU tests only load it)."""
    lib = os.path.join(ROOT, "polus_b200", "libpolus_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    yield
