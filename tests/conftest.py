import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """Make sure the in-tree library exists (cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from densereg_b200 import _ffi
    return _ffi.load()
