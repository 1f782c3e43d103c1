import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The suite's small fixtures must keep exercising the streaming kernels: the automatic one-kernel path for small dense NIPALS
# fits (mbpls_b200/smallfit.py) is switched off here and tested explicitly (tests/test_gpu_small.py, small_path=True).
os.environ.setdefault("MBPLS_SMALL_PATH", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure the in-tree CUDA library exists (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    yield
