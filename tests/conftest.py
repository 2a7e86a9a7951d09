import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def rbl():
    """The product package, with the C-ABI library built (nvcc cross-compiles on CPU-only boxes)."""
    import __graft_entry__
    lib = os.path.join(ROOT, "gpu-randomized-block-lanczos_b200", "lib", "librbl_b200.so")
    if not os.path.exists(lib):
        __graft_entry__.build()
    import rbl_b200
    rbl_b200.load_library()
    return rbl_b200


@pytest.fixture(scope="session")
def gpu(rbl):
    if rbl.lib().rbl_device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box (no CPU fallback exists)")
    return rbl
