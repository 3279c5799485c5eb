import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count():
    """Devices the CUDA driver sees (0 without a driver): asked of libcuda itself, not of /dev names (those differ between boxes)."""
    import ctypes
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        return n.value if cu.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """`-m gpu` on a box without a CUDA device: skip instead of erroring inside busca_create."""
    if not any("gpu" in it.keywords for it in items) or os.environ.get("BUSCA_FORCE_GPU_TESTS") or _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def weights():
    from busca_b200 import synth
    return synth.make_weights(0)
