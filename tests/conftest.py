import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` on a box without a CUDA device: skip instead of erroring inside busca_create."""
    if os.path.exists("/dev/nvidia0") or os.environ.get("BUSCA_FORCE_GPU_TESTS"):
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
