import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI tests need libscldm_b200.so; build it in-tree when it is missing or older than its sources (nvcc cross-compiles
    for sm_100a without a GPU, ~20 s; a no-op when the library is fresh, e.g. on the GPU box where the built .so travels)."""
    from scldm_b200 import build

    try:
        build.build()
    except Exception:
        if not os.path.isfile(build.LIB):   # a stale-looking but present library (copied snapshot, no compiler) is still usable
            raise
