import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import fft_lines_b200 as F
        return F.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/): test infrastructure only."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def lib():
    """The product library; built in-tree if it is missing."""
    from fft_lines_b200 import build as B
    B.build()
    import fft_lines_b200 as F
    return F
