import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _autocast_state_does_not_leak():
    """Some CPU tests answer `torch.is_autocast_enabled` with a monkeypatched True (CUDA autocast cannot be switched on without a
    device).  torch.autocast.__exit__ restores the state it READ on entry, so such a test leaves the real thread-local flag set;
    reset it so that the next test starts from the documented default."""
    yield
    import torch
    for dev in ("cuda", "cpu"):
        try:
            torch.set_autocast_enabled(dev, False)
        except Exception:
            pass
