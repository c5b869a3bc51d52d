import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly on a box with a GPU but no built extension; on a CPU-only
    container they are skipped unless explicitly selected with -m gpu."""
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _deterministic_inputs():
    """Every test starts from the same torch seed (CPU and CUDA generators): a failure must be reproducible."""
    import torch

    torch.manual_seed(1234)
    yield
