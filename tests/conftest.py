import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

REFERENCE_PKG = Path("/root/reference/hrl_ws/src/hrl_trainer")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the live reference under /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = REFERENCE_PKG.exists()
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


@pytest.fixture(autouse=True)
def _pin_global_rngs():
    """Every test starts from the same global numpy / torch (CPU and CUDA) generator state: a test that draws without an explicit
    generator sees the same problem instance whatever ran before it."""
    import numpy as np

    np.random.seed(12345)
    try:
        import torch

        torch.manual_seed(12345)
    except Exception:  # pragma: no cover
        pass
    yield
