import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_ok() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a host without a CUDA device skips the gpu-marked tests instead of erroring."""
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (libskm_b200 has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ctx():
    """The default libskm_b200 context on cuda:0 (GPU tests only)."""
    from sparsifiedkmeans_b200 import default_context
    return default_context(0)
