import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build libhikari_cuda.so (cross-compiles without a GPU), the oracle and the spectrum table once."""
    import __graft_entry__ as g
    g.build()
    yield


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
