import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def cuda_device():
    """The -m gpu tests must run the CUDA path; fail loudly otherwise."""
    import ctypes
    try:
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        cudart = None
    import torch
    assert torch.cuda.is_available(), "gpu-marked test without a CUDA device"
    return 0
