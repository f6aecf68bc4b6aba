"""Test configuration.  `-m "not gpu"` runs on the CPU-only build box (oracle vs golden vectors, host logic, ABI
symbols, gloo multi-process plumbing); `-m gpu` runs the parity tests proper on a B200 through the C ABI."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def las2():
    """The product package with its CUDA library built (nvcc cross-compiles without a GPU)."""
    import svdlibrs_b200
    from svdlibrs_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    _lib.lib()
    return svdlibrs_b200


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
