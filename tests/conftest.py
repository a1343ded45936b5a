import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """C oracle (oracle/oracle.c) -- the checker, never the thing under test."""
    from oracle import cref
    cref.build()
    return cref


@pytest.fixture(scope="session")
def b381():
    """The product's C ABI (libb381_cuda.so); builds in-tree if missing."""
    from midnight_bls12_381_cuda_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from midnight_bls12_381_cuda_b200 import build
        build.build(icicle=True)
    return _lib


@pytest.fixture(scope="session")
def cuda(b381):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("test is marked gpu but no CUDA device is visible (there is no CPU fallback)")
    torch.cuda.set_device(0)
    return torch
