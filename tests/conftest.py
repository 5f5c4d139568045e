import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    return oracle.lib()


@pytest.fixture(scope="session")
def gpu_ctx():
    """One library context on cuda:0 for the whole GPU test session."""
    import ctypes as C
    from ctsm_b200 import abi
    L = abi.lib()
    prm = abi.default_params()
    ctx = C.c_void_p()
    rc = L.ctsm_b200_init(C.byref(prm), C.byref(ctx))
    assert rc == 0, "ctsm_b200_init failed with %d (no CUDA device? there is no CPU fallback)" % rc
    yield L, ctx, prm
    L.ctsm_b200_finalize(ctx)
