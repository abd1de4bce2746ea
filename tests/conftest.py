import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def port():
    """our CPU restatement of the reference (oracle/dab_oracle.c)"""
    from oracle import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    """the unmodified reference compiled into oracle/_ref (build container, or prebuilt .so)"""
    from oracle import oracle
    r = oracle.ref()
    if r is None:
        pytest.skip("oracle/_ref not available (no /root/reference and no prebuilt libdabref.so)")
    return r


@pytest.fixture(scope="session")
def dab():
    """the ctypes binding of libdabgpu.so"""
    from dabtools_b200 import lib
    lib.load()
    return lib


@pytest.fixture(scope="session")
def gpu(dab):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; there is no CPU fallback")
    torch.cuda.init()
    from dabtools_b200 import lib
    lib.check(lib.load().dabgpu_set_device(0))
    return dab


GOLDEN = os.path.join(ROOT, "tests", "golden")
