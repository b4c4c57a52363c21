import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle backend (test infrastructure; built on demand from oracle/)."""
    from oracle_backend import OracleBackend
    return OracleBackend()


@pytest.fixture(scope="session")
def gpu():
    """The product backend: libfacgpu.so through the C ABI.  No fallback of any kind."""
    if not _has_gpu():
        pytest.skip("no CUDA device")
    from fac_b200 import GpuBackend
    return GpuBackend()


def pytest_generate_tests(metafunc):
    # `backend` parametrizes a test over the oracle (CPU, always) and the GPU product path
    # (marked gpu, so `-m "not gpu"` pins the oracle and `-m gpu` is the parity run).
    if "backend" in metafunc.fixturenames:
        metafunc.parametrize("backend", ["oracle", pytest.param("gpu", marks=pytest.mark.gpu)], indirect=True)


@pytest.fixture
def backend(request):
    return request.getfixturevalue(request.param)
