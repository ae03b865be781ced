import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(autouse=True)
def _exact_fp32_gemms_by_default(request):
    """GPU parity tests run the exact-fp32 SIMT GEMMs unless they ask for the tensor-core path with the
    ``tc_mode`` fixture; the library's own default (tcgen05) is restored afterwards."""
    if "gpu" not in request.keywords:
        yield
        return
    from reactranker_b200 import _lib
    L = _lib.lib()
    L.rr_set_gemm_mode(1 if "tc_mode" in request.fixturenames else 0)
    yield
    L.rr_set_gemm_mode(1)


@pytest.fixture
def tc_mode():
    yield
