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


# Which GEMM implementation a GPU test runs on (rr_set_gemm_mode): 0 = exact-fp32 SIMT kernels, 1 = the product default and the
# benchmarked path (tcgen05: forward 3 x TF32, backward 3 x bf16).
#   * a test that takes ``gemm_mode`` runs TWICE, once per mode, with the per-mode tolerances of tests/helpers.py (every model-, RankNet-,
#     composite-loss- and optimiser-step-level parity test does);
#   * any other GPU test runs in the PRODUCT DEFAULT (mode 1), except the kernel-level tests of test_gpu_kernels.py, which check the SIMT
#     kernels at 3e-6 and ask for the tensor-core kernels explicitly with ``tc_mode``.
@pytest.fixture(autouse=True)
def _gemm_mode_for_gpu_tests(request):
    if "gpu" not in request.keywords:
        yield
        return
    from reactranker_b200 import _lib
    L = _lib.lib()
    if "gemm_mode" in request.fixturenames:
        mode = None                               # the gemm_mode fixture sets it
    elif "tc_mode" in request.fixturenames:
        mode = 1
    else:
        mode = 0 if request.module.__name__.endswith("test_gpu_kernels") else 1
    if mode is not None:
        L.rr_set_gemm_mode(mode)
    L.rr_set_backward_bf16(1)
    yield
    L.rr_set_gemm_mode(1)
    L.rr_set_backward_bf16(1)
    L.rr_reload_switches()                        # whatever RR_* switch a test monkeypatched is gone again


@pytest.fixture(params=[0, 1], ids=["simt", "tc"])
def gemm_mode(request):
    from reactranker_b200 import _lib
    _lib.lib().rr_set_gemm_mode(request.param)
    yield request.param
    _lib.lib().rr_set_gemm_mode(1)


@pytest.fixture
def tc_mode():
    yield
