import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def oracle_backend(monkeypatch):
    """Host-logic tests only: stand the C oracle in for the CUDA kernels behind i2pnet_b200._cabi
    so that the nn.Module / autograd plumbing can be checked on a machine without a GPU.  The
    product never does this -- without libi2p_b200.so and a CUDA tensor it raises."""
    from tests import cpu_backend
    cpu_backend.patch(monkeypatch)
    yield


@pytest.fixture(params=[0, 7, 23, 55, 119], ids=["fma", "tcgen05", "tcgen05-sw128", "tcgen05-sw128-maxk", "tcgen05-sw128-maxk-wide"])
def tensor_cores(request):
    """The shared-MLP kernel families: f32 FMA (mask 0), tcgen05 tf32 with the 3-term hi/lo split for the
    forward, dX and dW GEMMs (mask 7), and the same with SWIZZLE_128B operand tiles + bulk-copied weights
    (mask 7 | 16), and with max-over-K gradient sources on the
    tensor cores as well (mask 7 | 16 | 32, the default; include/i2p_b200.h)."""
    from i2pnet_b200 import _cabi
    L = _cabi.lib()
    before = L.i2p_get_mlp_tensor_cores()
    L.i2p_set_mlp_tensor_cores(request.param)
    yield request.param
    L.i2p_set_mlp_tensor_cores(before)
