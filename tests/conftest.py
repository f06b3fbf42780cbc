import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "fast_numerics: GPU test that runs the tolerance-parity kernels (the library default)")


@pytest.fixture(scope="session")
def capi():
    """The C-ABI library; built in-tree if it is missing or stale (nvcc cross-compiles without a GPU)."""
    from mpcgpu_b200 import build, _capi
    build.build_lib()
    return _capi


@pytest.fixture(autouse=True)
def _numerics_for_module(request):
    """The library defaults to the tolerance-parity (fast) kernels.  Every GPU test module written against bit-exactness
    runs with GBD_PCG_NUMERICS_BITEXACT; tests/test_gpu_fast.py (and tests marked `fast_numerics`) run with the default."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    from mpcgpu_b200 import build, _capi
    build.build_lib()
    fast = request.node.get_closest_marker("fast_numerics") is not None
    prev = _capi.set_numerics(_capi.NUMERICS_FAST if fast else _capi.NUMERICS_BITEXACT)
    try:
        yield
    finally:
        _capi.set_numerics(prev)


@pytest.fixture(scope="session")
def oracle_pcg():
    from oracle import pcg
    pcg.build()
    return pcg


# GBD-PCG demo system (GBD-PCG/examples/pcg_solve.cu:14-25): n=2, N=3, stored [L|D|R] column-major tiles.
G1_S = [0, 0, 0, 0, -.999, 0, 0, -.999, .999, .0999, -.98, .999,
        .999, -.98, .0999, .999, -2.008, .8801, .8801, -3.0584, .999, .0999, -.98, .999,
        .999, -.98, .0999, .999, -1.019, .8801, .8801, -2.0694, 0, 0, 0, 0]
G1_GAMMA = [3.1385, 0, 0, 3.0788, .0031, 3.0788]
# fp64 direct solution of the demo system
G1_LAMBDA = [-303.70298609, -46.41593968, -315.17630263, -14.89830942, -298.79086192, 13.50378269]


@pytest.fixture(scope="session")
def g1():
    import numpy as np
    from mpcgpu_b200 import synth
    S = np.array(G1_S, np.float32)
    return dict(n=2, N=3, S=S, Pinv=synth.stair_preconditioner(S, 2, 3).astype(np.float32),
                gamma=np.array(G1_GAMMA, np.float32), lam=np.array(G1_LAMBDA))
