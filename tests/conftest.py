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
    from oracle import lvk_oracle
    lvk_oracle.build_native()
    return lvk_oracle


@pytest.fixture(scope="session")
def gpu_stream():
    """One lvkb200 stream on cuda:0.  Fails loudly (no fallback) if the extension or the device is missing."""
    import livevisionkit_b200 as L
    assert L.device_count() > 0, "no CUDA device visible — gpu tests must run on the GPU box"
    s = L.Stream(L.StabilizationFilterSettings.obs_homography_preset(), device=0)
    yield s
    s.close()


@pytest.fixture
def exact_build():
    """Runs a test with the EXACT arithmetic build of the EASU kernels (bit-identical to oracle/easu_ref.c); the
    default is the contract build (include/lvkb200.h: lvkb200_set_remap_exact)."""
    import livevisionkit_b200 as L
    L.set_remap_exact(True)
    yield
    L.set_remap_exact(False)
