import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def rbslam_lib():
    """The product library.  GPU tests must run the CUDA path: fail loudly otherwise."""
    import rbslam
    from rbslam import _capi
    L = _capi.lib()
    assert L.rbslam_device_count() > 0, "no CUDA device visible: GPU tests cannot run"
    return rbslam


def assert_close_norm(a, b, tol=1e-8, what=""):
    """max|a-b| <= tol * max(|b|_max, tiny)  (norm-wise, as SURVEY 8c prescribes for P)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    err = np.max(np.abs(a - b)) if a.size else 0.0
    assert err <= tol * scale, "%s: max abs err %.3e > %.1e * %.3e" % (what, err, tol, scale)
