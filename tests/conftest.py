import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name))


CUBES = ["density_rec", "magsus_rec", "drill_rec", "density_var", "magsus_var", "drill_var"]


def normwise_err(a, ref):
    """max|a-ref| / max|ref| per cube -- the norm BASELINE.md section 4 states the 1e-5 target in."""
    import numpy as np
    a, ref = np.asarray(a), np.asarray(ref)
    if np.isnan(ref).all():
        return 0.0 if np.isnan(a).all() else float("inf")
    return float(np.abs(a - ref).max() / np.abs(ref).max())
