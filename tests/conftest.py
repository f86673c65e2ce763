import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# compile command of the host builds of the device sources (tests/host_harness/*.cpp).  GEOBO_B200_HARNESS_CXXFLAGS adds flags:
# tests/test_host_harness_asan.py re-runs those tests with -fsanitize=address so that an out-of-bounds index in the per-thread
# arithmetic of a kernel is caught on the CPU (the harness buffers have exactly the device sizes).
HARNESS_CXX = ["g++", "-O2", "-ffp-contract=off"] + os.environ.get("GEOBO_B200_HARNESS_CXXFLAGS", "").split() + ["-shared", "-fPIC"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


# GPU tests written after the round's GPU minutes were spent have not run on hardware yet.  They are collected
# AFTER the hardware-verified parity tests, so that under `-x` a surprise in new code cannot hide the parity result
# of the path itself.  Drop a name from this list once a `gpurun` log under profiles/ shows it green.
NOT_YET_RUN_ON_HARDWARE = (
    # ordered by what a failure under -x would hide: first the full-size parity of the (hardware-verified) main path, then the
    # small additions around it, last the opt-in structured projections
    "test_fullsize_cubing_vs_cpu_oracle",
    "test_two_level_cholesky_flag_vs_oracle",
    "test_our_arm_line",
    "test_gpu_align_drill_",
    "test_gpu_create_synsurvey_",
    "test_gpu_optimize_gp_",
    "test_gpu_proposal_drivers_vs_oracle",
    "test_cholesky_lookahead_flag_vs_oracle",
    "test_gpu_kron_",
    "test_gpu_compact_",
    "test_gpu_fft_",
    "test_gpu_structured_arm_line",
)


def pytest_collection_modifyitems(config, items):
    def rank(item):
        for i, prefix in enumerate(NOT_YET_RUN_ON_HARDWARE):
            if item.name.startswith(prefix):
                return 1 + i
        return 0
    items.sort(key=rank)        # stable: file order is kept inside each group


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
