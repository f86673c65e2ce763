"""The @gpu tests' CPU side (fixtures, oracle-side arrange code, fixture files, shapes) must run on a box WITHOUT a device up to the
first C-ABI call: `pytest -m gpu --gpu-dry-run` (tests/conftest.py).  Round 1 ended with a red GPU run because an oracle-side
reshape inside a GPU test raised before any CUDA call and `-x` hid 45 tests behind it; this CPU test makes that class of failure
show up here."""
import os
import subprocess
import sys

from conftest import ROOT


def test_every_gpu_test_reaches_its_first_device_call_on_a_cpu_box():
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "--gpu-dry-run", "-q", "-p", "no:cacheprovider"],
                       capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
