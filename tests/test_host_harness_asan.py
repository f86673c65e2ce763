"""The host builds of the device sources (tests/host_harness/*.cpp: covariance formulas, A_sens, digit extraction, align_drill,
the Kronecker / tap-sum / FFT projections) re-run under AddressSanitizer: every buffer of a harness has exactly the size its
device counterpart has (scratch lattices, "shared memory", factor lines, tables, rows without padding), so an index of the
per-thread arithmetic that leaves its buffer -- the kind of bug that corrupts memory silently on the GPU -- aborts here."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

HOST_TESTS = ["tests/test_kron.py", "tests/test_compact.py", "tests/test_fftconv.py", "tests/test_align_drill.py", "tests/test_formulas_host.py",
              "tests/test_digit_slices.py"]


def test_host_harness_tests_pass_under_address_sanitizer():
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan.so not available")
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1",
               GEOBO_B200_HARNESS_CXXFLAGS="-fsanitize=address -fno-omit-frame-pointer -g")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider",
                        "-k", "not test_oracle and not test_structured_oracle"] + HOST_TESTS,      # NumPy-only tests have no harness to check cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1200)
    tail = r.stdout[-3000:] + r.stderr[-3000:]
    assert r.returncode == 0 and "AddressSanitizer" not in tail, tail
    assert " passed" in r.stdout
