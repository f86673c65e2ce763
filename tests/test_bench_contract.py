"""bench.py prints ONE JSON line with the keys the measurement contract names.  CPU: the reference
arm (``--impl reference``: the oracle's lean path on the host cores, bounded sample) on the small 16^3 workload, and its
behaviour on a non-zero rank.  GPU: our arm on the same workload (one step, no CPU baseline)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, check_subprocess

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _run(args, env=None, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, **(env or {})))
    check_subprocess(r)
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_line():
    lines = _run(["--impl", "reference", "--workload", "cfg1b", "--steps", "1", "--warmup", "0"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("voxels/sec joint-inversion") and d["unit"] == "voxels/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["config"]["workload"] and d["config"]["voxels"] == 4096
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"] and cb["unit"] == "voxels/s"
    assert d["e2e"] == {"value": d["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # 16^3 fits the per-step budget: the sample is the whole inversion
    assert cb["full_inversion"] is True and cb["sample"].startswith("full")
    # both arms print the same config dict (the driver compares them): it is built by one function from the workload alone
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(bench.WORKLOADS["cfg1b"], "cfg1b")


def test_reference_arm_uses_all_host_cores_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use every host core for BLAS and say how many."""
    lines = _run(["--impl", "reference", "--workload", "cfg1b", "--steps", "1", "--warmup", "0"],
                 env={"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    d = json.loads(lines[0])
    try:
        ncores = len(os.sched_getaffinity(0))
    except Exception:
        ncores = os.cpu_count()
    assert d["cpu_baseline"]["cores"] == ncores


def test_reference_arm_is_silent_on_other_ranks():
    assert _run(["--impl", "reference", "--gpus", "2", "--workload", "cfg1b", "--steps", "1", "--warmup", "0"],
                env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


@pytest.mark.gpu
def test_our_arm_line():
    lines = _run(["--workload", "cfg1b", "--steps", "2", "--warmup", "3", "--e2e-steps", "1", "--no-cpu-baseline"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3 and d["value"] > 0 and d["scaling"] in ("weak", "strong")
    assert d["gpu_launches"] > 0 and d["finite"] and d["info"] == 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["achieved"] > 0 and {"peak", "unit", "frac", "traffic", "kernel"} <= set(r)
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["stage_ms"]["project"] > 0 and d["stage_ms"]["chol"] > 0 and d["stage_ms"]["total"] <= d["ms_per_step"] * 1.001


@pytest.mark.gpu
@pytest.mark.parametrize("workload,structure", [("cfg1b", "kron"), ("cfg1", "compact"), ("cfg1", "fft")])
def test_gpu_structured_arm_line(workload, structure):
    """The opt-in structured projections through bench.py: same contract, the structure named in config, an HBM roofline."""
    lines = _run(["--workload", workload, "--structure", structure, "--precision", "fp64", "--steps", "2", "--warmup", "3", "--e2e-steps", "1",
                  "--no-cpu-baseline"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl_config"]["structure"].startswith(structure) and d["value"] > 0 and d["finite"] and d["info"] == 0
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["achieved"] > 0 and d["gpu_launches"] > 0
