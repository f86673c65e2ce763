"""HBM-write roofline of the dense covariance assembly kernel (kernels.create_cov on the grid spec, K1 of SURVEY.md 8(d)):
9 N^2 fp64 values written per launch; achieved GB/s = 9 N^2 * 8 B / kernel time against MEASURED_PEAKS.json hbm_gbs.
usage: python tools/assembly_bench.py [xN yN zN]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from geobo_b200 import _lib

shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 32, 16)
ctx = _lib.default_context()
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
N = int(np.prod(shape))
vox = (3050.0 / shape[0], 1952.0 / shape[1], 800.0 / shape[2])
res = []
for kf, gl in (("exp", [2 * vox[0], 2.04 * vox[0], 2 * vox[0]]), ("matern32", [2 * vox[0], 2.02 * vox[0], 2.04 * vox[0]]), ("sparse", [2 * vox[0], 2.04 * vox[0], 2 * vox[0]])):
    ts = []
    for it in range(6):
        _, ms = ctx.create_cov_grid(shape, vox, gl, [1.0, 0.2, 0.2], 1.0, kf, want_output=False)
        if it >= 2:
            ts.append(ms)
    ms = float(np.median(ts))
    gbs = 9.0 * N * N * 8 / (ms * 1e-3) / 1e9
    res.append({"kernel": kf, "cube": "%dx%dx%d" % shape, "N": N, "bytes_written": 9 * N * N * 8, "ms": ms, "achieved_gbs": gbs,
                "peak_gbs": peaks.get("hbm_gbs"), "frac": gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None})
print(json.dumps({"what": "create_cov_grid_kernel: dense 3N x 3N covariance assembly, HBM-write bound", "runs": res}, indent=1))
