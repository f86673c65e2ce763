"""Time the GEMM-engine variants (GEOBO_B200_GEMM="BK,STAGES,LATE") on one workload; one subprocess per variant."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, numpy as np
sys.path.insert(0, %r)
from geobo_b200 import _lib, config_loader, synth, inversion
import bench
wl = bench.WORKLOADS[%r]
xN, yN, zN = wl["shape"]
cfg = synth.settings(xN, yN, zN, kernelfunc=wl["kernel"]); config_loader.load_settings(cfg, make_outpath=False)
ctx = _lib.default_context()
f = synth.make_inputs(nd=wl["nd"], seed=0, ctx=ctx)
inv = inversion.Inversion(); inv.create_cubegeometry()
out = inv.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
h = inv._hyper()
ts = []
for i in range(4):
    inv._problem.predict(h, want_host=False); ts.append(inv._problem.timings())
t = ts[-1]
print(json.dumps(dict(project=min(x["project"] for x in ts[1:]), trsm=t["trsm"], chol=t["chol"], aka=t["aka"], total=min(x["total"] for x in ts[1:]),
                      checksum=float(np.nansum(out[0]) + np.nansum(out[3])), logl=inv.logl)))
'''

def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    variants = sys.argv[2:] or ["16,3,0", "16,4,0", "16,3,1", "16,4,1", "32,3,0", "32,3,1", "32,2,0"]
    for v in variants:
        env = dict(os.environ, GEOBO_B200_GEMM=v)
        r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, wl)], env=env, capture_output=True, text=True, timeout=900)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
        print(v, line, flush=True)

if __name__ == "__main__":
    main()
