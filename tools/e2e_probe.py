"""Where does Inversion.cubing() spend host wall time?  usage: python tools/e2e_probe.py cfg3"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from geobo_b200 import _lib, config_loader, dist, inversion, synth

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
xN, yN, zN = wl["shape"]
cfg = synth.settings(xN, yN, zN, kernelfunc=wl["kernel"], precision="int8x5")
config_loader.load_settings(cfg, make_outpath=False)
ctx = _lib.default_context()
rank, world = dist.init_from_env(ctx)
f = synth.make_inputs(nd=wl["nd"], seed=0, ctx=ctx)
inv = inversion.Inversion()
inv.create_cubegeometry()
gl0 = inv.gp_length * np.asarray(wl.get("gl_mult", (1.0, 1.0, 1.0)))
orig_build, orig_predict = inv._build_problem, inv.predict3
acc = {}
def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return r
    return w
inv._build_problem = timed("build_problem", orig_build)
inv.predict3 = timed("predict3", orig_predict)
dist.allgather_columns = timed("allgather", dist.allgather_columns)
_lib.Problem.predict = timed("problem.predict", _lib.Problem.predict)
for it in range(4):
    acc.clear()
    inv.gp_length = gl0.copy()
    t = time.perf_counter()
    inv.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    tot = time.perf_counter() - t
    tm = inv.timings
    print("rank %d iter %d: cubing %.1f ms | build_problem %.1f  predict3 %.1f (problem.predict %.1f, allgather %.1f) | device total %.1f (project %.1f)"
          % (rank, it, tot * 1e3, acc["build_problem"] * 1e3, acc["predict3"] * 1e3, acc.get("problem.predict", 0) * 1e3, acc.get("allgather", 0) * 1e3,
             tm["total"], tm["project"]), flush=True)
