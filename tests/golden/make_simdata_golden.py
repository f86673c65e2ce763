#!/usr/bin/env python
"""Fixtures for the steps next to the hot path, from the UNMODIFIED reference.

Build-container only (needs ``/root/reference``); one subprocess per job because the reference binds its YAML at import.

    python tests/golden/make_simdata_golden.py

``align_drill.npz``  ``align_drill`` (``geobo/run_geobo.py:132-159``) -- the function definition is taken out of the
                     reference source with ``ast`` (the module is a script that runs the whole pipeline at import) and
                     executed with the globals it reads -- on (a) synthetic holes on a 12x10x8 cube with samples on window
                     edges, NaN and inf values, and (b) the reference's committed ``simdrill_cylinders.csv`` on the
                     example-1 cube (the result is the ``drilldata0`` the example run feeds to ``cubing``).
``simdata.npz``      ``create_syncube`` (``geobo/simcube.py:34-117``, same extraction, file output stubbed) for the three
                     model names on the example-1 geometry, plus the reference's committed simulator outputs
                     ``examples/testdata/synthetic/simcube_*.csv`` (DENSITY, MAGSUS) and ``simsurveydata_*.csv``
                     (GRAVITY, MAGNETIC) = its own golden vectors for the simulator and the forward model.
"""
import ast
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
SYN = os.path.join(REF, "examples", "testdata", "synthetic")


def _extract(path, names):
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(body) == len(names), (path, names)
    return compile(ast.Module(body=body, type_ignores=[]), "reference:" + os.path.basename(path), "exec")


def _cfg_dict(cl):
    keys = ["xmin", "xmax", "ymin", "ymax", "zmax", "zoff", "zLcube", "xNcube", "yNcube", "zNcube",
            "gp_lengthscale", "gp_err", "gp_coeff", "kernelfunc", "optimize_gp", "XMAG", "YMAG", "ZMAG",
            "c_G", "c_SI_TO_MILLIGALS", "c_GCM3_TO_SI", "fcor_grav", "fcor_mag"]
    return {k: getattr(cl, k) for k in keys}


def _reference_align(cl, voxelpos):
    import numpy as np
    shape = (cl.xNcube, cl.yNcube, cl.zNcube)                      # run_geobo.py:399-403
    g = dict(np=np, xvoxsize=cl.xvoxsize, yvoxsize=cl.yvoxsize, zvoxsize=cl.zvoxsize,
             xxx=voxelpos[0].reshape(shape), yyy=voxelpos[1].reshape(shape), zzz=voxelpos[2].reshape(shape))
    exec(_extract(os.path.join(REF, "geobo", "run_geobo.py"), ["align_drill"]), g)
    return g["align_drill"]


def job_align_small():
    import numpy as np
    from oracle import ref_loader
    mods = ref_loader.load(ref_loader.write_yaml(dict(xNcube=12, yNcube=10, zNcube=8)))
    cl = mods["config_loader"]
    vp = mods["inversion"].Inversion().create_cubegeometry()
    align = _reference_align(cl, vp)
    rng = np.random.default_rng(11)
    dx, dy, dz = cl.xvoxsize, cl.yvoxsize, cl.zvoxsize
    pts, vals = [], []
    for _ in range(5):                                             # five holes, samples every 0.3 voxel heights
        x, y = rng.uniform(0, cl.xLcube), rng.uniform(0, cl.yLcube)
        z = -np.arange(0.0, rng.uniform(0.3, 1.0) * cl.zLcube, 0.3 * dz)
        pts.append(np.column_stack([x + 0 * z, y + 0 * z, z]))
        vals.append(rng.uniform(0.5, 3.0, z.size))
    pts.append(np.column_stack([rng.uniform(0, cl.xLcube, 60), rng.uniform(0, cl.yLcube, 60), -rng.uniform(0, cl.zLcube, 60)]))
    vals.append(rng.standard_normal(60))
    c0 = vp[:, 333]                                                # samples exactly on the window edges of voxel 333
    pts.append(np.array([[c0[0] - dx, c0[1], c0[2]], [c0[0] + dx, c0[1], c0[2]], [c0[0], c0[1] - dy, c0[2] + dz], [c0[0], c0[1], c0[2] - dz]]))
    vals.append(np.array([10.0, 20.0, 30.0, 40.0]))
    coord, data = np.vstack(pts), np.hstack(vals)
    data[[3, 40]] = np.nan                                         # skipped by nanmean
    data[70] = np.inf                                              # infinite mean -> voxel stays 0
    with np.errstate(all="ignore"):
        res = align(coord, data)
    out = dict(cfg_small=json.dumps(_cfg_dict(cl)), coord_small=coord, data_small=data, res_small=res)
    np.savez_compressed(os.path.join(HERE, "_align_small.npz"), **out)
    print("align small:", res.shape, int((res != 0).sum()), "non-zero voxels")


def job_align_example():
    import numpy as np
    import pandas as pd
    from oracle import ref_loader
    mods = ref_loader.load(ref_loader.write_yaml({}))
    cl = mods["config_loader"]
    vp = mods["inversion"].Inversion().create_cubegeometry()
    align = _reference_align(cl, vp)
    drill = pd.read_csv(os.path.join(SYN, "simdrill_cylinders.csv"))       # run_geobo.py:101-121
    drill = drill[(drill.x >= cl.xmin) & (drill.x <= cl.xmax) & (drill.y >= cl.ymin) & (drill.y <= cl.ymax)
                  & (drill.z <= cl.zmax) & (drill.z >= cl.zmin)]
    coord = np.vstack([drill["x"].values - cl.xmin, drill["y"].values - cl.ymin, drill["z"].values]).T
    data = drill["DENSITY"].values
    res = align(coord, data)
    np.savez_compressed(os.path.join(HERE, "_align_example.npz"), cfg_example=json.dumps(_cfg_dict(cl)), coord_example=coord,
                        data_example=data, res_example=res)
    print("align example:", res.shape, int((res != 0).sum()), "non-zero voxels of", coord.shape[0], "samples")


def job_simcube():
    import random
    import numpy as np
    import pandas as pd
    from oracle import ref_loader
    mods = ref_loader.load(ref_loader.write_yaml({}))
    cl = mods["config_loader"]
    vp = mods["inversion"].Inversion().create_cubegeometry()

    class _NoFiles:
        @staticmethod
        def create_vtkcube(*a, **k):
            pass

    class _Frame:                                                   # the csv output of create_syncube is not under test
        def __init__(self, *a, **k):
            self.df = pd.DataFrame(*a, **k)

        def __getattr__(self, n):
            return getattr(self.df, n)

    out = dict(cfg=json.dumps(_cfg_dict(cl)))
    tmp = ref_loader.tempfile.mkdtemp(prefix="geobo_sim_") + os.sep
    g = dict(np=np, pd=pd, os=os, random=random, cs=_NoFiles, inpath=tmp, print=lambda *a, **k: None)
    for k in ("xNcube", "yNcube", "zNcube", "xLcube", "yLcube", "zLcube", "xvoxsize", "yvoxsize", "zvoxsize", "gp_coeff"):
        g[k] = getattr(cl, k)
    exec(_extract(os.path.join(REF, "geobo", "simcube.py"), ["create_syncube"]), g)
    for model in ("cylinders", "layers_2", "layers_3"):
        dens, mags = g["create_syncube"](model, vp)
        out["density_" + model], out["magsus_" + model] = dens, mags
        print(model, dens.shape, float(dens.min()), float(dens.max()))
    for model in ("cylinders", "layers_3"):
        cube = pd.read_csv(os.path.join(SYN, "simcube_%s.csv" % model))
        surv = pd.read_csv(os.path.join(SYN, "simsurveydata_%s.csv" % model))
        out["csv_xyz_" + model] = cube[["x", "y", "z"]].values
        out["csv_density_" + model], out["csv_magsus_" + model] = cube["DENSITY"].values, cube["MAGSUS"].values
        out["csv_sensor_xy_" + model] = surv[["X", "Y"]].values
        out["csv_grav_" + model], out["csv_mag_" + model] = surv["GRAVITY"].values, surv["MAGNETIC"].values
    np.savez_compressed(os.path.join(HERE, "simdata.npz"), **out)
    print("simdata.npz written")


def main(argv):
    if len(argv) > 1:
        globals()["job_" + argv[1]]()
        return
    for job in ("align_small", "align_example", "simcube"):
        subprocess.run([sys.executable, os.path.abspath(__file__), job], check=True)
    import numpy as np
    merged = {}
    for part in ("_align_small.npz", "_align_example.npz"):
        p = os.path.join(HERE, part)
        merged.update(np.load(p))
        os.unlink(p)
    np.savez_compressed(os.path.join(HERE, "align_drill.npz"), **merged)
    print("align_drill.npz written")


if __name__ == "__main__":
    main(sys.argv)
