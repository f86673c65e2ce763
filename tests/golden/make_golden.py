#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Build-container only (needs ``/root/reference``).  Each job runs in its own
subprocess because the reference binds its YAML at import time
(``geobo/config_loader.py:20-36``).

    python tests/golden/make_golden.py            # all jobs
    python tests/golden/make_golden.py small      # tiny-shape fixtures
    python tests/golden/make_golden.py example 1  # committed example 1 (+ VTK goldens)

Outputs (``tests/golden/*.npz``):
  kernels_small.npz   create_cov / calcGridPoints3D / calcDistanceMatrix on a 4x3x2 grid
  sens_8x6x5.npz      A_sens('grav'|'magn') + A_drill on an 8x6x5 cube
  cubing_<k>_<tag>.npz  full Inversion.cubing on 8x6x5 (k in exp, sparse, matern32; nd>0 and nd=0)
  acquisition.npz     futility_vertical / futility_drill (run_geobo.py:175-235) on random 10x12x8 cubes
  example1.npz / example2.npz
                      the five ``cubing`` inputs the reference's own driver
                      (``geobo/run_geobo.py:396-415``) produced from the committed
                      input files, the live reference outputs, and the six
                      committed result cubes (``examples/results/*/cube_*.vtk``)
                      = the reference's own golden vectors for this path.
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SMALL = dict(xNcube=8, yNcube=6, zNcube=5)


def _cfg_dict(cl):
    keys = ["xmin", "xmax", "ymin", "ymax", "zmax", "zoff", "zLcube", "xNcube", "yNcube", "zNcube",
            "gp_lengthscale", "gp_err", "gp_coeff", "kernelfunc", "optimize_gp", "XMAG", "YMAG", "ZMAG",
            "c_G", "c_SI_TO_MILLIGALS", "c_GCM3_TO_SI", "fcor_grav", "fcor_mag"]
    return {k: getattr(cl, k) for k in keys}


def _sensor_locations(cl):
    # geobo/run_geobo.py:61-65
    import numpy as np
    x_s = np.linspace(0.5, cl.xNcube - 0.5, cl.xNcube) * cl.xvoxsize
    y_s = np.linspace(0.5, cl.yNcube - 0.5, cl.yNcube) * cl.yvoxsize
    z_s = cl.zmax + cl.zoff
    xs, ys, zs = np.meshgrid(x_s, y_s, z_s)
    return np.asarray([xs.flatten(), ys.flatten(), zs.flatten()]).T


def job_kernels_small():
    import numpy as np
    from oracle import ref_loader
    mods = ref_loader.load(ref_loader.write_yaml(dict(xNcube=4, yNcube=3, zNcube=2)))
    k = mods["kernels"]
    out = {}
    pts = k.calcGridPoints3D((4, 3, 2), (122.0, 61.0, 50.0))
    D2 = k.calcDistanceMatrix(pts)
    out["points"], out["D2"] = pts, D2
    for fk in ("exp", "sparse", "matern32"):
        # distinct scales: no de-dup (kernels.py:174-180)
        gl = np.array([244.0, 250.0, 260.0])
        out["cov_%s_distinct" % fk] = k.create_cov(D2, gl, [1.0, 0.2, 0.3], fkernel=fk)
        out["gl_%s_distinct_after" % fk] = gl
        if fk != "matern32":  # equal scales give NaN for matern32
            gl = np.array([244.0, 244.0, 244.0])
            out["cov_%s_equal" % fk] = k.create_cov(D2, gl, [1.0, 0.2, 0.2], fkernel=fk)
            out["gl_%s_equal_after" % fk] = gl  # mutated in place -> [244, 248.88, 244]
    np.savez_compressed(os.path.join(HERE, "kernels_small.npz"), **out)
    print("kernels_small ok", {n: v.shape for n, v in out.items()})


def job_sens_small():
    import numpy as np
    from oracle import ref_loader
    mods = ref_loader.load(ref_loader.write_yaml(SMALL))
    cl, sm = mods["config_loader"], mods["sensormodel"]
    inv = mods["inversion"].Inversion()
    voxelpos = inv.create_cubegeometry()
    loc = _sensor_locations(cl)
    Ag, _ = sm.A_sens(cl.magneticField * 0.0, loc, inv.Edges, "grav")
    Am, _ = sm.A_sens(cl.magneticField, loc, inv.Edges, "magn")
    # a tilted field exercises every term of magn_func (sensormodel.py:113-133)
    Bt = np.array([0.3, -0.2, 0.9]) * 1e-3
    Amt, _ = sm.A_sens(Bt, loc, inv.Edges, "magn")
    drill_idx = np.array([3, 17, 100, 239])
    Ad = sm.A_drill(voxelpos[:, drill_idx], voxelpos)
    np.savez_compressed(os.path.join(HERE, "sens_8x6x5.npz"), cfg=json.dumps(_cfg_dict(cl)),
                        Edges=inv.Edges, voxelpos=voxelpos, locations=loc, A_grav=Ag, A_magn=Am,
                        B_tilt=Bt, A_magn_tilt=Amt, drill_idx=drill_idx, A_drill=Ad)
    print("sens ok", Ag.shape, Am.shape, float(Ag.sum()), float(Am.sum()))


def _truth_cubes(cl, voxelpos):
    # 'cylinders' model, geobo/simcube.py:83-92 evaluated by formula on the reference geometry
    import numpy as np
    x3, y3, z3 = (v.reshape(cl.yNcube, cl.xNcube, cl.zNcube) for v in voxelpos)
    rad = cl.yLcube / 18.0
    rc1 = (y3 - cl.yLcube / 1.3 - rad) ** 2 + (z3 + cl.zLcube / 4 - rad) ** 2
    rc2 = (y3 - cl.yLcube / 4.0 - rad) ** 2 + (z3 + cl.zLcube / 4 - rad) ** 2
    density = x3 * 0.0 + 0.1
    density[rc2 <= rad ** 2] = 1.0
    density[rc1 <= rad ** 2] = 1.0
    density[(x3 < cl.xLcube / 5.0) | (x3 > cl.xLcube * 4.0 / 5.0)] = 0.1
    # smooth perturbation so the tiny cube has non-degenerate data
    density = density + 0.05 * np.sin(x3 / 400.0) * np.cos(y3 / 300.0) + 0.02 * z3 / cl.zLcube
    return density, cl.gp_coeff[1] * density


def job_cubing_small(kernelfunc, nd, gl_mult=None, tag=None):
    import numpy as np
    from oracle import ref_loader
    mods = ref_loader.load(ref_loader.write_yaml(dict(SMALL, kernelfunc=kernelfunc)))
    cl, sm = mods["config_loader"], mods["sensormodel"]
    inv = mods["inversion"].Inversion()
    voxelpos = inv.create_cubegeometry()
    xN, yN, zN = cl.xNcube, cl.yNcube, cl.zNcube
    inv.xxx = voxelpos[0].reshape(xN, yN, zN)
    inv.yyy = voxelpos[1].reshape(xN, yN, zN)
    inv.zzz = voxelpos[2].reshape(xN, yN, zN)
    loc = _sensor_locations(cl)
    dens, mags = _truth_cubes(cl, voxelpos)
    Ag, _ = sm.A_sens(cl.magneticField * 0.0, loc, inv.Edges, "grav")
    Am, _ = sm.A_sens(cl.magneticField, loc, inv.Edges, "magn")
    grav = (Ag @ dens.flatten()).astype(np.float32).astype(np.float64)
    mag = (Am @ mags.flatten()).astype(np.float32).astype(np.float64)
    rng = np.random.default_rng(0)
    drilldata0 = np.zeros(xN * yN * zN)
    if nd:
        idx = np.sort(rng.choice(xN * yN * zN, nd, replace=False))
        drilldata0[idx] = dens.flatten()[idx]
    drilldata0 = drilldata0.reshape(xN, yN, zN)
    drillfield = drilldata0[drilldata0 != 0]
    if gl_mult is not None:
        inv.gp_length = inv.gp_length * np.asarray(gl_mult)
    gl_before = np.array(inv.gp_length, dtype=float)
    with np.errstate(all="ignore"):
        res = inv.cubing(grav, mag, drillfield, loc, drilldata0)
    names = ["density_rec", "magsus_rec", "drill_rec", "density_var", "magsus_var", "drill_var"]
    out = {n: r for n, r in zip(names, res)}
    out.update(cfg=json.dumps(_cfg_dict(cl)), grav=grav, mag=mag, drillfield=drillfield,
               sensor_locations=loc, drilldata0=drilldata0, logl=inv.logl, gl_before=gl_before,
               gl_after=np.array(inv.gp_length, dtype=float), mu_rec=inv.mu_rec, Fs3=inv.Fs3)
    tag = tag or ("nd%d" % nd)
    np.savez_compressed(os.path.join(HERE, "cubing_%s_%s.npz" % (kernelfunc, tag)), **out)
    print("cubing", kernelfunc, tag, "logl", inv.logl, "gl_after", inv.gp_length)


def job_acquisition():
    """futility_vertical / futility_drill of the UNMODIFIED reference (geobo/run_geobo.py:175-235): the two function
    definitions are extracted from the reference source with ``ast`` (the module itself is a script that runs the
    whole pipeline at import) and executed with the module globals they read set explicitly."""
    import ast
    import numpy as np
    from oracle import ref_loader
    mods = ref_loader.load(ref_loader.write_yaml(dict(xNcube=12, yNcube=10, zNcube=8)))
    cl = mods["config_loader"]
    src = open("/root/reference/geobo/run_geobo.py").read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("futility_vertical", "futility_drill")]
    usrc = open("/root/reference/geobo/utils.py").read()
    utree = ast.parse(usrc)
    wanted = [n for n in utree.body if isinstance(n, ast.FunctionDef) and n.name == "spherical2cartes"] + wanted
    rng = np.random.default_rng(7)
    shape = (cl.yNcube, cl.xNcube, cl.zNcube)
    rec = rng.standard_normal(shape)
    var = rng.uniform(0.05, 1.0, shape)
    costs = rng.uniform(0.0, 2.0, shape)
    g = dict(np=np, drill_rec=rec, drill_var=var, kappa=1.5, beta=0.3, zLcube=cl.zLcube, zmax=cl.zmax,
             xvoxsize=cl.xvoxsize, yvoxsize=cl.yvoxsize, zvoxsize=cl.zvoxsize)
    exec(compile(ast.Module(body=wanted, type_ignores=[]), "reference_acquisition", "exec"), g)
    pv = np.array([[a, b] for a in range(-1, shape[0] + 1) for b in range(-1, shape[1] + 1)], dtype=float)
    pv = np.vstack([pv, [[2.4, 3.6], [np.nan, 1.0], [4.5, 5.5]]])
    fv = np.array([g["futility_vertical"](p) for p in pv])
    fvc = np.array([g["futility_vertical"](p, costs) for p in pv])
    pd_ = np.column_stack([rng.uniform(0.0, shape[0] * cl.xvoxsize, 400), rng.uniform(0.0, shape[1] * cl.yvoxsize, 400),
                           rng.uniform(0, 360, 400), rng.uniform(30, 90, 400)])
    with np.errstate(all="ignore"):
        fd = np.array([g["futility_drill"](p) for p in pd_])
        fdc = np.array([g["futility_drill"](p, costs) for p in pd_])
    np.savez_compressed(os.path.join(HERE, "acquisition.npz"), cfg=json.dumps(_cfg_dict(cl)), rec=rec, var=var, costs=costs,
                        kappa=1.5, beta=0.3, pv=pv, fv=fv, fvc=fvc, pd=pd_, fd=fd, fdc=fdc)
    print("acquisition.npz: %d vertical, %d drill candidates (%d inside the cube)" % (len(pv), len(pd_), int((fd != 0).sum())))


def _install_driver_stubs():
    """Stub the plotting / raster / VTK dependencies of geobo/run_geobo.py (absent here)."""
    import types
    from unittest.mock import MagicMock
    import numpy as np
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d", "skimage", "skimage.measure"]:
        sys.modules[name] = MagicMock()

    def read_tiff(path):
        """Minimal baseline-TIFF reader: uncompressed strips, float32/float64 samples."""
        import struct
        with open(path, "rb") as f:
            raw = f.read()
        bo = "<" if raw[:2] == b"II" else ">"
        (ifd,) = struct.unpack(bo + "I", raw[4:8])
        (n,) = struct.unpack(bo + "H", raw[ifd:ifd + 2])
        tags = {}
        tsize = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 11: 4, 12: 8, 16: 8}
        tfmt = {1: "B", 3: "H", 4: "I", 11: "f", 12: "d", 16: "Q"}
        for i in range(n):
            e = raw[ifd + 2 + 12 * i: ifd + 14 + 12 * i]
            tag, typ, cnt = struct.unpack(bo + "HHI", e[:8])
            if typ not in tfmt:
                continue
            nbytes = tsize[typ] * cnt
            data = e[8:8 + nbytes] if nbytes <= 4 else raw[struct.unpack(bo + "I", e[8:12])[0]:][:nbytes]
            tags[tag] = struct.unpack(bo + tfmt[typ] * cnt, data)
        w, h, bits = tags[256][0], tags[257][0], tags[258][0]
        assert tags.get(259, (1,))[0] == 1, "compressed TIFF not supported"
        buf = b"".join(raw[o:o + c] for o, c in zip(tags[273], tags[279]))
        dt = {32: "f4", 64: "f8"}[bits]
        return np.frombuffer(buf, dtype=bo + dt, count=w * h).reshape(h, w).astype(dt)

    class _Img:
        def __init__(self, path):
            self.path = path

        def read(self, band):
            return read_tiff(self.path)

    rasterio = types.ModuleType("rasterio")
    rasterio.open = lambda path, *a, **k: _Img(path)
    rasterio.float32 = "float32"
    sys.modules["rasterio"] = rasterio

    captured = []

    class UniformGrid:
        def __init__(self):
            self.cell_arrays = {}

        def save(self, fname):
            captured.append((fname, np.array(self.cell_arrays["values"]), tuple(self.dimensions)))

    pyvista = types.ModuleType("pyvista")
    pyvista.UniformGrid = UniformGrid
    sys.modules["pyvista"] = pyvista
    return captured


def job_example(which):
    """Run the reference's own driver on its committed inputs; capture cubing inputs/outputs."""
    import numpy as np
    from oracle import ref_loader, vtkio
    ref = ref_loader.REFERENCE_ROOT
    sub = {"1": "synthetic", "2": "sample"}[which]
    res = {"1": "cylinders", "2": "sample"}[which]
    base = os.path.join(ref, "examples", "settings_example%s.yaml" % which)
    ypath = ref_loader.write_yaml(dict(inpath=os.path.join(ref, "examples", "testdata", sub) + os.sep,
                                       gen_simulation=False, plot3d=False, plot_vertical=False,
                                       bayesopt_vertical=False, bayesopt_nonvertical=False), base=base)
    _install_driver_stubs()
    mods = ref_loader.load(ypath)
    inv_mod = mods["inversion"]
    rec = {}
    orig = inv_mod.Inversion.cubing

    def spy(self, gravfield, magfield, drillfield, sensor_locations, drilldata0):
        rec.update(grav=np.array(gravfield), mag=np.array(magfield), drillfield=np.array(drillfield),
                   sensor_locations=np.array(sensor_locations), drilldata0=np.array(drilldata0))
        out = orig(self, gravfield, magfield, drillfield, sensor_locations, drilldata0)
        rec.update(logl=self.logl, gl_after=np.array(self.gp_length, dtype=float),
                   M=self.Asens3.shape[0])
        for n, r in zip(["density_rec", "magsus_rec", "drill_rec", "density_var", "magsus_var", "drill_var"], out):
            rec["live_" + n] = r
        return out

    inv_mod.Inversion.cubing = spy
    import importlib
    importlib.import_module("geobo.run_geobo")  # the import *is* the run (geobo/main.py:16-26)
    cl = mods["config_loader"]
    gold = {}
    for n, f in [("density_rec", "cube_density"), ("magsus_rec", "cube_magsus"), ("drill_rec", "cube_drill"),
                 ("density_var", "cube_density_variance"), ("magsus_var", "cube_magsus_variance"),
                 ("drill_var", "cube_drill_variance")]:
        gold["gold_" + n] = vtkio.read_cube(os.path.join(ref, "examples", "results", res, f + ".vtk"))
        live = rec["live_" + n]
        err = np.abs(live - gold["gold_" + n]).max() / np.abs(gold["gold_" + n]).max()
        print("example", which, n, "live-vs-committed-VTK rel err %.2e" % err)
        # the live double-precision outputs are redundant with the goldens to <=4e-8: keep only goldens + mu check
    keep = {k: v for k, v in rec.items() if not k.startswith("live_")}
    keep["live_density_rec"] = rec["live_density_rec"]
    keep["live_drill_var"] = rec["live_drill_var"]
    np.savez_compressed(os.path.join(HERE, "example%s.npz" % which), cfg=json.dumps(_cfg_dict(cl)), **keep, **gold)
    print("example", which, "logl", rec["logl"], "M", rec["M"], "gl_after", rec["gl_after"])


JOBS = [
    ["kernels_small"], ["sens_small"],
    ["cubing_small", "exp", "7"], ["cubing_small", "sparse", "7"], ["cubing_small", "matern32", "7"],
    ["cubing_small", "exp", "0"],
    ["example", "1"], ["example", "2"],
    ["acquisition"],
]


def main(argv):
    if not argv:
        for job in JOBS:
            print("==", job, flush=True)
            subprocess.run([sys.executable, "-W", "ignore", __file__] + job, check=True)
        return
    cmd = argv[0]
    if cmd == "small":
        for job in JOBS[:6]:
            subprocess.run([sys.executable, "-W", "ignore", __file__] + job, check=True)
    elif cmd == "kernels_small":
        job_kernels_small()
    elif cmd == "sens_small":
        job_sens_small()
    elif cmd == "cubing_small":
        kf, nd = argv[1], int(argv[2])
        # matern32 needs distinct scales (equal scales -> 0/0 in the cross term)
        job_cubing_small(kf, nd, gl_mult=[1.0, 1.01, 1.02] if kf == "matern32" else None)
    elif cmd == "example":
        job_example(argv[1])
    elif cmd == "acquisition":
        job_acquisition()
    else:
        raise SystemExit("unknown job %r" % (argv,))


if __name__ == "__main__":
    main(sys.argv[1:])
