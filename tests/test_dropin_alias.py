"""The ``geobo`` alias package: the reference's module names resolve to the B200 implementation, so its driver's import lines
(``geobo/run_geobo.py:385-389``) and call sequence (``:391-425``) run unedited.

CPU: every alias is the ``geobo_b200`` module object itself; the star-import of the settings works through the alias; in the build
container the UNMODIFIED reference driver ``run_geobo.py`` is executed inside this package (I/O and plotting dependencies stubbed as
in tests/golden/make_golden.py) up to its first device call.  GPU: the driver's call sequence on the reference's committed example-1
inputs reproduces the committed VTK goldens on the tensor-core path."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import CUBES, ROOT, check_subprocess, load_golden, normwise_err

ALIASES = ["config_loader", "kernels", "sensormodel", "inversion", "utils", "simcube", "cubeshow", "acquisition"]


def test_aliases_are_the_implementation_modules():
    import importlib
    for name in ALIASES:
        a = importlib.import_module("geobo." + name)
        b = importlib.import_module("geobo_b200." + name)
        assert a is b, name
    import geobo
    from geobo import inversion, kernels                  # `from . import inversion` style (run_geobo.py:388)
    assert inversion.Inversion.__module__ == "geobo_b200.inversion" and geobo.kernels is kernels


def test_star_import_of_settings_through_the_alias():
    f = load_golden("example1.npz")
    import json
    from geobo import config_loader
    config_loader.load_settings(json.loads(str(f["cfg"])), make_outpath=False)
    ns = {}
    exec("from geobo.config_loader import *", ns)            # run_geobo.py:385
    assert ns["xNcube"] == 25 and ns["xvoxsize"] == config_loader.xvoxsize and "magneticField" in ns


DRIVER = r'''
import sys, json, numpy as np
sys.path.insert(0, %(root)r)
f = np.load(%(root)r + "/tests/golden/example1.npz")      # (not through tests/conftest.py: it pins the suite's default precision)
import geobo_b200.config_loader as _cl
_cl.load_settings(json.loads(str(f["cfg"])), make_outpath=False)
# ---- the reference driver's own lines (geobo/run_geobo.py:385-415), imports unedited
from geobo.config_loader import *  # loads settings
from geobo.utils import *
from geobo import cubeshow as cs
from geobo import inversion
from geobo import simcube
inv = inversion.Inversion()
voxelpos = inv.create_cubegeometry()
xxx, yyy, zzz = voxelpos
xxx = inv.xxx = xxx.reshape(xNcube, yNcube, zNcube)
yyy = inv.yyy = yyy.reshape(xNcube, yNcube, zNcube)
zzz = inv.zzz = zzz.reshape(xNcube, yNcube, zNcube)
gravfield, magfield, sensor_locations = f["grav"], f["mag"], f["sensor_locations"]       # read_surveydata() of the example
drilldata0 = f["drilldata0"]
drillfield = drilldata0[drilldata0 != 0]
density_rec, magsus_rec, drill_rec, density_var, magsus_var, drill_var = inv.cubing(gravfield, magfield, drillfield, sensor_locations, drilldata0)
origin = (voxelpos[0].min(), voxelpos[1].min(), voxelpos[2].min())
voxelsize = (xvoxsize, yvoxsize, zvoxsize)
cs.create_vtkcube(density_rec, origin, voxelsize, fname=sys.argv[1] + "/cube_density.vtk")
np.save(sys.argv[1] + "/cubes.npy", np.stack([density_rec, magsus_rec, drill_rec, density_var, magsus_var, drill_var]))
print("PRECISION", inv.precision_used)
'''


@pytest.mark.gpu
def test_gpu_reference_driver_sequence_through_the_alias_reproduces_the_vtk_goldens(tmp_path):
    env = {k: v for k, v in os.environ.items() if k != "GEOBO_B200_DEFAULT_PRECISION"}      # the product default: precision auto
    r = subprocess.run([sys.executable, "-c", DRIVER % dict(root=ROOT), str(tmp_path)], capture_output=True, text=True, timeout=600, env=env)
    check_subprocess(r)
    assert "PRECISION int8x5" in r.stdout                    # zNcube = 16: the tensor-core path, without touching the YAML
    f = load_golden("example1.npz")
    cubes = np.load(tmp_path / "cubes.npy")
    for n, a in zip(CUBES, cubes):
        assert normwise_err(a, f["gold_" + n]) < 1e-6, n
    head = open(tmp_path / "cube_density.vtk", "rb").read(200)
    assert head.startswith(b"# vtk DataFile Version")


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/geobo"), reason="needs the reference tree (build container only)")
def test_unmodified_reference_driver_imports_resolve_to_the_alias_package(tmp_path):
    """geobo/run_geobo.py of the reference, byte for byte, executed as ``geobo.run_geobo`` inside a package whose hot-path modules
    are this repo's aliases: it must get through its imports, settings, geometry and survey reading and reach its first device
    call (no GPU here: the library refuses at gb_ctx_create)."""
    pkg = tmp_path / "site" / "geobo"
    pkg.mkdir(parents=True)
    for name in os.listdir(os.path.join(ROOT, "geobo")):
        if name.endswith(".py"):
            (pkg / name).write_text(open(os.path.join(ROOT, "geobo", name)).read())
    (pkg / "run_geobo.py").symlink_to("/root/reference/geobo/run_geobo.py")
    out = tmp_path / "out"
    out.mkdir()
    import yaml
    cfg = yaml.safe_load(open("/root/reference/examples/settings_example1.yaml"))
    cfg.update(inpath="/root/reference/examples/testdata/synthetic/", outpath=str(out) + os.sep, plot3d=False, plot_vertical=False,
               bayesopt_vertical=False, bayesopt_nonvertical=False, gen_simulation=False)
    y = tmp_path / "settings.yaml"
    y.write_text(yaml.safe_dump(cfg))
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests/golden')
import make_golden
sys.path.insert(0, %r)    # the temp package before the repo root
sys.argv = ['main.py', %r]
make_golden._install_driver_stubs()          # matplotlib / rasterio / pyvista / skimage are absent here (I/O and plotting only)
try:
    import geobo.run_geobo
except Exception as e:
    import traceback; tb = traceback.extract_tb(e.__traceback__)
    print('STOPPED', type(e).__name__, str(e)[:200])
    print('FRAMES', [(fr.filename.split('/')[-1], fr.name) for fr in tb])
""" % (ROOT, ROOT, str(tmp_path / "site"), str(y))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "STOPPED GeoboB200Error" in r.stdout and "no CUDA device" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    # the first device call of the driver is align_drill (geobo.utils through the star-import, run_geobo.py:128) inside its
    # read_drilldata -- already the CUDA implementation; with a GPU the run continues into inv.cubing
    assert "('run_geobo.py', '<module>')" in r.stdout and "('utils.py', 'align_drill')" in r.stdout and "('_lib.py', 'default_context')" in r.stdout, r.stdout[-3000:]
