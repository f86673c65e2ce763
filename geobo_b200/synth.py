"""Deterministic synthetic inputs for benchmarks and smoke tests.

Same recipe as the reference's simulator: analytic 'cylinders' density / susceptibility cubes
(``geobo/simcube.py:83-92``, pure functions of the voxel coordinates, so any cube size works),
forward-simulated surveys ``grav = A_g . rho``, ``mag = A_m . chi`` (``simcube.py:147-150``) rounded
through float32 like the GeoTIFF round trip (``:196-199``), sensors over the voxel-column centres at
height ``zmax + zoff`` (``run_geobo.py:61-65``) and ``nd`` drilled voxels drawn with
``numpy.random.default_rng(seed)``.  The forward simulation runs on the GPU (``gb_forward``).
"""
import numpy as np

from . import _lib
from . import config_loader as _cfg

BASE_SETTINGS = dict(
    xmin=0, xmax=3050, ymin=0, ymax=1952, zmax=0, zoff=1, zLcube=800.0, xNcube=25, yNcube=16, zNcube=16,
    gp_lengthscale=2, gp_err=[0.1, 0.1, 0.1], gp_coeff=[1.0, 0.2, 0.2], kernelfunc="sparse", optimize_gp=False,
    XMAG=0, YMAG=0, ZMAG=1, c_G=6.673848e-11, c_SI_TO_MILLIGALS=10000, c_GCM3_TO_SI=1000.0, fcor_grav=1.0,
    fcor_mag=0.001, gen_simulation=False, modelname="cylinders")   # examples/settings_example1.yaml


def settings(xN, yN, zN, kernelfunc="exp", **extra):
    cfg = dict(BASE_SETTINGS, xNcube=int(xN), yNcube=int(yN), zNcube=int(zN), kernelfunc=kernelfunc)
    cfg.update(extra)
    return cfg


def sensor_grid():
    """run_geobo.py:61-65"""
    x_s = np.linspace(0.5, _cfg.xNcube - 0.5, _cfg.xNcube) * _cfg.xvoxsize
    y_s = np.linspace(0.5, _cfg.yNcube - 0.5, _cfg.yNcube) * _cfg.yvoxsize
    xs, ys, zs = np.meshgrid(x_s, y_s, _cfg.zmax + _cfg.zoff)
    return np.asarray([xs.flatten(), ys.flatten(), zs.flatten()]).T


def cylinders(voxelpos):
    """simcube.py:83-92 plus a smooth background term."""
    shp = (_cfg.yNcube, _cfg.xNcube, _cfg.zNcube)
    x3, y3, z3 = (np.asarray(v).reshape(shp) for v in voxelpos)
    rad = _cfg.yLcube / 18.
    rc1 = (y3 - _cfg.yLcube / 1.3 - rad)**2 + (z3 + _cfg.zLcube / 4 - rad)**2
    rc2 = (y3 - _cfg.yLcube / 4. - rad)**2 + (z3 + _cfg.zLcube / 4 - rad)**2
    density = x3 * 0. + 0.1
    density[rc2 <= rad**2] = 1.
    density[rc1 <= rad**2] = 1.
    density[(x3 < _cfg.xLcube / 5.) | (x3 > _cfg.xLcube * 4. / 5.)] = 0.1
    # smooth background so that drill samples are never all equal (constant samples have zero std
    # and normalise to NaN in the reference, inversion.py:213-214)
    density = density + 0.05 * np.sin(x3 / 400.) * np.cos(y3 / 300.) + 0.02 * z3 / _cfg.zLcube
    return density, _cfg.gp_coeff[1] * density


def make_inputs(nd=0, seed=0, ctx=None):
    """Returns dict(grav, mag, drillfield, sensor_locations, drilldata0) for the loaded settings."""
    from .inversion import Inversion
    ctx = ctx or _lib.default_context()
    inv = Inversion()
    voxelpos = inv.create_cubegeometry()
    xN, yN, zN = _cfg.xNcube, _cfg.yNcube, _cfg.zNcube
    N = xN * yN * zN
    loc = sensor_grid()
    density, magsus = cylinders(voxelpos)
    prob = _lib.Problem(ctx, (xN, yN, zN), (_cfg.xvoxsize, _cfg.yvoxsize, _cfg.zvoxsize), inv.Edges, loc,
                        _cfg.magneticField, _cfg.c_MILLIGALS_UNITS, _cfg.fcor_grav, 1.0, _cfg.fcor_mag,
                        np.zeros(0, dtype=np.int64), 0, min(N, 128))   # tiny shard: only A is needed here
    grav = prob.forward("grav", density.ravel()).astype(np.float32).astype(np.float64)
    mag = prob.forward("magn", magsus.ravel()).astype(np.float32).astype(np.float64)
    prob.close()
    drilldata0 = np.zeros(N)
    if nd:
        idx = np.random.default_rng(seed).choice(N, nd, replace=False)
        drilldata0[idx] = density.ravel()[idx]      # non-zero by construction (density >= 0.1)
    drilldata0 = drilldata0.reshape(xN, yN, zN)
    return dict(grav=grav, mag=mag, drillfield=drilldata0[drilldata0 != 0], sensor_locations=loc, drilldata0=drilldata0)
