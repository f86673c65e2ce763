"""Synthetic cubes and their simulated surveys, with the surface of ``geobo/simcube.py``.

``create_syncube`` evaluates the three analytic models of the reference (``simcube.py:54-94``: 'cylinders', 'layers_2',
'layers_3') on the voxel centres -- O(N) host arithmetic, any cube size.  ``create_synsurvey`` is the expensive half in
the reference (two ``A_sens`` triple loops, ``simcube.py:143-146``): here the sensitivities are built on the GPU and the
surveys are the device products ``A_grav . density`` and ``A_magn . magsus`` (``gb_forward``).  File output follows the
reference (``simcube_<model>.vtk/.csv``, ``simdrill_<model>.csv``, ``simsurveydata_<model>.csv`` under ``inpath``) and is
skipped when the settings carry no ``inpath``; the GeoTIFF pair of ``create_simdata`` needs ``rasterio`` and is written
only where that package exists.
"""
import os
import random

import numpy as np

from . import _lib
from . import config_loader as _cfg
from . import cubeshow as cs

MODELS = ("layers_2", "layers_3", "cylinders")
# (amplitude, upper edge, lower edge) of the layers as fractions of the cube depth, simcube.py:57,63,69
_LAYERS = {"layer1": (4., 0.3, 0.325), "layer2": (8., 0.25, 0.275), "layer3": (6., 0.35, 0.375)}


def _layer(z3, zshift, amp, upper, lower):
    """One buried layer: a difference of two logistic steps in depth, cut at its 90th percentile into {0, max}."""
    zL = _cfg.zLcube
    with np.errstate(over="ignore"):
        step_u = 1. / (1 + np.exp(-2 * (-z3 - zL * upper + zshift)))
        step_l = 1. / (1 + np.exp(-2 * (-z3 - zL * lower + zshift)))
    layer = amp * (step_u - step_l)
    cut = np.percentile(layer, 90)
    top = layer >= cut
    layer[~top] = 0.
    layer[top] = layer.max()
    return layer


def _csv(path, header, columns):
    """``DataFrame.to_csv(index=False)`` of float columns: shortest round-trip ``repr`` of every value."""
    with open(path, "w") as f:
        f.write(",".join(header) + "\n")
        for row in zip(*columns):
            f.write(",".join(v if isinstance(v, str) else repr(float(v)) for v in row) + "\n")


def create_syncube(modelname, voxelpos):
    """Density and magnetic-susceptibility cubes (yNcube, xNcube, zNcube) of a synthetic model (``simcube.py:34-117``).

    modelname: 'layers_2', 'layers_3' or 'cylinders'; voxelpos: voxel centres (x, y, z) from
    ``Inversion.create_cubegeometry``.  With ``inpath`` in the settings also writes ``simcube_<model>.vtk/.csv`` and a
    ``simdrill_<model>.csv`` of four vertical holes at random columns, like the reference."""
    if modelname not in MODELS:
        raise ValueError("modelname must be one of %s, got %r" % (", ".join(MODELS), modelname))
    print("Creating simulated cube data ...")
    shape = (_cfg.yNcube, _cfg.xNcube, _cfg.zNcube)
    x3, y3, z3 = (np.asarray(v, dtype=float).reshape(shape) for v in voxelpos)
    if modelname == "cylinders":
        # two horizontal cylinders along x at a quarter of the depth, cut off near both x ends (simcube.py:83-92)
        rad = _cfg.yLcube / 18.
        depth2 = (z3 + _cfg.zLcube / 4 - rad)**2
        in_c1 = (y3 - _cfg.yLcube / 1.3 - rad)**2 + depth2 <= rad**2
        in_c2 = (y3 - _cfg.yLcube / 4. - rad)**2 + depth2 <= rad**2
        density = np.where(in_c1 | in_c2, 1., x3 * 0. + 0.1)
        density[(x3 < _cfg.xLcube / 5.) | (x3 > _cfg.xLcube * 4. / 5.)] = 0.1
    else:
        # layers dipping along y; the dip is centred on zLcube / 2 for 'layers_2' (simcube.py:56) and on yLcube / 2 for 'layers_3' (:68)
        centre = _cfg.zLcube / 2 if modelname == "layers_2" else _cfg.yLcube / 2.
        with np.errstate(over="ignore"):
            zshift = _cfg.zLcube / 8. * 1. / (1 + np.exp(2. * (-y3 + centre)))
        density = 0.5 + _layer(z3, zshift, *_LAYERS["layer1"]) + _layer(z3, zshift, *_LAYERS["layer2"])
        if modelname == "layers_3":
            density = density + _layer(z3, zshift, *_LAYERS["layer3"])
    magsus = _cfg.gp_coeff[1] * density          # simple correlation between the two properties
    inpath = getattr(_cfg, "inpath", None)
    if inpath:
        os.makedirs(inpath, exist_ok=True)
        origin = (np.min(voxelpos[0]), np.min(voxelpos[1]), np.min(voxelpos[2]))
        cs.create_vtkcube(density, origin, (_cfg.xvoxsize, _cfg.yvoxsize, _cfg.zvoxsize),
                          fname=os.path.join(inpath, 'simcube_' + modelname + '.vtk'))
        cols = [x3.flatten(), y3.flatten(), z3.flatten(), density.flatten(), magsus.flatten()]
        _csv(os.path.join(inpath, 'simcube_' + modelname + '.csv'), ['x', 'y', 'z', 'DENSITY', 'MAGSUS'], cols)
        # four vertical drill holes: two random x columns times two random y columns (simcube.py:104-113)
        xd = np.asarray([random.randint(2, _cfg.xNcube - 2) for _ in range(2)]) * _cfg.xvoxsize + 0.5 * _cfg.xvoxsize
        yd = np.asarray([random.randrange(2, _cfg.yNcube - 2) for _ in range(2)]) * _cfg.yvoxsize + 0.5 * _cfg.yvoxsize
        sel = np.isin(cols[0], xd) & np.isin(cols[1], yd)
        site = ['SiteID_' + str(x) + str(y) for x, y in zip(cols[0][sel], cols[1][sel])]
        _csv(os.path.join(inpath, 'simdrill_' + modelname + '.csv'), ['x', 'y', 'z', 'DENSITY', 'MAGSUS', 'SiteID'],
             [col[sel] for col in cols] + [site])
    return density, magsus


def create_synsurvey(modelname, density, magsus, ctx=None):
    """Simulated gravity and magnetic maps (yNcube, xNcube) of the cubes (``simcube.py:119-159``): one sensor above every
    voxel column at height ``zoff``; both forward models are evaluated on the GPU."""
    print("Creating simulated sensor data...")
    from .inversion import Inversion
    inv = Inversion()
    inv.create_cubegeometry()
    xN, yN, zN = _cfg.xNcube, _cfg.yNcube, _cfg.zNcube
    xnew = np.arange(_cfg.xvoxsize / 2., _cfg.xLcube + _cfg.xvoxsize / 2., _cfg.xvoxsize)
    ynew = np.arange(_cfg.yvoxsize / 2., _cfg.yLcube + _cfg.yvoxsize / 2., _cfg.yvoxsize)
    xx, yy = np.meshgrid(xnew, ynew)
    zz = xx * 0. + _cfg.zoff
    sensor_locations = np.asarray([xx.flatten(), yy.flatten(), zz.flatten()]).T
    N = xN * yN * zN
    prob = _lib.Problem(ctx or _lib.default_context(), (xN, yN, zN), (_cfg.xvoxsize, _cfg.yvoxsize, _cfg.zvoxsize), inv.Edges,
                        sensor_locations, _cfg.magneticField, _cfg.c_MILLIGALS_UNITS, _cfg.fcor_grav, 1.0, _cfg.fcor_mag,
                        np.zeros(0, dtype=np.int64), 0, min(N, 128))     # only the sensitivities are needed: smallest voxel shard
    try:
        gravfield = prob.forward("grav", np.asarray(density, dtype=float).flatten())
        magfield = prob.forward("magn", np.asarray(magsus, dtype=float).flatten())
    finally:
        prob.close()
    inpath = getattr(_cfg, "inpath", None)
    if inpath:
        os.makedirs(inpath, exist_ok=True)
        _csv(os.path.join(inpath, 'simsurveydata_' + modelname + '.csv'), ['X', 'Y', 'GRAVITY', 'MAGNETIC'],
             [xx.flatten(), yy.flatten(), gravfield, magfield])
    return gravfield.reshape(xx.shape), magfield.reshape(xx.shape)


def create_simdata(modelname="cylinders", plot=True):
    """Cubes + surveys of one model written under ``inpath`` (``simcube.py:162-220``).  The float32 GeoTIFF maps the
    reference's reader expects are written when ``rasterio`` is importable; plotting is not part of this package."""
    from .inversion import Inversion
    voxelpos = Inversion().create_cubegeometry()
    density, magsus = create_syncube(modelname, voxelpos)
    grav2D, magn2D = create_synsurvey(modelname, density, magsus)
    inpath = getattr(_cfg, "inpath", None)
    if inpath:
        try:
            import rasterio
        except ImportError:
            print("rasterio not installed: gravity_simdata_%s.tif / magnetic_simdata_%s.tif not written" % (modelname, modelname))
        else:
            for name, img in (("gravity", grav2D), ("magnetic", magn2D)):
                with rasterio.open(os.path.join(inpath, name + '_simdata_' + modelname + '.tif'), 'w', driver='GTiff', width=img.shape[1],
                                   height=img.shape[0], count=1, dtype='float32') as dst:
                    dst.write(img.astype('float32'), 1)
    return density, magsus, grav2D, magn2D
