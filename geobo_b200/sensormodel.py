"""``geobo.sensormodel`` surface on the B200 (reference: ``geobo/sensormodel.py``).

Gravity / magnetic prism sensitivities after Li & Oldenburg and the drill-core selection
matrix, computed by the CUDA library in float64.
"""
import numpy as np

from . import _lib
from . import config_loader as _cfg


def A_sens(magneticField, locations, Edges, func):
    """Gravity ('grav') or magnetic ('magn') forward-model matrix (``sensormodel.py:29-93``).

    Returns ``(sens, None)``: the reference also returns the per-sensor corner potentials
    ``result_ez`` (Ns x Nedge), which every caller in the reference discards; they are not
    materialised here.  The sensor count is ``xNcube*yNcube`` as in the reference (``:54,58``).
    """
    _cfg.require("xNcube", "yNcube", "zNcube")
    if func not in ("grav", "magn"):
        print('function not supported')                         # sensormodel.py:75-76
        raise ValueError("func must be 'grav' or 'magn'")
    Edges = np.asarray(Edges, dtype=float)
    ncube = (_cfg.xNcube, _cfg.yNcube, _cfg.zNcube)
    nsens = _cfg.xNcube * _cfg.yNcube
    loc = np.asarray(locations, dtype=float)[:nsens]
    if func == 'grav':
        mul, div = _cfg.c_MILLIGALS_UNITS, _cfg.fcor_grav      # :88-89
    else:
        mul, div = 1.0, _cfg.fcor_mag                           # :90-91
    sens = _lib.default_context().a_sens(func, np.asarray(magneticField, dtype=float), loc, Edges, ncube, mul, div)
    return sens, None


def grav_func(x, y, z):
    """Vertical gravity corner potential (``sensormodel.py:96-110``), elementwise."""
    shape = np.broadcast(np.asarray(x), np.asarray(y), np.asarray(z)).shape
    return _lib.default_context().corner_func("grav", x, y, z).reshape(shape)


def magn_func(x, y, z, bx, by, bz):
    """Magnetic corner potential (``sensormodel.py:113-133``), elementwise."""
    shape = np.broadcast(np.asarray(x), np.asarray(y), np.asarray(z)).shape
    return _lib.default_context().corner_func("magn", x, y, z, [bx, by, bz]).reshape(shape)


def A_drill(loc, voxelpos):
    """Drill-hole filter matrix with sensitivity 1 (``sensormodel.py:136-153``): loc (3, Ndrill), voxelpos (3, Nvoxel)."""
    loc = np.asarray(loc, dtype=float)
    vp = np.asarray([np.asarray(voxelpos[0]).flatten(), np.asarray(voxelpos[1]).flatten(), np.asarray(voxelpos[2]).flatten()])
    if loc.ndim != 2 or loc.shape[1] == 0:
        return np.zeros((0 if loc.ndim != 2 else loc.shape[1], vp.shape[1]))
    return _lib.default_context().a_drill(loc, vp)
