"""Bayesian-optimisation acquisition functions of ``geobo/run_geobo.py:175-235`` evaluated on the GPU.

The reference calls ``futility_vertical`` / ``futility_drill`` one point at a time from ``scipy.optimize.shgo``
(``run_geobo.py:246-362``) on module globals ``drill_rec``, ``drill_var``, ``kappa``, ``beta``.  The objective is
piecewise constant in the voxel indices, so the exhaustive sweep over every voxel column (vertical drillholes) or
over a candidate batch (non-vertical) is cheap on the device and needs no optimiser.  The scalar functions keep the
reference's signatures and sign convention (they return MINUS the utility); the cubes are passed explicitly or
taken from ``set_cubes``.
"""
import numpy as np

from . import _lib
from . import config_loader as _cfg

_state = {"drill_rec": None, "drill_var": None}


def set_cubes(drill_rec, drill_var):
    """The reference reads module globals; set them once after ``Inversion.cubing``."""
    _state["drill_rec"] = np.ascontiguousarray(drill_rec, dtype=float)
    _state["drill_var"] = np.ascontiguousarray(drill_var, dtype=float)


def _cubes(drill_rec, drill_var):
    rec = _state["drill_rec"] if drill_rec is None else drill_rec
    var = _state["drill_var"] if drill_var is None else drill_var
    if rec is None or var is None:
        raise RuntimeError("no cubes: call acquisition.set_cubes(drill_rec, drill_var) or pass them explicitly")
    return rec, var


def sweep_vertical(drill_rec=None, drill_var=None, costs=None, kappa=None, beta=None, top=10):
    """Utility of a vertical drillhole at EVERY voxel column (``futility_vertical`` for all (xd, yd) at once).

    Returns ``(utility, proposals)``: ``utility[a, b]`` (``-inf`` on the border columns the reference excludes) and the
    ``top`` best ``(a, b, utility)`` rows sorted by decreasing utility."""
    rec, var = _cubes(drill_rec, drill_var)
    kappa = _cfg.kappa if kappa is None else kappa
    beta = _cfg.beta if beta is None else beta
    util = _lib.default_context().acquisition_vertical(rec, var, kappa, beta, costs)
    order = np.argsort(-util, axis=None, kind="stable")[:top]
    a, b = np.unravel_index(order, util.shape)
    props = np.column_stack([a, b, util[a, b]])
    return util, props[np.isfinite(props[:, 2])]


def futility_vertical(params, costs=None, drill_rec=None, drill_var=None, kappa=None, beta=None):
    """``run_geobo.py:175-200``: minus the utility of a vertical drillhole at voxel column ``round(params)``."""
    params = np.asarray(params, dtype=float)
    rec, var = _cubes(drill_rec, drill_var)
    if not np.isfinite(params).all():
        return np.inf
    xd, yd = int(np.round(params[0])), int(np.round(params[1]))
    if not (0 < xd < rec.shape[0] - 1 and 0 < yd < rec.shape[1] - 1):
        return np.inf
    util, _ = sweep_vertical(rec, var, costs, kappa, beta, top=1)
    return -util[xd, yd]


def futility_drill(params, costs=None, drill_rec=None, drill_var=None, kappa=None, beta=None):
    """``run_geobo.py:203-235`` for one ``[x0, y0, azimuth, dip]`` (scalar returned) or a batch ``(n, 4)`` (array):
    minus the utility of a drillcore of length ``zLcube`` starting at the surface."""
    rec, var = _cubes(drill_rec, drill_var)
    kappa = _cfg.kappa if kappa is None else kappa
    beta = _cfg.beta if beta is None else beta
    p = np.asarray(params, dtype=float)
    out = _lib.default_context().acquisition_drill(rec, var, (_cfg.xvoxsize, _cfg.yvoxsize, _cfg.zvoxsize), _cfg.zmax, _cfg.zLcube,
                                                   p.reshape(-1, 4), kappa, beta, costs)
    return -out[0] if p.ndim == 1 else -out
