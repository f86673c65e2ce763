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

_state = {"drill_rec": None, "drill_var": None, "util_key": None, "util": None}


def set_cubes(drill_rec, drill_var):
    """The reference reads module globals; set them once after ``Inversion.cubing``."""
    _state["drill_rec"] = np.ascontiguousarray(drill_rec, dtype=float)
    _state["drill_var"] = np.ascontiguousarray(drill_var, dtype=float)
    _state["util_key"] = _state["util"] = None


def _utility_map(rec, var, costs, kappa, beta):
    """The whole utility map, computed once per (cubes, costs, kappa, beta) and then indexed: as a drop-in ``shgo`` objective
    ``futility_vertical`` is called point by point, and a full-cube upload + launch + sync per evaluation would be slower than the
    NumPy original.  Cached by object identity of the arrays (``set_cubes`` invalidates it)."""
    key = (id(rec), id(var), id(costs), float(kappa), float(beta), rec.shape)
    if _state["util_key"] != key:
        _state["util"] = _lib.default_context().acquisition_vertical(rec, var, kappa, beta, costs)
        _state["util_key"] = key
        _state["util_refs"] = (rec, var, costs)          # keeps the ids alive while the cache entry exists
    return _state["util"]


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
    kappa = _cfg.kappa if kappa is None else kappa
    beta = _cfg.beta if beta is None else beta
    return -_utility_map(rec, var, costs, kappa, beta)[xd, yd]


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


# ------------------------------------------------------------------------------------------------ proposals (run_geobo.py:246-362)
def sweep_drill(drill_rec=None, drill_var=None, costs=None, kappa=None, beta=None, n_start=None, n_azimuth=36, n_dip=7, top=10):
    """Utility of a non-vertical drillcore for a full factorial grid over the reference's search box
    (``run_geobo.py:322``: start point one voxel inside the cube, azimuth 0..360, dip 30..90 degrees).

    ``n_start``: grid points per start coordinate (default: one per voxel column).  Returns ``(candidates, utility,
    proposals)``: all ``(x0, y0, azimuth, dip)`` rows, their utilities, and the ``top`` best rows with the utility appended."""
    rec, var = _cubes(drill_rec, drill_var)
    n0 = int(n_start[0]) if n_start is not None else max(2, int(round(_cfg.yLcube / _cfg.yvoxsize)) - 1)
    n1 = int(n_start[1]) if n_start is not None else max(2, int(round(_cfg.xLcube / _cfg.xvoxsize)) - 1)
    # the reference's bounds: params[0] in (yvoxsize, yLcube - yvoxsize), params[1] in (xvoxsize, xLcube - xvoxsize)
    p0 = np.linspace(_cfg.yvoxsize, _cfg.yLcube - _cfg.yvoxsize, n0)
    p1 = np.linspace(_cfg.xvoxsize, _cfg.xLcube - _cfg.xvoxsize, n1)
    az = np.linspace(0., 360., int(n_azimuth), endpoint=False)
    dip = np.linspace(30., 90., int(n_dip))
    cand = np.stack(np.meshgrid(p0, p1, az, dip, indexing="ij"), axis=-1).reshape(-1, 4)
    util = -futility_drill(cand, costs, rec, var, kappa, beta)
    order = np.argsort(-util, kind="stable")[:top]
    return cand, util, np.column_stack([cand[order], util[order]])


def _write_csv(path, header, rows):
    with open(path, "w") as f:
        f.write(",".join(header) + "\n")
        for row in rows:
            f.write(",".join(repr(float(v)) for v in row) + "\n")


def bayesopt_vert(drillcoord=None, top=10, costs=None, drill_rec=None, drill_var=None):
    """New vertical drillhole proposals (``run_geobo.py:246-305``): the reference runs ``shgo`` on ``futility_vertical`` and
    lists its local minima; the objective only depends on the voxel column, so here every column is evaluated and the
    ``top`` best are listed.  Returns rows ``(NORTHING, EASTING, BO_GAIN)`` in the reference's units and rounding and
    writes ``newdrill_proposals_vertical.csv`` under ``outpath`` when the settings have one."""
    print('Calculating propoals list of new vertical drillholes...')
    _, props = sweep_vertical(drill_rec, drill_var, costs, top=top)
    rows = np.column_stack([np.round(props[:, 0]) * _cfg.yvoxsize + _cfg.ymin + 0.5 * _cfg.yvoxsize,
                            np.round(props[:, 1]) * _cfg.xvoxsize + _cfg.xmin + 0.5 * _cfg.xvoxsize,
                            np.round(props[:, 2], 4)])
    if len(rows):
        print('New vertical Drillcore Proposal:')
        print('EASTING [meters]: ', rows[0, 1])
        print('NORTHING [meters]: ', rows[0, 0])
    outpath = getattr(_cfg, "outpath", None)
    if outpath:
        import os
        os.makedirs(outpath, exist_ok=True)
        _write_csv(os.path.join(outpath, 'newdrill_proposals_vertical.csv'), ['NORTHING', 'EASTING', 'BO_GAIN'], rows)
    return rows


def bayesopt_nonvert(drillcoord=None, top=10, costs=None, drill_rec=None, drill_var=None, **grid):
    """New non-vertical drillcore proposals (``run_geobo.py:308-362``) from the factorial sweep of ``sweep_drill``.
    Returns rows ``(NORTHING, EASTING, AZIMUTH, DIP, BO_GAIN)`` and writes ``newdrill_proposals_non-vertical.csv``."""
    _, _, props = sweep_drill(drill_rec, drill_var, costs, top=top, **grid)
    rows = np.column_stack([np.round(np.round(props[:, 0], 2) + _cfg.ymin, 1), np.round(np.round(props[:, 1], 2) + _cfg.xmin, 1),
                            np.round(props[:, 2], 2), np.round(props[:, 3], 2), np.round(props[:, 4], 4)])
    if len(rows):
        print('New non-vertical Drillcore Proposal:')
        print('EASTING [meters]: ', rows[0, 1])
        print('NORTHING [meters]: ', rows[0, 0])
        print('Azimuth Angle [degree]: ', rows[0, 2])
        print('Dip Angle [degree]: ', rows[0, 3])
    outpath = getattr(_cfg, "outpath", None)
    if outpath:
        import os
        os.makedirs(outpath, exist_ok=True)
        _write_csv(os.path.join(outpath, 'newdrill_proposals_non-vertical.csv'), ['NORTHING', 'EASTING', 'AZIMUTH', 'DIP', 'BO_GAIN'], rows)
    return rows
