"""Drill-data helpers with the surface of ``geobo/utils.py`` / ``geobo/run_geobo.py``.

``align_drill(coord, data)`` is the reference's voxelisation of drill-core samples (``run_geobo.py:132-159``;
``utils.align_drill2``, ``utils.py:55-83``, is the same loop): every voxel takes the ``nanmean`` of the samples inside
a window of one voxel size to either side of its centre -- an O(voxels x samples) Python loop in the reference, one
CUDA kernel here (``csrc/drill.cu``).  The reference reads the voxel-centre arrays ``xxx, yyy, zzz`` and the voxel
sizes from module globals; here they come from the loaded settings unless passed explicitly.
"""
import numpy as np

from . import _lib
from . import config_loader as _cfg


def _centres(xxx, yyy, zzz):
    if xxx is None or yyy is None or zzz is None:
        # run_geobo.py:399-403: the flat voxel order of create_cubegeometry viewed as (xNcube, yNcube, zNcube)
        from .inversion import Inversion
        vp = Inversion().create_cubegeometry()
        shape = (_cfg.xNcube, _cfg.yNcube, _cfg.zNcube)
        return vp, shape
    xxx = np.asarray(xxx, dtype=float)
    return np.vstack([xxx.ravel(), np.asarray(yyy, dtype=float).ravel(), np.asarray(zzz, dtype=float).ravel()]), xxx.shape


def align_drill(coord, data, xxx=None, yyy=None, zzz=None, voxelsize=None, ctx=None):
    """Drill-core samples in model-cube shape (``run_geobo.py:132-159``).

    coord: (N_drill, 3) sample coordinates, data: (N_drill,) values.  Returns an array shaped like ``xxx`` holding, per
    voxel, the mean of the non-NaN samples with ``centre - size <= coordinate < centre + size`` on every axis, 0 where
    there is none (or the mean is not finite)."""
    vp, shape = _centres(xxx, yyy, zzz)
    vs = (_cfg.xvoxsize, _cfg.yvoxsize, _cfg.zvoxsize) if voxelsize is None else voxelsize
    coord = np.asarray(coord, dtype=float).reshape(-1, 3)
    ctx = ctx or _lib.default_context()
    return ctx.align_drill(vp, coord, np.asarray(data, dtype=float), vs).reshape(shape)


align_drill2 = align_drill      # utils.py:55-83 differs only in how it tests for an empty selection


def downsample_survey(grav_img, mag_img):
    """The survey part of ``read_surveydata`` after the GeoTIFF read (``run_geobo.py:53-65``): both survey maps resampled to one
    value per voxel column with ``scipy.ndimage.zoom`` (factor ``xNcube / image width``, the reference's call with its defaults:
    cubic spline, prefiltered) and the sensor coordinates over the voxel-column centres at height ``zmax + zoff``.

    Host code on purpose: O(Ns) work on two 2-D images, done once per settings file; the identical library call keeps the result
    bit-identical to the reference's.  Returns ``(grav.flatten(), mag.flatten(), sensor_locations)`` -- the first, second and
    fourth argument of ``Inversion.cubing``."""
    from scipy.ndimage import zoom
    from .synth import sensor_grid
    out = []
    for img in (np.asarray(grav_img), np.asarray(mag_img)):
        img2 = zoom(img, _cfg.xNcube * 1. / img.shape[1])
        if img2.shape != (_cfg.yNcube, _cfg.xNcube):
            raise AssertionError("zoomed survey has shape %r, expected (yNcube, xNcube) = %r" % (img2.shape, (_cfg.yNcube, _cfg.xNcube)))
        out.append(img2.flatten())
    return out[0], out[1], sensor_grid()
