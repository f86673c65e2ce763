"""align_drill (geobo/run_geobo.py:132-159): drill-core samples -> voxel cube.

Fixture ``tests/golden/align_drill.npz`` comes from the unmodified reference function (``make_simdata_golden.py``):
synthetic holes on a 12x10x8 cube with samples exactly on window edges, NaN and inf values, and the reference's
committed ``simdrill_cylinders.csv`` on the example-1 cube.  CPU: the oracle restatement is bit-exact against it, and the
per-voxel source the CUDA kernel is built from (``csrc/drill.cuh``), compiled for the host, selects the same samples.
GPU: ``geobo_b200.utils.align_drill`` through the C ABI.  Tolerance 1e-13 relative on the means (NumPy sums the selected
samples pairwise, the kernel left to right); the set of non-zero voxels must be identical.
"""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import HARNESS_CXX, ROOT, load_golden
from oracle import numpy_oracle as o

RTOL = 1e-13


def _centres(cfg):
    c = o.make_config(cfg)
    _, vp = o.cube_geometry(c)
    shape = (c.xNcube, c.yNcube, c.zNcube)                          # run_geobo.py:399-403
    return c, np.asarray(vp), shape


@pytest.mark.parametrize("which", ["small", "example"])
def test_oracle_align_drill_is_bit_exact_against_the_reference(which):
    f = load_golden("align_drill.npz")
    c, vp, shape = _centres(json.loads(str(f["cfg_" + which])))
    got = o.align_drill(f["coord_" + which], f["data_" + which], vp[0].reshape(shape), vp[1].reshape(shape), vp[2].reshape(shape),
                        (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    assert np.array_equal(got, f["res_" + which])
    assert o.align_drill(np.zeros((0, 3)), np.zeros(0), vp[0].reshape(shape), vp[1].reshape(shape), vp[2].reshape(shape),
                         (c.xvoxsize, c.yvoxsize, c.zvoxsize)).any() == False  # noqa: E712


def test_example_fixture_is_the_drilldata0_of_the_example_run():
    """The committed drill csv voxelised by the reference = the ``drilldata0`` its driver passed to ``cubing`` (example 1)."""
    f, ex = load_golden("align_drill.npz"), load_golden("example1.npz")
    assert np.array_equal(f["res_example"], ex["drilldata0"])


@pytest.fixture(scope="module")
def host_kernel(tmp_path_factory):
    """csrc/drill.cuh compiled for the host (the same drill_accumulate / drill_finish the CUDA kernel calls)."""
    so = tmp_path_factory.mktemp("drill_host") / "drill_host.so"
    src = os.path.join(ROOT, "tests", "host_harness", "drill_host.cpp")
    subprocess.run(HARNESS_CXX + [src, "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    P = ctypes.c_void_p
    lib.drill_host.argtypes = [P, ctypes.c_long, P, P, ctypes.c_long, P, ctypes.c_long, P]
    lib.drill_host.restype = None

    def run(vp, coord, data, vs, tile=1024):
        vp = np.ascontiguousarray(vp, dtype=float)
        coord = np.ascontiguousarray(coord, dtype=float).reshape(-1, 3)
        data = np.ascontiguousarray(data, dtype=float)
        vs = np.ascontiguousarray(vs, dtype=float)
        out = np.empty(vp.shape[1])
        lib.drill_host(vp.ctypes.data, vp.shape[1], coord.ctypes.data, data.ctypes.data, coord.shape[0], vs.ctypes.data, tile, out.ctypes.data)
        return out
    return run


def _check(got, ref):
    ref = np.asarray(ref).ravel()
    assert np.array_equal(got != 0, ref != 0)                       # same voxels own samples (window edges, NaN, inf rules)
    assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max()


@pytest.mark.parametrize("which", ["small", "example"])
def test_kernel_source_on_the_host_matches_the_reference(host_kernel, which):
    f = load_golden("align_drill.npz")
    c, vp, shape = _centres(json.loads(str(f["cfg_" + which])))
    vs = (c.xvoxsize, c.yvoxsize, c.zvoxsize)
    got = host_kernel(vp, f["coord_" + which], f["data_" + which], vs)
    _check(got, f["res_" + which])
    assert np.array_equal(got, host_kernel(vp, f["coord_" + which], f["data_" + which], vs, tile=7))   # tiling does not change the order
    assert not host_kernel(vp, np.zeros((0, 3)), np.zeros(0), vs).any()


def test_window_edges_nan_and_inf_rules(host_kernel):
    """One voxel at the origin, window [-1, 1) per axis."""
    vp = np.zeros((3, 1))
    vs = (1.0, 1.0, 1.0)
    cases = [([[-1.0, 0, 0]], [5.0], 5.0),                          # lower edge belongs to the window
             ([[1.0, 0, 0]], [5.0], 0.0),                           # upper edge does not
             ([[0, 0, np.nextafter(1.0, 0)]], [5.0], 5.0),
             ([[0, 0, 0], [0.5, 0.5, -0.5]], [np.nan, 3.0], 3.0),   # NaN skipped by nanmean
             ([[0, 0, 0]], [np.nan], 0.0),                          # only NaN -> mean undefined -> 0
             ([[0, 0, 0], [0, 0, 0]], [np.inf, 1.0], 0.0),          # infinite mean -> 0
             ([[0, 0, 0], [0, 0, 0]], [np.inf, -np.inf], 0.0),
             ([[0, 0, 0], [0.1, 0, 0], [5, 0, 0]], [1.0, 2.0, 100.0], 1.5)]
    for coord, data, want in cases:
        got = host_kernel(vp, np.array(coord, dtype=float), np.array(data), vs)[0]
        ref = o.align_drill(np.array(coord, dtype=float), np.array(data), np.zeros((1, 1, 1)), np.zeros((1, 1, 1)), np.zeros((1, 1, 1)), vs)[0, 0, 0]
        assert got == want == ref, (coord, data, got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["small", "example"])
def test_gpu_align_drill_vs_reference_fixture(which):
    from geobo_b200 import config_loader, utils
    f = load_golden("align_drill.npz")
    cfg = json.loads(str(f["cfg_" + which]))
    config_loader.load_settings(cfg, make_outpath=False)
    got = utils.align_drill(f["coord_" + which], f["data_" + which])                     # geometry from the settings
    assert got.shape == f["res_" + which].shape
    _check(got.ravel(), f["res_" + which])
    c, vp, shape = _centres(cfg)
    got2 = utils.align_drill2(f["coord_" + which], f["data_" + which], vp[0].reshape(shape), vp[1].reshape(shape), vp[2].reshape(shape))
    assert np.array_equal(got, got2)                                                     # explicit centre arrays
    assert not utils.align_drill(np.zeros((0, 3)), np.zeros(0)).any()                    # no samples -> zeros


@pytest.mark.gpu
def test_gpu_align_drill_many_samples_vs_oracle():
    """More samples than one shared-memory tile (1024) and a voxel count that is not a multiple of the block size."""
    from geobo_b200 import config_loader, utils
    cfg = dict(json.loads(str(load_golden("align_drill.npz")["cfg_small"])), xNcube=13, yNcube=9, zNcube=11)   # 9, not 7: the reference's np.arange centres overshoot at yNcube=7 (8 centres)
    config_loader.load_settings(cfg, make_outpath=False)
    c, vp, shape = _centres(cfg)
    assert vp.shape[1] == int(np.prod(shape))                       # geometry valid for the reference's centre formula
    rng = np.random.default_rng(5)
    n = 2500
    coord = np.column_stack([rng.uniform(-50, c.xLcube + 50, n), rng.uniform(-50, c.yLcube + 50, n), -rng.uniform(-20, c.zLcube + 20, n)])
    data = rng.standard_normal(n)
    data[rng.choice(n, 40, replace=False)] = np.nan
    ref = o.align_drill(coord, data, vp[0].reshape(shape), vp[1].reshape(shape), vp[2].reshape(shape), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    got = utils.align_drill(coord, data)
    assert np.array_equal(got != 0, ref != 0)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
