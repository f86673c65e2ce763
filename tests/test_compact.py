"""Compact-support form of the 'sparse' covariance blocks -- settings key ``structure: compact``.

CPU: (1) the algebra -- ``oracle/stencil.py`` (tap sum over the support window, taps from the oracle's own ``cov_block``)
against the oracle's dense ``pt_panel`` / ``create_cov``; (2) the device source -- ``csrc/stencil.cuh`` compiled for the host
and driven with the launch geometry of ``csrc/stencil.cu`` (ragged strips, voxel-column shards cutting through a plane,
accumulation over data blocks, windows wider than the cube) against the dense oracle.
GPU (``-m gpu``): ``Inversion.cubing`` with ``structure: compact`` against the oracle, the reference's committed VTK goldens of
both examples (sparse kernel) for both precisions, the loud refusal for other kernels.
"""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import CUBES, HARNESS_CXX, ROOT, load_golden, normwise_err
from oracle import numpy_oracle as o
from oracle import stencil as st

KID = {"sparse": 0, "exp": 1, "matern32": 2}


def _cfg(shape, **kw):
    cfg = json.loads(str(load_golden("sens_8x6x5.npz")["cfg"]))
    cfg.update(xNcube=shape[0], yNcube=shape[1], zNcube=shape[2], kernelfunc="sparse", **kw)
    return cfg


def _dense_pt(c, params, w, amp, A_list, didx):
    pts = o.grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    return o.pt_panel(c, params, w, amp, A_list, didx, pts, np.arange(pts.shape[0]))


def _case(shape, seed, nrows=None, w=(0.9, 0.4, 0.7), amp=1.3, mult=2.0):
    c = o.make_config(_cfg(shape))
    N = shape[0] * shape[1] * shape[2]
    rng = np.random.default_rng(seed)
    nrows = shape[0] * shape[1] if nrows is None else nrows
    A = [rng.standard_normal((nrows, N)), rng.standard_normal((nrows, N))]
    params = o.dedup_lengths(mult * np.asarray([c.xvoxsize] * 3))          # Q1: [L, 1.02 L, L]
    return c, N, A, params, np.asarray(w, dtype=float), amp


# ------------------------------------------------------------------------------------------------ (1) the algebra
@pytest.mark.parametrize("shape,mult", [((6, 5, 7), 2.0), ((9, 3, 4), 1.0), ((4, 4, 4), 40.0), ((5, 6, 8), 0.3)])
def test_oracle_tap_sum_equals_dense_projection(shape, mult):
    """mult = 40: the support covers the whole cube (window clipped to n - 1); mult = 0.3: only the zero offset survives."""
    c, N, A, params, w, amp = _case(shape, 1, mult=mult)
    didx = np.array([1, N // 2, N - 2])
    dense = _dense_pt(c, params, w, amp, A, didx)
    got = st.pt_compact(c, params, w, amp, A, didx)
    assert np.abs(got - dense).max() <= 1e-13 * np.abs(dense).max()
    ry, rx, rz = st.window(c, params)
    assert 0 <= ry <= c.yNcube - 1 and 0 <= rx <= c.xNcube - 1 and 0 <= rz <= c.zNcube - 1


def test_oracle_window_contains_the_whole_support():
    """Every non-zero of the dense blocks lies inside the window (so culling the other offsets drops exact zeros only)."""
    c, N, A, params, w, amp = _case((7, 6, 9), 2, mult=2.5)
    ry, rx, rz = st.window(c, params)
    pts = o.grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    K = o.create_cov(o.sqdist(pts), params.copy(), w, "sparse")
    idx = np.arange(N)
    iy, ix, iz = idx // (c.xNcube * c.zNcube), (idx // c.zNcube) % c.xNcube, idx % c.zNcube
    outside = ((np.abs(iy[:, None] - iy[None, :]) > ry) | (np.abs(ix[:, None] - ix[None, :]) > rx) | (np.abs(iz[:, None] - iz[None, :]) > rz))
    assert outside.any() and not K[np.tile(outside, (3, 3))].any()
    assert (K != 0).mean() < 0.5                                                     # the blocks really are sparse here


def test_oracle_tap_matvec_equals_create_cov_matvec():
    c, N, A, params, w, amp = _case((5, 4, 6), 3)
    pts = o.grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    K = amp * o.create_cov(o.sqdist(pts), params.copy(), w, "sparse")
    W = np.random.default_rng(3).standard_normal((3, N))
    want = (K @ W.ravel()).reshape(3, N)
    got = st.kw_compact(c, params, w, amp, W)
    assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()


# ------------------------------------------------------------------------------------------------ (2) the device source on the host
@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = tmp_path_factory.mktemp("stencil_host") / "stencil_host.so"
    src = os.path.join(ROOT, "tests", "host_harness", "stencil_host.cpp")
    subprocess.run(HARNESS_CXX + [src, "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    P, L, D, I = ctypes.c_void_p, ctypes.c_long, ctypes.c_double, ctypes.c_int
    lib.stencil_host_apply.argtypes = [I, P, P, D, P, P, I, P, L, L, L, L, P, L, L, I, P]
    lib.stencil_host_apply.restype = None

    def apply(c, params, w, amp, blk0, A, c0, c1, out, ldo, r_stride_out, accumulate, kernel="sparse"):
        p = lambda a: a.ctypes.data_as(P)                                            # noqa: E731
        l = np.ascontiguousarray(params, dtype=float)
        ww = np.ascontiguousarray(w, dtype=float)
        ncube = np.array([c.xNcube, c.yNcube, c.zNcube], dtype=np.int64)
        vox = np.array([c.xvoxsize, c.yvoxsize, c.zvoxsize], dtype=float)
        A = np.ascontiguousarray(A, dtype=float)
        win = np.zeros(3, dtype=np.int32)
        assert out.flags.c_contiguous and out.dtype == np.float64
        lib.stencil_host_apply(KID[kernel], p(l), p(ww), amp, p(ncube), p(vox), blk0, p(A), A.shape[1], A.shape[0], c0, c1, p(out), ldo,
                               r_stride_out, accumulate, p(win))
        return tuple(int(v) for v in win)
    return apply


def _host_projection(host, c, params, w, amp, A, c0, c1):
    Ns, ncol = A[0].shape[0], c1 - c0
    ncp = -(-ncol // 32) * 32
    Pt = np.full((2 * Ns, 3 * ncp), np.nan)
    for cb in range(2):
        win = host(c, params, w, amp, cb * 3, A[cb], c0, c1, Pt[cb * Ns:], 3 * ncp, ncp, 0)
    return Pt.reshape(2 * Ns, 3, ncp), win


@pytest.mark.parametrize("shape,shard,mult", [
    ((6, 5, 7), None, 2.0),            # zN = 7: two strips of 4, the second ragged
    ((6, 5, 7), (16, 160), 2.0),       # shard starts and ends inside a plane (42 voxels per y-row)
    ((4, 4, 4), None, 40.0),           # support wider than the cube: window clipped to n - 1 on every axis
    ((5, 6, 8), (32, 208), 0.3),       # only the zero offset inside the support
    ((70, 2, 16), None, 1.5),          # 280 work items per plane: two tiles of 256 threads, the second partly empty
    ((1, 7, 1), None, 3.0),            # degenerate axes
])
def test_host_compiled_kernel_vs_dense_oracle(host, shape, shard, mult):
    c, N, A, params, w, amp = _case(shape, 5, nrows=5, mult=mult)
    c0, c1 = shard if shard else (0, N)
    dense = _dense_pt(c, params, w, amp, A, np.zeros(0, dtype=int))
    Pt, win = _host_projection(host, c, params, w, amp, A, c0, c1)
    assert win == st.window(c, params)                                               # the library's window = the oracle's
    got = Pt[:, :, :c1 - c0]
    assert np.isfinite(got).all() and np.isnan(Pt[:, :, c1 - c0:]).all()             # exactly the shard's columns were written
    assert np.abs(got - dense[:, :, c0:c1]).max() <= 1e-13 * np.abs(dense).max()


def test_host_compiled_matvec_accumulates_over_data_blocks(host):
    c, N, A, params, w, amp = _case((6, 5, 4), 6)
    W = np.random.default_rng(7).standard_normal((3, N))
    c0, c1, ncp = 16, 112, 96
    z = np.full((3, ncp), np.nan)
    for cb in range(3):
        host(c, params, w, amp, cb * 3, W[cb][None, :], c0, c1, z, 0, ncp, int(cb > 0))
    want = st.kw_compact(c, params, w, amp, W)[:, c0:c1]
    assert np.abs(z - want).max() <= 1e-13 * np.abs(want).max()


def test_reference_example_geometry_window():
    """examples/settings_example1.yaml (25 x 16 x 16, gp_lengthscale 2 -> scales [L, 1.02 L, L] after Q1): the window is
    5 x 5 x 9 offsets = 225 taps instead of N = 6400 per output."""
    cfg = json.loads(str(load_golden("example1.npz")["cfg"]))
    c = o.make_config(cfg)
    params = o.dedup_lengths(c.gp_lengthscale * np.asarray([c.xvoxsize] * 3))
    assert st.window(c, params) == (2, 2, 4)


# ------------------------------------------------------------------------------------------------ GPU
def _gpu_cubing(cfg, f, gl=None):
    from geobo_b200 import _lib, config_loader, inversion
    config_loader.load_settings(cfg, make_outpath=False)
    inv = inversion.Inversion()
    inv.create_cubegeometry()
    if gl is not None:
        inv.gp_length = np.array(gl, dtype=float)
    try:
        out = inv.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    finally:
        if inv._problem is not None:
            inv._problem.close()
            inv._problem = None
        _lib.default_context().release_cache()
    return inv, out


@pytest.mark.gpu
@pytest.mark.parametrize("shape,nd,prec", [((9, 2, 1), 2, "fp64"), ((17, 13, 9), 0, "fp64"), ((6, 5, 7), 4, "fp64"),
                                           ((9, 8, 16), 4, "int8x5"), ((20, 16, 16), 0, "int8x6")])
def test_gpu_compact_cubing_vs_oracle(shape, nd, prec):
    from test_gpu_parity import synthetic_inputs
    cfg = _cfg(shape, structure="compact", precision=prec)
    c = o.make_config(cfg)
    f = synthetic_inputs(c, nd)
    with np.errstate(all="ignore"):
        ref, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    inv, out = _gpu_cubing(cfg, f)
    tol = 1e-7 if prec in ("fp64", "int8x6") else 1e-6
    for n, a, r in zip(CUBES, out, ref):
        assert normwise_err(a, r) < tol, n
    assert abs(inv.logl - ex["logl"]) < (1e-7 if prec == "fp64" else 1e-4) * abs(ex["logl"])


@pytest.mark.gpu
@pytest.mark.parametrize("which,prec", [("1", "fp64"), ("2", "fp64"), ("1", "int8x5")])
def test_gpu_compact_examples_vs_committed_vtk_goldens(which, prec):
    """The reference's own golden cubes (25 x 16 x 16, sparse kernel) through the tap-sum path."""
    f = load_golden("example%s.npz" % which)
    cfg = dict(json.loads(str(f["cfg"])), structure="compact", precision=prec)
    inv, out = _gpu_cubing(cfg, f)
    for n, a in zip(CUBES, out):
        assert normwise_err(a, f["gold_" + n]) < (1e-7 if prec == "fp64" else 1e-6), n
    assert abs(inv.logl - float(f["logl"])) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("kf", ["exp", "matern32"])
def test_gpu_compact_refuses_kernels_without_compact_support(kf):
    from geobo_b200 import _lib
    from test_gpu_parity import synthetic_inputs
    cfg = dict(_cfg((5, 4, 3), structure="compact"), kernelfunc=kf)
    f = synthetic_inputs(o.make_config(cfg), 0)
    with pytest.raises(_lib.GeoboB200Error, match="compact"):
        _gpu_cubing(cfg, f)


# ------------------------------------------------------------------------------------------------ CPU: the reference's own golden vectors
@pytest.mark.parametrize("which,structure", [("1", "compact"), ("2", "compact"), ("1", "fft")])
def test_structured_oracle_reproduces_the_committed_vtk_goldens(monkeypatch, which, structure):
    """The reference's committed result cubes (examples/results/*/cube_*.vtk, 25 x 16 x 16, sparse kernel) from the oracle's lean
    pipeline with Pt = Asens3 . kcov taken from the structured restatement instead of the dense panels: the tap-sum / FFT forms
    are pinned against the reference's own golden vectors, not only against the dense oracle."""
    from oracle import fftconv as fc
    f = load_golden("example%s.npz" % which)
    cfg = json.loads(str(f["cfg"]))
    c = o.make_config(cfg)
    cache = {}
    dense_panel = o.pt_panel

    def structured_panel(c_, params, w, amp, A_list, didx, pts, cols, **kw):
        if cache.get("building"):                  # the drill rows inside pt_compact / pt_fft stay dense gathers
            return dense_panel(c_, params, w, amp, A_list, didx, pts, cols, **kw)
        if "pt" not in cache:
            cache["building"] = True
            cache["pt"] = (st.pt_compact(c_, params, w, amp, A_list, didx) if structure == "compact"
                           else fc.pt_fft(c_, params, w, amp, c_.kernelfunc, A_list, didx))
            cache["building"] = False
        return cache["pt"][:, :, cols]

    monkeypatch.setattr(o, "pt_panel", structured_panel)
    with np.errstate(all="ignore"):
        cubes, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    assert "pt" in cache
    for n, a in zip(CUBES, cubes):
        assert normwise_err(a, f["gold_" + n]) < 2e-7, n          # the VTK files come from an older BLAS stack (<= 4e-8)
    assert abs(ex["logl"] - float(f["logl"])) < 1e-6


def _pipeline_with_projection(monkeypatch, c, f, project, gp_length=None):
    """The oracle's lean pipeline with the two survey blocks of Pt supplied by ``project(params, w, amp, A_list)`` -> (2 Ns, 3, N);
    the drill rows stay the oracle's dense gathers."""
    cache = {}
    dense_panel = o.pt_panel

    def panel(c_, params, w, amp, A_list, didx, pts, cols, **kw):
        if "pt" not in cache:
            Ns = A_list[0].shape[0]
            N = A_list[0].shape[1]
            full = np.zeros((2 * Ns + didx.size, 3, N))
            full[:2 * Ns] = project(params, w, amp, A_list)
            if didx.size:
                zero_rows = [np.zeros((1, N)), np.zeros((1, N))]           # only the drill rows of the dense panel are wanted
                full[2 * Ns:] = dense_panel(c_, params, w, amp, zero_rows, didx, pts, np.arange(N))[2:]
            cache["pt"] = full
        return cache["pt"][:, :, cols]

    monkeypatch.setattr(o, "pt_panel", panel)
    with np.errstate(all="ignore"):
        kw = {} if gp_length is None else {"gp_length": np.array(gp_length, dtype=float)}
        cubes, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"], **kw)
    assert "pt" in cache
    return cubes, ex


def test_host_compiled_kernel_reproduces_the_committed_vtk_goldens(host, monkeypatch):
    """Example 1 of the reference (25 x 16 x 16, sparse kernel, M = 1056) with Pt computed by the DEVICE SOURCE of the tap-sum kernel
    (csrc/stencil.cuh compiled for the host, tables from csrc/formulas.cuh): the committed VTK cubes are reproduced to 2e-7."""
    f = load_golden("example1.npz")
    c = o.make_config(json.loads(str(f["cfg"])))

    def project(params, w, amp, A_list):
        N = A_list[0].shape[1]
        Pt, win = _host_projection(host, c, params, w, amp, A_list, 0, N)
        assert win == (2, 2, 4)
        return Pt[:, :, :N]

    cubes, ex = _pipeline_with_projection(monkeypatch, c, f, project)
    for n, a in zip(CUBES, cubes):
        assert normwise_err(a, f["gold_" + n]) < 2e-7, n
    assert abs(ex["logl"] - float(f["logl"])) < 1e-6
