"""Block-Toeplitz (FFT) form of the covariance blocks -- settings key ``structure: fft``, all kernels.

CPU: (1) the algebra -- ``oracle/fftconv.py`` (circulant embedding with numpy.fft, offsets from the oracle's own
``cov_block``) against the oracle's dense ``pt_panel`` / ``create_cov`` for the three kernel families; (2) the device source
-- ``csrc/fftconv.cuh`` compiled for the host and driven with the pass sequence, chunking and scratch carve-up of
``csrc/fftconv.cu`` (radix-2 passes in "shared memory", phase by phase) against the dense oracle: ragged shapes, odd row
counts (the unpaired last row), shards cutting through a plane, accumulation over data blocks.
GPU (``-m gpu``): ``Inversion.cubing`` with ``structure: fft`` against the oracle for the three kernels and both precisions,
the reference's committed VTK goldens, BASELINE config 3 (Matern-3/2, 64x64x32) at full size against the CPU fixture.
"""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import CUBES, GOLDEN, HARNESS_CXX, ROOT, load_golden, normwise_err
from oracle import fftconv as fc
from oracle import numpy_oracle as o

KID = {"sparse": 0, "exp": 1, "matern32": 2}
TOL = 1e-12          # FFT rounding: eps * log2(Py Px Pz) relative to the largest entry


def _cfg(shape, kernel, **kw):
    cfg = json.loads(str(load_golden("sens_8x6x5.npz")["cfg"]))
    cfg.update(xNcube=shape[0], yNcube=shape[1], zNcube=shape[2], kernelfunc=kernel, **kw)
    return cfg


def _dense_pt(c, params, w, amp, A_list, didx):
    pts = o.grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    return o.pt_panel(c, params, w, amp, A_list, didx, pts, np.arange(pts.shape[0]))


def _case(shape, kernel, seed, nrows=None, w=(0.9, 0.4, 0.7), amp=1.3, mult=2.0):
    c = o.make_config(_cfg(shape, kernel))
    N = shape[0] * shape[1] * shape[2]
    rng = np.random.default_rng(seed)
    nrows = shape[0] * shape[1] if nrows is None else nrows
    A = [rng.standard_normal((nrows, N)), rng.standard_normal((nrows, N))]
    params = mult * c.xvoxsize * np.array([1.0, 1.01, 1.02])                # distinct scales (matern32 is singular otherwise)
    return c, N, A, params, np.asarray(w, dtype=float), amp


# ------------------------------------------------------------------------------------------------ (1) the algebra
@pytest.mark.parametrize("kernel", ["exp", "sparse", "matern32"])
@pytest.mark.parametrize("shape", [(5, 4, 6), (3, 7, 2), (1, 5, 4)])
def test_oracle_fft_projection_equals_dense_projection(kernel, shape):
    c, N, A, params, w, amp = _case(shape, kernel, 1)
    didx = np.array([0, N // 2, N - 1])
    dense = _dense_pt(c, params, w, amp, A, didx)
    got = fc.pt_fft(c, params, w, amp, kernel, A, didx)
    assert np.abs(got - dense).max() <= TOL * np.abs(dense).max()
    assert all(p >= 2 * n - 1 and p & (p - 1) == 0 for p, n in zip(fc.padded(c), (c.yNcube, c.xNcube, c.zNcube)))


@pytest.mark.parametrize("kernel", ["exp", "sparse", "matern32"])
def test_oracle_fft_matvec_equals_create_cov_matvec(kernel):
    c, N, A, params, w, amp = _case((4, 3, 5), kernel, 2)
    pts = o.grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    K = amp * o.create_cov(o.sqdist(pts), params.copy(), w, kernel)
    W = np.random.default_rng(3).standard_normal((3, N))
    want = (K @ W.ravel()).reshape(3, N)
    got = fc.kw_fft(c, params, w, amp, kernel, W)
    assert np.abs(got - want).max() <= TOL * np.abs(want).max()


# ------------------------------------------------------------------------------------------------ (2) the device source on the host
@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = tmp_path_factory.mktemp("fftconv_host") / "fftconv_host.so"
    src = os.path.join(ROOT, "tests", "host_harness", "fftconv_host.cpp")
    subprocess.run(HARNESS_CXX + [src, "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    P, L, D, I = ctypes.c_void_p, ctypes.c_long, ctypes.c_double, ctypes.c_int
    lib.fftconv_host_apply.argtypes = [I, P, P, D, P, P, I, P, L, L, L, L, L, P, L, L, I, P]
    lib.fftconv_host_apply.restype = None
    lib.fftconv_host_set_thread_order.argtypes = [I]

    def apply(c, params, w, amp, blk0, A, c0, c1, B, out, ldo, r_stride_out, accumulate):
        p = lambda a: a.ctypes.data_as(P)                                            # noqa: E731
        l = np.ascontiguousarray(params, dtype=float)
        ww = np.ascontiguousarray(w, dtype=float)
        ncube = np.array([c.xNcube, c.yNcube, c.zNcube], dtype=np.int64)
        vox = np.array([c.xvoxsize, c.yvoxsize, c.zvoxsize], dtype=float)
        A = np.ascontiguousarray(A, dtype=float)
        pad = np.zeros(3, dtype=np.int32)
        assert out.flags.c_contiguous and out.dtype == np.float64
        lib.fftconv_host_apply(KID[c.kernelfunc], p(l), p(ww), amp, p(ncube), p(vox), blk0, p(A), A.shape[1], A.shape[0], c0, c1, B, p(out), ldo,
                               r_stride_out, accumulate, p(pad))
        return tuple(int(v) for v in pad)
    apply.set_thread_order = lib.fftconv_host_set_thread_order
    return apply


def _host_projection(host, c, params, w, amp, A, c0, c1, B):
    Ns, ncol = A[0].shape[0], c1 - c0
    ncp = -(-ncol // 32) * 32
    Pt = np.full((2 * Ns, 3 * ncp), np.nan)
    for cb in range(2):
        pad = host(c, params, w, amp, cb * 3, A[cb], c0, c1, B, Pt[cb * Ns:], 3 * ncp, ncp, 0)
    return Pt.reshape(2 * Ns, 3, ncp), pad


@pytest.mark.parametrize("shape,kernel,shard,nrows,B", [
    ((5, 6, 7), "matern32", None, 5, 8),            # odd row count: the last complex transform carries one real row
    ((5, 6, 7), "sparse", (32, 176), 4, 1),         # shard inside planes (35 voxels per y-row); one row pair per chunk
    ((9, 3, 4), "exp", None, 6, 2),                 # padded lengths 8 x 32 x 8 ... several chunks
    ((1, 5, 1), "matern32", None, 3, 1),            # degenerate axes: P = 1 transforms (no butterfly stage)
    ((3, 2, 20), "matern32", (0, 96), 2, 1),        # Pz = 64: 16 lines x 64 elements per block
])
def test_host_compiled_kernels_vs_dense_oracle(host, shape, kernel, shard, nrows, B):
    c, N, A, params, w, amp = _case(shape, kernel, 5, nrows=nrows)
    c0, c1 = shard if shard else (0, N)
    dense = _dense_pt(c, params, w, amp, A, np.zeros(0, dtype=int))
    Pt, pad = _host_projection(host, c, params, w, amp, A, c0, c1, B)
    assert pad == fc.padded(c)
    got = Pt[:, :, :c1 - c0]
    assert np.isfinite(got).all() and np.isnan(Pt[:, :, c1 - c0:]).all()             # exactly the shard's columns were written
    assert np.abs(got - dense[:, :, c0:c1]).max() <= TOL * np.abs(dense).max()


def test_host_compiled_kernels_do_not_depend_on_the_thread_order_inside_a_phase(host):
    """Between two barriers (load | every butterfly stage | store) the harness may run the threads of a block in any order: a
    stage in which two threads touch the same element would give a different result."""
    c, N, A, params, w, amp = _case((5, 6, 7), "matern32", 9, nrows=3)
    try:
        res = []
        for order in (0, 1, 2):
            host.set_thread_order(order)
            res.append(_host_projection(host, c, params, w, amp, A, 32, 176, 1)[0])
    finally:
        host.set_thread_order(0)
    assert np.array_equal(res[0], res[1], equal_nan=True) and np.array_equal(res[0], res[2], equal_nan=True)


def test_host_compiled_matvec_accumulates_over_data_blocks(host):
    c, N, A, params, w, amp = _case((6, 5, 4), "matern32", 6)
    W = np.random.default_rng(7).standard_normal((3, N))
    c0, c1, ncp = 16, 112, 96
    z = np.full((3, ncp), np.nan)
    for cb in range(3):
        host(c, params, w, amp, cb * 3, W[cb][None, :], c0, c1, 1, z, 0, ncp, int(cb > 0))
    want = fc.kw_fft(c, params, w, amp, "matern32", W)[:, c0:c1]
    assert np.abs(z - want).max() <= TOL * np.abs(want).max()


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
@pytest.mark.parametrize("shape,kernel,nd,prec", [((7, 5, 3), "exp", 3, "fp64"), ((9, 2, 1), "sparse", 2, "fp64"), ((3, 4, 19), "matern32", 5, "fp64"),
                                                  ((17, 13, 9), "sparse", 0, "fp64"), ((6, 5, 32), "matern32", 5, "int8x5"),
                                                  ((11, 4, 48), "exp", 0, "int8x6")])
def test_gpu_fft_cubing_vs_oracle(shape, kernel, nd, prec):
    from test_gpu_parity import synthetic_inputs
    cfg = _cfg(shape, kernel, structure="fft", precision=prec)
    c = o.make_config(cfg)
    f = synthetic_inputs(c, nd)
    gl = c.gp_lengthscale * c.xvoxsize * (np.array([1.0, 1.01, 1.02]) if kernel == "matern32" else np.ones(3))
    with np.errstate(all="ignore"):
        ref, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"], gp_length=gl.copy())
    inv, out = _gpu_cubing(cfg, f, gl=gl.copy())
    tol = 1e-7 if prec in ("fp64", "int8x6") else 1e-6
    for n, a, r in zip(CUBES, out, ref):
        assert normwise_err(a, r) < tol, n
    assert abs(inv.logl - ex["logl"]) < (1e-7 if prec == "fp64" else 1e-4) * abs(ex["logl"])


@pytest.mark.gpu
def test_gpu_fft_example_vs_committed_vtk_goldens():
    f = load_golden("example1.npz")
    cfg = dict(json.loads(str(f["cfg"])), structure="fft")
    inv, out = _gpu_cubing(cfg, f)
    for n, a in zip(CUBES, out):
        assert normwise_err(a, f["gold_" + n]) < 1e-7, n
    assert abs(inv.logl - float(f["logl"])) < 1e-3


@pytest.mark.gpu
def test_gpu_fft_matches_dense_device_path_and_scratch_chunks(monkeypatch):
    from test_gpu_parity import synthetic_inputs
    cfg = _cfg((16, 16, 16), "matern32")
    f = synthetic_inputs(o.make_config(cfg), 50)
    c = o.make_config(cfg)
    gl = c.gp_lengthscale * c.xvoxsize * np.array([1.0, 1.01, 1.02])
    _, dense = _gpu_cubing(dict(cfg, structure="dense"), f, gl=gl.copy())
    _, fft = _gpu_cubing(dict(cfg, structure="fft"), f, gl=gl.copy())
    monkeypatch.setenv("GEOBO_B200_FFT_SCRATCH_MB", "1")
    _, fft_small = _gpu_cubing(dict(cfg, structure="fft"), f, gl=gl.copy())
    for n, a, b, b2 in zip(CUBES, dense, fft, fft_small):
        assert normwise_err(b, a) < 1e-8, n
        assert normwise_err(b2, b) < 1e-12, n


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp64", "int8x5"])
def test_gpu_fft_fullsize_cfg3_vs_cpu_oracle(prec):
    """BASELINE config 3 (64x64x32, Matern-3/2, nd = 50) through structure fft against the oracle's full-size result (1e-5)."""
    if not os.path.exists(os.path.join(GOLDEN, "fullsize_cfg3.npz")):
        pytest.skip("fixture fullsize_cfg3.npz not generated")
    g = load_golden("fullsize_cfg3.npz")
    cfg = dict(json.loads(str(g["cfg"])), precision=prec, structure="fft")
    c = o.make_config(cfg)
    N = c.xNcube * c.yNcube * c.zNcube
    d0 = np.zeros(N)
    d0[g["didx"]] = g["drillvals"]
    d0 = d0.reshape(c.xNcube, c.yNcube, c.zNcube)
    f = dict(grav=g["grav"], mag=g["mag"], drillfield=d0[d0 != 0], sensor_locations=o.sensor_grid(c), drilldata0=d0)
    inv, out = _gpu_cubing(cfg, f, gl=g["gl0"])
    stride = int(g["stride"])
    for n, cube in zip(CUBES, out):
        sub, ref_max = g["sub_" + n], float(g["max_" + n])
        got = np.asarray(cube).ravel()
        assert np.abs(got[::stride] - sub).max() / ref_max < 1e-5, n
        assert abs(float(got.sum()) - float(g["sum_" + n])) / (N * ref_max) < 1e-5, n
    assert abs(inv.logl - float(g["logl"])) < (1e-6 if prec == "fp64" else 1e-4) * abs(float(g["logl"]))


@pytest.mark.gpu
@pytest.mark.parametrize("kernel,structure", [("exp", "kron"), ("sparse", "compact"), ("matern32", "fft"), ("exp", "fft")])
def test_gpu_fft_and_friends_calc_logl_vs_oracle(kernel, structure):
    """Inversion.calc_logl (inversion.py:125-152; the objective optimize_gp evaluates hundreds of times) through the structured
    projections against the oracle's objective."""
    from geobo_b200 import _lib, config_loader, inversion
    from test_gpu_parity import synthetic_inputs
    cfg = _cfg((6, 5, 4), kernel, structure=structure)
    c = o.make_config(cfg)
    f = synthetic_inputs(c, 4)
    config_loader.load_settings(cfg, make_outpath=False)
    inv = inversion.Inversion()
    inv.create_cubegeometry()
    if kernel == "matern32":
        inv.gp_length = inv.gp_length * np.array([1.0, 1.01, 1.02])
    try:
        inv.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
        E, vp = o.cube_geometry(c)
        A = [o.a_sens(c, c.magneticField * 0, f["sensor_locations"], E, "grav"), o.a_sens(c, c.magneticField, f["sensor_locations"], E, "magn")]
        didx = o.drill_indices(f["drilldata0"])
        for params in ([1.0, 2.0, 1.0, 0.2, 0.2], [1.7, 3.1, 0.6, 0.5, 0.9]):
            with np.errstate(all="ignore"):
                ref = o.calc_logl(c, A, didx, inv.Fs3, params)
            got = inv.calc_logl(params)
            if np.isfinite(ref):
                assert abs(got - ref) < 1e-7 * abs(ref), (params, got, ref)
            else:
                assert not np.isfinite(got)                                           # matern32 with one common scale is singular
    finally:
        if inv._problem is not None:
            inv._problem.close()
            inv._problem = None
        _lib.default_context().release_cache()


@pytest.mark.parametrize("name", ["matern32_nd7", "sparse_nd7", "exp_nd7"])
def test_host_compiled_kernels_reproduce_the_live_reference_fixture(host, monkeypatch, name):
    """Outputs of the unmodified reference (all three kernels, tiny cube) with Pt computed by the DEVICE SOURCE of the FFT passes
    (csrc/fftconv.cuh compiled for the host, tables from csrc/formulas.cuh) inside the oracle's lean pipeline."""
    from test_compact import _pipeline_with_projection
    f = load_golden("cubing_%s.npz" % name)
    c = o.make_config(json.loads(str(f["cfg"])))

    def project(params, w, amp, A_list):
        N = A_list[0].shape[1]
        return _host_projection(host, c, params, w, amp, A_list, 0, N, 3)[0][:, :, :N]

    cubes, ex = _pipeline_with_projection(monkeypatch, c, f, project, gp_length=f["gl_before"])
    for n, a in zip(CUBES, cubes):
        assert normwise_err(a, f[n]) < 1e-7, n
    assert abs(ex["logl"] - float(f["logl"])) < 1e-7 * abs(float(f["logl"]))


def test_host_compiled_kernels_reproduce_the_committed_vtk_goldens(host, monkeypatch):
    """Example 1 of the reference (25 x 16 x 16, sparse kernel, M = 1056) with Pt computed by the DEVICE SOURCE of the FFT passes
    (padded lattice 32 x 64 x 32) inside the oracle's lean pipeline: the committed VTK cubes are reproduced to 2e-7."""
    from test_compact import _pipeline_with_projection
    f = load_golden("example1.npz")
    c = o.make_config(json.loads(str(f["cfg"])))

    def project(params, w, amp, A_list):
        N = A_list[0].shape[1]
        Pt, pad = _host_projection(host, c, params, w, amp, A_list, 0, N, 4)
        assert pad == (32, 64, 32)
        return Pt[:, :, :N]

    cubes, ex = _pipeline_with_projection(monkeypatch, c, f, project)
    for n, a in zip(CUBES, cubes):
        assert normwise_err(a, f["gold_" + n]) < 2e-7, n
    assert abs(ex["logl"] - float(f["logl"])) < 1e-6
