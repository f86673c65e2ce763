"""Separable (Kronecker) form of the exp covariance blocks -- settings key ``structure: kron``.

CPU: (1) the algebra -- the oracle's restatement ``oracle/kron.py`` (three Toeplitz mode products per block, factor lines
from the oracle's own ``cov_block``) against the oracle's dense ``pt_panel`` / ``create_cov``; (2) the device source --
``csrc/kron.cuh`` compiled for the host and driven with the launch geometry of ``csrc/kron.cu`` (row chunks, ragged tiles,
voxel-column shards that cut through a plane, accumulation over data blocks) against the dense oracle.
GPU (``-m gpu``): ``Inversion.cubing`` with ``structure: kron`` for both precisions against the oracle and against the
dense device path, the loud refusal for non-separable kernels, and BASELINE config 2 at full size.
"""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import CUBES, GOLDEN, HARNESS_CXX, ROOT, load_golden, normwise_err
from oracle import kron as kr
from oracle import numpy_oracle as o

KID = {"sparse": 0, "exp": 1, "matern32": 2}


def _cfg(shape, **kw):
    cfg = json.loads(str(load_golden("sens_8x6x5.npz")["cfg"]))
    cfg.update(xNcube=shape[0], yNcube=shape[1], zNcube=shape[2], kernelfunc="exp", **kw)
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
    params = o.dedup_lengths(mult * np.asarray([c.xvoxsize] * 3))          # Q1: three distinct scales
    return c, N, A, params, np.asarray(w, dtype=float), amp


# ------------------------------------------------------------------------------------------------ (1) the algebra
@pytest.mark.parametrize("shape", [(5, 4, 6), (3, 7, 2), (1, 5, 4)])
def test_oracle_kron_projection_equals_dense_projection(shape):
    c, N, A, params, w, amp = _case(shape, 1)
    didx = np.array([0, N // 2, N - 1])
    dense = _dense_pt(c, params, w, amp, A, didx)
    got = kr.pt_kron(c, params, w, amp, A, didx)
    assert got.shape == dense.shape
    assert np.abs(got - dense).max() <= 1e-13 * np.abs(dense).max()


def test_oracle_kron_matvec_equals_create_cov_matvec():
    c, N, A, params, w, amp = _case((4, 3, 5), 2)
    pts = o.grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    K = amp * o.create_cov(o.sqdist(pts), params.copy(), w, "exp")                   # (3N, 3N), kernels.py:158-195
    W = np.random.default_rng(3).standard_normal((3, N))
    want = (K @ W.ravel()).reshape(3, N)
    got = kr.kw_kron(c, params, w, amp, W)
    assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()


def test_oracle_kron_zero_cross_weight_zeroes_the_block():
    c, N, A, params, w, amp = _case((3, 4, 5), 4, w=(0.0, 0.5, 0.0))
    dense = _dense_pt(c, params, w, amp, A, np.zeros(0, dtype=int))
    got = kr.pt_kron(c, params, w, amp, A, np.zeros(0, dtype=int))
    assert not got[: A[0].shape[0], 2].any() and not got[: A[0].shape[0], 1].any()   # density-drill and density-magsus weights are 0
    assert np.abs(got - dense).max() <= 1e-13 * np.abs(dense).max()


# ------------------------------------------------------------------------------------------------ (2) the device source on the host
@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = tmp_path_factory.mktemp("kron_host") / "kron_host.so"
    src = os.path.join(ROOT, "tests", "host_harness", "kron_host.cpp")
    subprocess.run(HARNESS_CXX + [src, "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    P, L, D, I = ctypes.c_void_p, ctypes.c_long, ctypes.c_double, ctypes.c_int
    lib.kron_host_apply.argtypes = [I, P, P, D, P, P, I, P, L, L, L, L, L, P, L, L, I]
    lib.kron_host_apply.restype = None
    lib.kron_host_set_thread_order.argtypes = [I]

    def apply(c, params, w, amp, blk0, A, c0, c1, chunk_rows, out, ldo, r_stride_out, accumulate, kernel="exp"):
        p = lambda a: a.ctypes.data_as(P)                                            # noqa: E731
        l = np.ascontiguousarray(params, dtype=float)
        ww = np.ascontiguousarray(w, dtype=float)
        ncube = np.array([c.xNcube, c.yNcube, c.zNcube], dtype=np.int64)
        vox = np.array([c.xvoxsize, c.yvoxsize, c.zvoxsize], dtype=float)
        A = np.ascontiguousarray(A, dtype=float)
        assert out.flags.c_contiguous and out.dtype == np.float64
        lib.kron_host_apply(KID[kernel], p(l), p(ww), amp, p(ncube), p(vox), blk0, p(A), A.shape[1], A.shape[0], c0, c1, chunk_rows, p(out),
                            ldo, r_stride_out, accumulate)
    apply.set_thread_order = lib.kron_host_set_thread_order
    return apply


def _host_projection(host, c, params, w, amp, A, c0, c1, chunk_rows):
    """Pt rows of the two survey blocks for the voxel columns [c0, c1), laid out like the device's Pt: [row][r][ncp]."""
    Ns, ncol = A[0].shape[0], c1 - c0
    ncp = -(-ncol // 32) * 32
    Pt = np.full((2 * Ns, 3 * ncp), np.nan)
    for cb in range(2):
        host(c, params, w, amp, cb * 3, A[cb], c0, c1, chunk_rows, Pt[cb * Ns:], 3 * ncp, ncp, 0)
    return Pt.reshape(2 * Ns, 3, ncp)


@pytest.mark.parametrize("shape,shard,chunk", [
    ((5, 11, 6), None, 1000),          # ragged everywhere; 11 y-rows = two groups of KRON_JT = 8 output rows
    ((5, 11, 6), (48, 272), 7),        # a shard that starts and ends inside a plane (30 voxels per y-row); several row chunks
    ((20, 3, 16), None, 3),            # 320 voxels per plane: two plane tiles, the second one partly empty
    ((7, 2, 33), (0, 400), 1),         # zN = 33: nine strips of 4, the last one ragged; one row per chunk
    ((1, 9, 1), None, 4),              # degenerate axes
])
def test_host_compiled_kernels_vs_dense_oracle(host, shape, shard, chunk):
    c, N, A, params, w, amp = _case(shape, 5, nrows=6)
    c0, c1 = shard if shard else (0, N)
    dense = _dense_pt(c, params, w, amp, A, np.zeros(0, dtype=int))                  # (2 * 6, 3, N)
    Pt = _host_projection(host, c, params, w, amp, A, c0, c1, chunk)
    got = Pt[:, :, :c1 - c0]
    assert np.isfinite(got).all()                                                    # every column of the shard was written
    assert np.isnan(Pt[:, :, c1 - c0:]).all()                                        # and nothing outside it
    assert np.abs(got - dense[:, :, c0:c1]).max() <= 1e-13 * np.abs(dense).max()


def test_host_compiled_kernels_do_not_depend_on_the_thread_order_inside_a_phase(host):
    """Between two barriers the harness may run the threads of a block in any order (ascending, descending, odd ids first): a
    phase in which one thread reads what another one writes would give a different result."""
    c, N, A, params, w, amp = _case((5, 11, 6), 9, nrows=3)
    try:
        res = []
        for order in (0, 1, 2):
            host.set_thread_order(order)
            res.append(_host_projection(host, c, params, w, amp, A, 48, 272, 2))
    finally:
        host.set_thread_order(0)
    assert np.array_equal(res[0], res[1], equal_nan=True) and np.array_equal(res[0], res[2], equal_nan=True)


def test_host_compiled_matvec_accumulates_over_data_blocks(host):
    """z = K w as the refinement uses it: three applications with one row each, the second and third accumulating."""
    c, N, A, params, w, amp = _case((6, 5, 4), 6)
    W = np.random.default_rng(7).standard_normal((3, N))
    c0, c1 = 16, 112
    ncp = 96
    z = np.full((3, ncp), np.nan)
    for cb in range(3):
        host(c, params, w, amp, cb * 3, W[cb][None, :], c0, c1, 1, z, 0, ncp, int(cb > 0))
    want = kr.kw_kron(c, params, w, amp, W)[:, c0:c1]
    assert np.abs(z - want).max() <= 1e-13 * np.abs(want).max()


def test_host_compiled_kernels_zero_block(host):
    c, N, A, params, w, amp = _case((4, 3, 5), 8, nrows=2, w=(0.0, 0.5, 0.0))
    Pt = _host_projection(host, c, params, w, amp, A, 0, N, 8)[:, :, :N]
    dense = _dense_pt(c, params, w, amp, A, np.zeros(0, dtype=int))
    assert not Pt[:2, 1].any() and not Pt[:2, 2].any() and np.isfinite(Pt).all()
    assert np.abs(Pt - dense).max() <= 1e-13 * np.abs(dense).max()


def test_structure_key_is_validated():
    from geobo_b200 import _lib, config_loader, inversion
    config_loader.load_settings(_cfg((4, 4, 4), structure="wavelet"), make_outpath=False)
    with pytest.raises(ValueError, match="structure"):
        inversion.Inversion()._structure()
    with pytest.raises(ValueError, match="structure"):
        _lib.Problem.hyper([1, 1, 1], [1, 1, 1], [1, 1, 1], 1.0, "exp", structure="wavelet")
    h = _lib.Problem.hyper([1, 1, 1], [1, 1, 1], [1, 1, 1], 1.0, "exp", structure="kron")
    assert h.structure == 1 and _lib.Problem.hyper([1, 1, 1], [1, 1, 1], [1, 1, 1], 1.0, "exp").structure == 0
    config_loader.load_settings(_cfg((4, 4, 4)), make_outpath=False)
    assert inversion.Inversion()._structure() == "dense"                             # reference YAMLs have no such key
    for kf, want in (("exp", "kron"), ("sparse", "compact"), ("matern32", "fft")):
        config_loader.load_settings(dict(_cfg((4, 4, 4), structure="auto"), kernelfunc=kf), make_outpath=False)
        assert inversion.Inversion()._structure() == want


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
        launches = inv._problem.timings().get("launches") if hasattr(inv._problem, "timings") else None
    finally:
        if inv._problem is not None:
            inv._problem.close()
            inv._problem = None
        _lib.default_context().release_cache()
    return inv, out, launches


@pytest.mark.gpu
@pytest.mark.parametrize("shape,nd,prec", [((7, 5, 3), 3, "fp64"), ((5, 11, 6), 0, "fp64"), ((20, 4, 16), 4, "fp64"),
                                           ((5, 4, 16), 3, "int8x5"), ((11, 4, 48), 0, "int8x6")])
def test_gpu_kron_cubing_vs_oracle(shape, nd, prec):
    from test_gpu_parity import synthetic_inputs
    cfg = _cfg(shape, structure="kron", precision=prec)
    c = o.make_config(cfg)
    f = synthetic_inputs(c, nd)
    with np.errstate(all="ignore"):
        ref, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    inv, out, _ = _gpu_cubing(cfg, f)
    tol = 1e-7 if prec in ("fp64", "int8x6") else 1e-6
    for n, a, r in zip(CUBES, out, ref):
        assert normwise_err(a, r) < tol, n
    assert abs(inv.logl - ex["logl"]) < (1e-7 if prec == "fp64" else 1e-4) * abs(ex["logl"])


@pytest.mark.gpu
def test_gpu_kron_matches_dense_device_path_and_scratch_chunks(monkeypatch):
    """Same cube through structure dense and kron (fp64): the six cubes agree far below the parity tolerance; with a 1 MB
    scratch the rows go through many chunks and the result does not change."""
    from test_gpu_parity import synthetic_inputs
    cfg = _cfg((16, 16, 16))
    f = synthetic_inputs(o.make_config(cfg), 50)
    _, dense, _ = _gpu_cubing(dict(cfg, structure="dense"), f)
    _, kron, _ = _gpu_cubing(dict(cfg, structure="kron"), f)
    monkeypatch.setenv("GEOBO_B200_KRON_SCRATCH_MB", "1")
    _, kron_small, _ = _gpu_cubing(dict(cfg, structure="kron"), f)
    for n, a, b, b2 in zip(CUBES, dense, kron, kron_small):
        assert normwise_err(b, a) < 1e-9, n
        assert normwise_err(b2, b) < 1e-12, n      # same arithmetic per row whatever the chunking


@pytest.mark.gpu
@pytest.mark.parametrize("kf", ["sparse", "matern32"])
def test_gpu_kron_refuses_non_separable_kernels(kf):
    from geobo_b200 import _lib
    from test_gpu_parity import synthetic_inputs
    cfg = dict(_cfg((5, 4, 3), structure="kron"), kernelfunc=kf)
    f = synthetic_inputs(o.make_config(cfg), 0)
    with pytest.raises(_lib.GeoboB200Error, match="kron"):
        _gpu_cubing(cfg, f)


@pytest.mark.gpu
@pytest.mark.parametrize("name,prec", [("cfg2", "fp64"), ("cfg2", "int8x5"), ("cfg3e", "int8x5")])
def test_gpu_kron_fullsize_vs_cpu_oracle(name, prec):
    """BASELINE config 2 (32x32x32, exp) and the north star's 64x64x32 two-property cube (exp) through structure kron against the
    dense oracle's full-size results (1e-5, north star)."""
    if not os.path.exists(os.path.join(GOLDEN, "fullsize_%s.npz" % name)):
        pytest.skip("fixture fullsize_%s.npz not generated" % name)
    g = load_golden("fullsize_%s.npz" % name)
    cfg = dict(json.loads(str(g["cfg"])), precision=prec, structure="kron")
    c = o.make_config(cfg)
    N = c.xNcube * c.yNcube * c.zNcube
    d0 = np.zeros(N)
    d0[g["didx"]] = g["drillvals"]
    d0 = d0.reshape(c.xNcube, c.yNcube, c.zNcube)
    f = dict(grav=g["grav"], mag=g["mag"], drillfield=d0[d0 != 0], sensor_locations=o.sensor_grid(c), drilldata0=d0)
    inv, out, _ = _gpu_cubing(cfg, f, gl=g["gl0"])
    stride = int(g["stride"])
    for n, cube in zip(CUBES, out):
        sub, ref_max = g["sub_" + n], float(g["max_" + n])
        got = np.asarray(cube).ravel()
        if np.isnan(sub).all():
            assert np.isnan(got).all(), n
            continue
        assert np.abs(got[::stride] - sub).max() / ref_max < 1e-5, n
        assert abs(float(got.sum()) - float(g["sum_" + n])) / (N * ref_max) < 1e-5, n
    assert abs(inv.logl - float(g["logl"])) < (1e-6 if prec == "fp64" else 1e-4) * abs(float(g["logl"]))


@pytest.mark.gpu
def test_gpu_kron_compact_fft_two_rank_shards_match_oracle():
    """Two ranks, voxel-column shards that meet inside an x-z plane: every rank applies the y mode for its own y-rows only and
    stores its own columns; same for the tap-sum and FFT paths (tests/mgpu_check.py with GEOBO_B200_MGPU_STRUCTURED=1)."""
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29950 + os.getpid() % 40), os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, GEOBO_B200_MGPU_STRUCTURED="1"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_OK world=2" in r.stdout


# ------------------------------------------------------------------------------------------------ CPU: outputs of the unmodified reference
@pytest.mark.parametrize("name,structure", [("exp_nd7", "kron"), ("exp_nd0", "kron"), ("exp_nd7", "fft"), ("matern32_nd7", "fft"), ("sparse_nd7", "compact")])
def test_structured_oracle_reproduces_the_live_reference_fixtures(monkeypatch, name, structure):
    """tests/golden/cubing_*.npz hold Inversion.cubing outputs of the UNMODIFIED reference (make_golden.py).  The oracle's lean
    pipeline with Pt from the structured restatements reproduces them to the same 1e-7 as with the dense panels."""
    from oracle import fftconv as fc
    from oracle import stencil as st
    f = load_golden("cubing_%s.npz" % name)
    c = o.make_config(json.loads(str(f["cfg"])))
    cache = {}
    dense_panel = o.pt_panel

    def structured_panel(c_, params, w, amp, A_list, didx, pts, cols, **kw):
        if cache.get("building"):                  # the drill rows inside the structured restatements stay dense gathers
            return dense_panel(c_, params, w, amp, A_list, didx, pts, cols, **kw)
        if "pt" not in cache:
            cache["building"] = True
            cache["pt"] = {"kron": lambda: kr.pt_kron(c_, params, w, amp, A_list, didx),
                           "compact": lambda: st.pt_compact(c_, params, w, amp, A_list, didx),
                           "fft": lambda: fc.pt_fft(c_, params, w, amp, c_.kernelfunc, A_list, didx)}[structure]()
            cache["building"] = False
        return cache["pt"][:, :, cols]

    monkeypatch.setattr(o, "pt_panel", structured_panel)
    with np.errstate(all="ignore"):
        cubes, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"], gp_length=np.array(f["gl_before"], dtype=float))
    assert "pt" in cache
    for n, a in zip(CUBES, cubes):
        assert normwise_err(a, f[n]) < 1e-7, n
    assert abs(ex["logl"] - float(f["logl"])) < 1e-7 * abs(float(f["logl"]))


@pytest.mark.parametrize("name", ["exp_nd7", "exp_nd0"])
def test_host_compiled_kernels_reproduce_the_live_reference_fixture(host, monkeypatch, name):
    """Outputs of the unmodified reference (exp kernel, tiny cube) with Pt computed by the DEVICE SOURCE of the Kronecker kernels
    (csrc/kron.cuh compiled for the host, tables from csrc/formulas.cuh) inside the oracle's lean pipeline."""
    from test_compact import _pipeline_with_projection
    f = load_golden("cubing_%s.npz" % name)
    c = o.make_config(json.loads(str(f["cfg"])))

    def project(params, w, amp, A_list):
        N = A_list[0].shape[1]
        return _host_projection(host, c, params, w, amp, A_list, 0, N, 7)[:, :, :N]

    cubes, ex = _pipeline_with_projection(monkeypatch, c, f, project, gp_length=f["gl_before"])
    for n, a in zip(CUBES, cubes):
        assert normwise_err(a, f[n]) < 1e-7, n
    assert abs(ex["logl"] - float(f["logl"])) < 1e-7 * abs(float(f["logl"]))
