"""CPU tests of the digit-slice arithmetic restated in oracle/digit_slices.py (the scheme the tcgen05 int8 path runs):
digit range, reconstruction error, exactness against the product of the rounded operands, error against the fp64
product the reference computes (inversion.py:96), and the int32 overflow bound of the accumulator flush interval."""
import numpy as np
import pytest

from oracle import digit_slices as ds


@pytest.mark.parametrize("S", [4, 5, 6])
def test_balanced_digits_range_and_reconstruction(S):
    rng = np.random.default_rng(S)
    x = np.concatenate([rng.uniform(-0.5, 0.5, 20000), [0.5, -0.5, 0.0, 2.0**-60, -2.0**-60, 0.5 - 2.0**-50]])
    d = ds.balanced_digits(x, S)
    assert d.min() >= -128 and d.max() <= 127
    rec = ds.reconstruct(d, 0)
    ulp = 2.0 ** -(7 + 8 * (S - 1))
    assert np.abs(rec - x).max() <= 0.5 * ulp * (1 + 1e-9)          # round to nearest at the last digit
    assert abs(np.mean(rec - x)) < 0.02 * ulp                        # unbiased (balanced digits)


@pytest.mark.parametrize("S", [4, 5, 6])
def test_sliced_product_error_model(S):
    rng = np.random.default_rng(10 + S)
    m, n, k = 24, 20, 700
    A = rng.standard_normal((m, k)) * np.exp(rng.uniform(-12, 0, (m, k)))      # wide dynamic range inside a row (like A_sens)
    B = rng.uniform(0, 1, (n, k))                                              # covariance-like operand
    C, worst = ds.sliced_matmul(A, B, S)
    assert worst < 2 ** 31
    ref = A @ B.T
    # operand rounding 2^-(8 + 8 (S-1)) of each row scale + dropped levels >= S: a few ulps of the last digit per term
    rowA = np.abs(A).max(axis=1)[:, None] * 2
    rowB = np.abs(B).max(axis=1)[None, :] * 2
    bound = (S + 2) * 2.0 ** -(7 + 8 * (S - 1)) * rowA * rowB * np.sqrt(k) * 4
    assert (np.abs(C - ref) <= bound).all()
    # exactness: with S large enough that nothing is rounded the product is the fp64 product of the operands
    Ai = rng.integers(-100, 100, (m, 64)).astype(float) / 128.0
    Bi = rng.integers(-100, 100, (n, 64)).astype(float) / 128.0
    Ci, _ = ds.sliced_matmul(Ai, Bi, S)
    assert np.array_equal(Ci, Ai @ Bi.T)


def test_flush_interval_cannot_overflow_int32():
    # level S-1 sums S products of digits bounded by 128 in magnitude over `chunk` contraction indices
    for S in (4, 5, 6):
        assert S * 128 * 128 * 16384 < 2 ** 31
    # and the bound is attained in the worst case only: all digits -128
    A = np.full((1, 16384), -0.5)
    d = ds.balanced_digits(np.ldexp(A, -ds.scale_exp(0.5)), 6)
    assert d.min() >= -128


@pytest.mark.parametrize("kf,S,tol_mu,tol_var", [("exp", 5, 1e-8, 1e-6), ("matern32", 5, 1e-8, 1e-6), ("exp", 4, 1e-5, 1e-4)])
def test_sliced_pipeline_error_model_vs_fp64_oracle(kf, S, tol_mu, tol_var):
    """The whole `precision: int8xS` pipeline restated on the CPU (digit-slice products + fp64 Cholesky + one refinement
    step + matrix-free mean) against the oracle's fp64 predict3 on a small cube: the mean is independent of the operand
    rounding after refinement, the variance error is the slice error."""
    import json
    from conftest import load_golden
    from oracle import numpy_oracle as o
    cfg = json.loads(str(load_golden("sens_8x6x5.npz")["cfg"]))
    cfg.update(xNcube=4, yNcube=4, zNcube=8, kernelfunc=kf)
    c = o.make_config(cfg)
    E, vp = o.cube_geometry(c)
    loc = o.sensor_grid(c)
    Ag = o.a_sens(c, c.magneticField * 0, loc, E, "grav")
    Am = o.a_sens(c, c.magneticField, loc, E, "magn")
    N, Ns = Ag.shape[1], Ag.shape[0]
    rng = np.random.default_rng(3)
    didx = np.sort(rng.choice(N, 5, replace=False))
    y = rng.standard_normal(2 * Ns + 5)
    gl = c.gp_lengthscale * c.xvoxsize * (np.array([1.0, 1.01, 1.02]) if kf == "matern32" else np.ones(3))
    sig, w, amp = np.asarray(c.gp_err, float), np.asarray(c.gp_coeff, float), 1.0
    mu_ref, var_ref, _, _ = o.predict_lean(c, [Ag, Am], didx, y, gl.copy(), sig, w, amp)
    pts = o.grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    kcov = amp * o.create_cov(o.sqdist(pts), gl.copy(), w, fkernel=kf)
    mu, var = ds.predict_sliced([Ag, Am], didx, kcov, y, sig, amp, S, refine=1)
    assert np.abs(mu - mu_ref).max() <= tol_mu * np.abs(mu_ref).max()
    assert np.abs(var - var_ref).max() <= tol_var * np.abs(var_ref).max()
    if S == 5:      # without refinement the mean carries the slice error of AkA and Pt
        mu0, _ = ds.predict_sliced([Ag, Am], didx, kcov, y, sig, amp, S, refine=0)
        assert np.abs(mu0 - mu_ref).max() >= np.abs(mu - mu_ref).max()


def test_device_digit_extraction_source_equals_the_restatement(tmp_path):
    """csrc/ozaki.cuh (ozaki::digits<S>, ozaki::scale_exp) compiled for the host: the bytes the device kernels would
    write are the digits of ``ds.balanced_digits`` bit for bit, and the scaling exponents agree."""
    import ctypes
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    so = tmp_path / "ozaki_host.so"
    from conftest import HARNESS_CXX
    subprocess.run(HARNESS_CXX + ["-I" + cuda_inc,
                    os.path.join(root, "tests", "host_harness", "ozaki_host.cpp"), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.host_scale_exp.argtypes = [ctypes.c_double]
    rng = np.random.default_rng(99)
    t = np.concatenate([rng.uniform(-0.5, 0.5, 50000), [0.5, -0.5, 0.0, -0.0, 2.0**-60, -2.0**-60, 0.5 - 2.0**-50, -0.5 + 2.0**-53],
                        np.ldexp(rng.uniform(-0.5, 0.5, 2000), rng.integers(-40, 0, 2000))])
    t = np.ascontiguousarray(t)
    for S in (4, 5, 6):
        out = np.empty((t.size, S), dtype=np.uint8)
        assert lib.host_digits(S, t.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(t.size), out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.array_equal(out.view(np.int8).astype(np.int64), ds.balanced_digits(t, S)), S
    for amax in [0.0, -1.0, np.nan, np.inf, 1.0, 0.5, 0.75, 1e-300, 3.7e5, 2.0**-20, np.nextafter(1.0, 0)]:
        assert lib.host_scale_exp(amax) == ds.scale_exp(amax), amax
