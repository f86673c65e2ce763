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
