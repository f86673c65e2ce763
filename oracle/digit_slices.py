"""CPU restatement (NumPy, exact integer arithmetic) of the balanced 8-bit digit-slice product that the tcgen05 path
runs on the int8 tensor cores (geobo_b200/csrc/ozaki.cuh `digits`, ozaki.cu / ozaki_gemm.cu).  TEST INFRASTRUCTURE ONLY:
imported by tests/ to pin the arithmetic of the slice scheme (digit range, reconstruction error, exactness of the
integer accumulation, int32 overflow bound); the product path never imports it.

There is no counterpart in the reference (its products are fp64 dgemm, geobo/inversion.py:96,114,117); this module
restates OUR algorithm so that its error model is checked on the CPU against the fp64 product the reference computes.
"""
import numpy as np


def scale_exp(amax):
    """exponent e with |x| <= 2^(e-1) for all |x| <= amax (ozaki.cuh scale_exp)."""
    if not (amax > 0.0) or not np.isfinite(amax):
        return 0
    _, e = np.frexp(amax)          # amax = m 2^e, m in [0.5, 1)
    return int(e) + 1


def balanced_digits(t, S):
    """t in [-1/2, 1/2] -> int array (..., S) of balanced digits d_q in [-128, 127] with
    t ~= sum_q d_q 2^-(7 + 8 q), rounded to nearest at the last digit (ozaki.cuh digits<S>)."""
    t = np.asarray(t, dtype=np.float64) + np.ldexp(1.0, -(7 + 8 * (S - 1) + 1))
    v = np.empty(t.shape + (S,), dtype=np.int64)
    x = t * 128.0
    f = np.floor(x)
    v[..., 0] = f.astype(np.int64)
    r = x - f
    for q in range(1, S):
        x = r * 256.0
        f = np.floor(x)
        v[..., q] = f.astype(np.int64)
        r = x - f
    for q in range(S - 1, 0, -1):
        carry = v[..., q] >= 128
        v[..., q] -= 256 * carry
        v[..., q - 1] += carry
    return v


def reconstruct(d, e):
    """value represented by digits d (..., S) with exponent(s) e."""
    S = d.shape[-1]
    acc = np.zeros(d.shape[:-1], dtype=object)
    for q in range(S):
        acc = acc * 256 + d[..., q].astype(object)
    # acc = sum_q d_q 256^(S-1-q); value = 2^e * acc * 2^-(7 + 8 (S-1))
    return np.asarray(acc, dtype=np.float64) * np.ldexp(1.0, np.asarray(e) - 7 - 8 * (S - 1))


def slice_rows(X, S):
    """per-row exponent + digits of a matrix (the layout-independent content of slice_rows_tiled_kernel)."""
    X = np.asarray(X, dtype=np.float64)
    e = np.array([scale_exp(np.abs(r).max()) if r.size else 0 for r in X], dtype=np.int64)
    d = balanced_digits(np.ldexp(X, -e[:, None]), S)
    return d, e


def sliced_matmul(A, B, S, chunk=16384):
    """C = A . B^T with A (m, k), B (n, k) as the device computes it: digit products (qa, qb) with qa + qb < S accumulate
    exactly (int64 here; the device uses int32 per level and flushes every `chunk` contraction indices), levels are
    recombined exactly and scaled in fp64.  Returns (C, max_abs_level_sum) -- the latter must stay below 2^31."""
    da, ea = slice_rows(A, S)
    db, eb = slice_rows(B, S)
    m, k = A.shape
    n = B.shape[0]
    C = np.zeros((m, n))
    worst = 0
    for k0 in range(0, k, chunk):
        k1 = min(k, k0 + chunk)
        lvl = [np.zeros((m, n), dtype=np.int64) for _ in range(S)]
        for qa in range(S):
            for qb in range(S - qa):
                lvl[qa + qb] += da[:, k0:k1, qa] @ db[:, k0:k1, qb].T
        worst = max(worst, max(int(np.abs(x).max()) for x in lvl))
        acc = np.zeros((m, n), dtype=object)
        for x in lvl:
            acc = acc * 256 + x.astype(object)
        scale = np.ldexp(1.0, (ea[:, None] + eb[None, :] - 14 - 8 * (S - 1)).astype(np.int64))
        C += scale * np.asarray(acc, dtype=np.float64)
    return C, worst
