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


def predict_sliced(A_list, didx, kcov, y, sig, amp, S, refine=1):
    """CPU restatement of the device pipeline of `precision: int8xS` on a small dense problem:
    Pt = A3.K, AkA and L^-1.Pt as digit-slice products (exact integer accumulation of S balanced digits per operand,
    `sliced_matmul`), fp64 Cholesky, `refine` steps of iterative refinement of alpha = (A K A^T + Sigma)^-1 y against
    the fp64 operator, mean = K A3^T alpha, variance = amp - colsumsq(Linv . Pt).

    A_list = [A_grav, A_magn] (Ns x N), didx = drilled voxel indices, kcov = dense 3N x 3N covariance (fp64, from the
    oracle), y = data vector.  Returns (mu, var).  Used by tests/test_digit_slices.py to check the ERROR MODEL of the
    scheme against the oracle's fp64 predict3 (inversion.py:96-117) without a GPU."""
    from scipy.linalg import cholesky, solve_triangular
    Ns, N = A_list[0].shape
    nd = len(didx)
    M = 2 * Ns + nd
    K = np.asarray(kcov).reshape(3, N, 3, N)
    # Pt[(c, s), (r, i)] = sum_j A_c[s, j] K[(c, j), (r, i)]  -- one exponent per sensor row and per covariance block
    Pt = np.zeros((M, 3 * N))
    for c in range(2):
        for r in range(3):
            Kb = K[c, :, r, :]                       # (j, i)
            scale = 2.0 ** scale_exp(np.abs(Kb).max())
            P, worst = sliced_matmul(A_list[c], (Kb / scale).T, S)       # operands: rows = A rows, rows = output voxels i
            assert worst < 2 ** 31
            Pt[c * Ns:(c + 1) * Ns, r * N:(r + 1) * N] = P * scale
    for r in range(3):
        Pt[2 * Ns:, r * N:(r + 1) * N] = K[2, didx, r, :]                 # one-hot drill rows: exact gathers
    # AkA: block (c', c) = Pt[(c', .), (c, .)] . A_c^T  as digit-slice products; drill rows / columns are gathers
    AkA = np.zeros((M, M))
    for cp in range(2):
        for c in range(2):
            blk, _ = sliced_matmul(Pt[cp * Ns:(cp + 1) * Ns, c * N:(c + 1) * N], A_list[c], S)
            AkA[cp * Ns:(cp + 1) * Ns, c * Ns:(c + 1) * Ns] = blk
    AkA[2 * Ns:, :] = Pt[:, 2 * N + np.asarray(didx, dtype=int)].T if nd else AkA[2 * Ns:, :]
    AkA[:, 2 * Ns:] = AkA[2 * Ns:, :].T
    AkA = np.tril(AkA) + np.tril(AkA, -1).T                               # the device only forms the lower triangle
    sig2 = np.hstack((np.full(Ns, sig[0] ** 2), np.full(Ns, sig[1] ** 2), np.full(nd, sig[2] ** 2)))
    AkA[np.diag_indices(M)] += sig2
    L = cholesky(AkA, lower=True)
    Linv = solve_triangular(L, np.eye(M), lower=True)
    alpha = Linv.T @ (Linv @ y)

    def a3t(v):                                                            # A3^T v
        w = np.zeros(3 * N)
        w[:N] = A_list[0].T @ v[:Ns]
        w[N:2 * N] = A_list[1].T @ v[Ns:2 * Ns]
        if nd:
            w[2 * N + np.asarray(didx, dtype=int)] = v[2 * Ns:]
        return w

    def a3(z):                                                             # A3 z
        t = np.empty(M)
        t[:Ns] = A_list[0] @ z[:N]
        t[Ns:2 * Ns] = A_list[1] @ z[N:2 * N]
        if nd:
            t[2 * Ns:] = z[2 * N + np.asarray(didx, dtype=int)]
        return t

    Kd = np.asarray(kcov)
    for _ in range(refine):
        r = y - a3(Kd @ a3t(alpha)) - sig2 * alpha                         # fp64 matrix-free residual
        alpha = alpha + Linv.T @ (Linv @ r)
    mu = Kd @ a3t(alpha)
    V, _ = sliced_matmul(Linv, Pt.T.copy(), S)                            # rows of Linv x columns of Pt
    var = amp - np.einsum("ij,ij->j", V, V)
    return mu, var
