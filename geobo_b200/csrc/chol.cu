// Blocked fp64 Cholesky of AkA (inversion.py:100, LAPACK dpotrf) and the blocked forward solve
// V = L^-1 Pt (inversion.py:105,114, LAPACK dtrtrs with 3N right-hand sides).
//
// The matrix is padded to a multiple of NB = 128 with an identity tail, so every panel is full.
// Per panel:  potrf128 (one CTA, shared memory; also emits the inverse of the diagonal block and
// accumulates log det) -> panel solve as a GEMM with that inverse -> trailing update (lower tiles
// only) on the DMMA engine.  The triangular solve with 3N right-hand sides is left-looking: each
// 128-row block of V is one wide GEMM against all previous rows, then a GEMM with the block inverse.
#include "common.cuh"

constexpr int NB = 128;
constexpr int PLD = NB + 1;   // padded leading dimension in shared memory

// Factor one 128x128 diagonal block in place (lower), write its inverse to linv (row-major, upper zero).
__global__ void __launch_bounds__(256, 1) potrf128_kernel(double* __restrict__ Bm, long ldb, int k0, int Mtrue,
                                                          double* __restrict__ linv, double* __restrict__ logdet,
                                                          int* __restrict__ info) {
    extern __shared__ double S[];   // [NB][PLD]; lower = L, strict upper = X^T (inverse), xd = diag of inverse
    __shared__ double xd[NB];
    __shared__ int bad;
    const int tid = threadIdx.x;
    double* blk = Bm + (long)k0 * ldb + k0;
    if (tid == 0) bad = 0;
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int r = e / NB, c = e % NB;
        S[r * PLD + c] = (c <= r) ? blk[(long)r * ldb + c] : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < NB; ++k) {
        if (tid == 0) {
            const double d = S[k * PLD + k];
            if (!(d > 0.0) && bad == 0) bad = k0 + k + 1;
            S[k * PLD + k] = sqrt(d);
        }
        __syncthreads();
        const double piv = S[k * PLD + k];
        for (int r = k + 1 + tid; r < NB; r += blockDim.x) S[r * PLD + k] = S[r * PLD + k] / piv;
        __syncthreads();
        // trailing update of the lower triangle: element (r, c), k < c <= r
        const int rem = NB - 1 - k;
        for (int e = tid; e < rem * rem; e += blockDim.x) {
            const int r = k + 1 + e / rem, c = k + 1 + e % rem;
            if (c <= r) S[r * PLD + c] -= S[r * PLD + k] * S[c * PLD + k];
        }
        __syncthreads();
    }
    // inverse of the triangular block: thread c owns column c of X = L^-1, stored transposed in the upper half
    if (tid < NB) {
        const int c = tid;
        const double xc = 1.0 / S[c * PLD + c];
        xd[c] = xc;
        for (int r = c + 1; r < NB; ++r) {
            double s0 = S[r * PLD + c] * xc, s1 = 0.0;
            int k = c + 1;
            for (; k + 1 < r; k += 2) {
                s0 = fma(S[r * PLD + k], S[c * PLD + k], s0);
                s1 = fma(S[r * PLD + k + 1], S[c * PLD + k + 1], s1);
            }
            if (k < r) s0 = fma(S[r * PLD + k], S[c * PLD + k], s0);
            S[c * PLD + r] = -(s0 + s1) / S[r * PLD + r];
        }
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int r = e / NB, c = e % NB;
        if (c <= r) blk[(long)r * ldb + c] = S[r * PLD + c];
        linv[e] = (c < r) ? S[c * PLD + r] : (c == r ? xd[r] : 0.0);
    }
    if (tid == 0) {
        double ld = 0.0;
        for (int k = 0; k < NB; ++k)
            if (k0 + k < Mtrue) {
                const double l = S[k * PLD + k];
                ld += log(l * l);                     // inversion.py:108: log(diag(L)**2)
            }
        *logdet += ld;
        if (bad && *info == 0) *info = bad;
    }
}

static gemm::Task make_task(const double* A, long lda, const double* B, long ldb, const double* Cin, long ldcin, double* C,
                            long ldc, int M, int N, int K, double alpha, double beta, int lower) {
    gemm::Task t;
    memset(&t, 0, sizeof t);
    t.A = A; t.lda = lda; t.B = B; t.ldb = ldb; t.Cin = Cin; t.ldcin = ldcin; t.C = C; t.ldc = ldc;
    t.M = M; t.N = N; t.K = K; t.alpha = alpha; t.beta = beta; t.lower = lower;
    return t;
}

cudaError_t chol_factor(double* Bm, long ldb, int Mp, int Mtrue, const CholWork& w, cudaStream_t s) {
    static bool attr_set = false;
    const int smem = NB * PLD * (int)sizeof(double);
    cudaError_t e;
    if (!attr_set) {
        e = cudaFuncSetAttribute(potrf128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int nblk = Mp / NB;
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB;
        double* linv = w.linv + (long)kb * NB * NB;
        potrf128_kernel<<<1, 256, smem, s>>>(Bm, ldb, k0, Mtrue, linv, w.logdet, w.info);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const int rest = Mp - k0 - NB;
        if (rest <= 0) break;
        double* panel = Bm + (long)(k0 + NB) * ldb + k0;
        // L21 = A21 . L11^-T : C[m, n] = sum_k A21[m, k] * Linv[n, k]   (in place: one CTA owns whole rows)
        gemm::TaskBatch b1;
        b1.n = 1;
        b1.t[0] = make_task(panel, ldb, linv, NB, nullptr, 0, panel, ldb, rest, NB, NB, 1.0, 0.0, 0);
        e = gemm::launch(b1, gemm::B_T, s);
        if (e != cudaSuccess) return e;
        // A22 -= L21 . L21^T  (lower tiles only)
        double* trail = Bm + (long)(k0 + NB) * ldb + (k0 + NB);
        gemm::TaskBatch b2;
        b2.n = 1;
        b2.t[0] = make_task(panel, ldb, panel, ldb, trail, ldb, trail, ldb, rest, rest, NB, -1.0, 1.0, 1);
        e = gemm::launch(b2, gemm::B_T, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t chol_forward_solve(const double* L, long ldl, int Mp, const CholWork& w, double* Pt, long ldp, int ncols,
                               double* tmp, cudaStream_t s) {
    const int nblk = Mp / NB;
    cudaError_t e;
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB;
        double* rowblk = Pt + (long)k0 * ldp;
        // tmp = Pt[kb] - L[kb, 0:k0] . V[0:k0]
        gemm::TaskBatch a;
        a.n = 1;
        a.t[0] = make_task(L + (long)k0 * ldl, ldl, Pt, ldp, rowblk, ldp, tmp, ldp, NB, ncols, k0, -1.0, 1.0, 0);
        e = gemm::launch(a, gemm::B_N, s);
        if (e != cudaSuccess) return e;
        // V[kb] = L[kb,kb]^-1 . tmp
        gemm::TaskBatch b;
        b.n = 1;
        b.t[0] = make_task(w.linv + (long)kb * NB * NB, NB, tmp, ldp, nullptr, 0, rowblk, ldp, NB, ncols, NB, 1.0, 0.0, 0);
        e = gemm::launch(b, gemm::B_N, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
