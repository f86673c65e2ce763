// Blocked fp64 Cholesky of AkA (inversion.py:100, LAPACK dpotrf) and the blocked forward solve
// V = L^-1 Pt (inversion.py:105,114, LAPACK dtrtrs with 3N right-hand sides).
//
// The matrix is padded to a multiple of NB = 128 with an identity tail, so every panel is full.
// Per panel:  potrf128 (one CTA, shared memory; also emits the inverse of the diagonal block and
// accumulates log det) -> panel solve as a GEMM with that inverse -> trailing update (lower tiles
// only) on the DMMA engine.  The triangular solve with 3N right-hand sides is left-looking: each
// 128-row block of V is one wide GEMM against all previous rows, then a GEMM with the block inverse.
#include "common.cuh"

#include <algorithm>
#include <utility>

constexpr int NB = 128;
constexpr int PLD = NB + 1;   // padded leading dimension in shared memory

// Factor one 128x128 diagonal block in place (lower), write its inverse to linv (row-major, upper zero).
//
// Register-tiled right-looking Cholesky: 256 threads as a 16 x 16 grid, thread (ty, tx) owns the 8 x 8
// cyclic sub-matrix {(ty + 16a, tx + 16b)} of the (symmetric) block in registers; per pivot k the owner
// publishes sqrt(d), the owners of column k publish the scaled column through shared memory, and every
// thread applies the rank-1 update to its registers (2 barriers per pivot, no shared-memory matrix traffic).
// The inverse of the factor (used for the panel solve and the blocked forward substitution) is then
// built column by column with two lanes per column.
__global__ void __launch_bounds__(256, 1) potrf128_kernel(double* __restrict__ Bm, long ldb, int k0, int Mtrue,
                                                          double* __restrict__ linv, double* __restrict__ logdet,
                                                          int* __restrict__ info) {
    extern __shared__ double S[];   // [NB][PLD]; lower = L, strict upper = X^T (inverse)
    __shared__ double colk[NB];
    __shared__ double xd[NB];
    __shared__ double red[NB];
    __shared__ double pivs;
    __shared__ int bad;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double* blk = Bm + (long)k0 * ldb + k0;
    double reg[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int r = ty + 16 * a, c = tx + 16 * b;
            reg[a][b] = (c <= r) ? blk[(long)r * ldb + c] : blk[(long)c * ldb + r];   // only the lower triangle is valid in memory
        }
    if (tid == 0) bad = 0;
    __syncthreads();
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        for (int kk = 0; kk < 16; ++kk) {
            const int k = kb * 16 + kk;
            if (ty == kk && tx == kk) {
                const double d = reg[kb][kb];
                if (!(d > 0.0) && bad == 0) bad = k0 + k + 1;
                pivs = sqrt(d);
            }
            __syncthreads();
            const double piv = pivs;
            if (tx == kk) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int r = ty + 16 * a;
                    if (r >= k) {
                        const double l = (r == k) ? piv : reg[a][kb] / piv;
                        colk[r] = l;
                        reg[a][kb] = l;
                    }
                }
            }
            __syncthreads();
            double lr[8], lc[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) lr[a] = colk[ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 8; ++b) lc[b] = colk[tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b)
                    if (ty + 16 * a > k && tx + 16 * b > k) reg[a][b] -= lr[a] * lc[b];
        }
    }
    // publish L (lower) to shared memory and to the matrix
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int r = ty + 16 * a, c = tx + 16 * b;
            if (c <= r) {
                S[r * PLD + c] = reg[a][b];
                blk[(long)r * ldb + c] = reg[a][b];
            }
        }
    __syncthreads();
    // X = L^-1, column c by the lane pair (2c, 2c+1); X^T goes to the strict upper triangle of S
    {
        const int c = tid >> 1, h = tid & 1;
        const int cmin = (tid & ~31) >> 1;          // smallest column handled by this warp: uniform loop bounds
        const double xc = 1.0 / S[c * PLD + c];
        if (h == 0) xd[c] = xc;
        for (int r = cmin + 1; r < NB; ++r) {
            double s0 = 0.0, s1 = 0.0;
            if (r > c) {
                int k = c + h;
                if (k == c && k < r) { s0 = S[r * PLD + c] * xc; k += 2; }
                for (; k + 2 < r; k += 4) {
                    s0 = fma(S[r * PLD + k], S[c * PLD + k], s0);
                    s1 = fma(S[r * PLD + k + 2], S[c * PLD + k + 2], s1);
                }
                for (; k < r; k += 2) s0 = fma(S[r * PLD + k], S[c * PLD + k], s0);
            }
            double tot = s0 + s1;
            tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            if (r > c && h == 0) S[c * PLD + r] = -tot / S[r * PLD + r];
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int r = e >> 7, c = e & (NB - 1);
        linv[e] = (c < r) ? S[c * PLD + r] : (c == r ? xd[r] : 0.0);
    }
    // log det: fixed-order tree reduction (deterministic)
    if (tid < NB) {
        const double l = S[tid * PLD + tid];
        red[tid] = (k0 + tid < Mtrue) ? log(l * l) : 0.0;     // inversion.py:108: log(diag(L)**2)
    }
    __syncthreads();
    for (int o = NB / 2; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        *logdet += red[0];
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

cudaError_t chol_forward_solve(const double* L, long ldl, int Mp, const CholWork& w, double* Pt, long ldp, int ncols_all,
                               double* tmp, cudaStream_t s, int tri) {
    const int nblk = Mp / NB;
    cudaError_t e;
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB;
        const int ncols = tri ? (ncols_all < k0 + NB ? ncols_all : k0 + NB) : ncols_all;
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

// Linv diagonal blocks <- inverses of the 128 x 128 diagonal blocks of L (everything else zero)
__global__ void linv_diag_init_kernel(const double* __restrict__ linv_blocks, double* __restrict__ Linv, long n) {
    const int kb = blockIdx.x;
    const double* src = linv_blocks + (long)kb * NB * NB;
    double* dst = Linv + ((long)kb * NB) * n + (long)kb * NB;
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) dst[(long)(e >> 7) * n + (e & (NB - 1))] = src[e];
}

cudaError_t chol_inverse(const double* L, long ldl, int Mp, const CholWork& w, double* Linv, double* Ltmp, cudaStream_t s, long* nlaunch) {
    cudaError_t e = cudaMemsetAsync(Linv, 0, (size_t)Mp * Mp * sizeof(double), s);
    if (e != cudaSuccess) return e;
    const int nblk = Mp / NB;
    linv_diag_init_kernel<<<nblk, 256, 0, s>>>(w.linv, Linv, Mp);
    if (nlaunch) *nlaunch += 1;
    // groups of already inverted diagonal blocks: (start row, rows); merged pairwise per level
    std::vector<std::pair<int, int>> groups;
    for (int kb = 0; kb < nblk; ++kb) groups.push_back({kb * NB, NB});
    while (groups.size() > 1) {
        std::vector<std::pair<int, int>> next;
        std::vector<gemm::Task> t1, t2;
        for (size_t g = 0; g + 1 < groups.size(); g += 2) {
            const int r0 = groups[g].first, n0 = groups[g].second, r1 = groups[g + 1].first, n1 = groups[g + 1].second;
            double* T = Ltmp + (long)r1 * Mp + r0;
            // T = L21 . X11          (B given as [K][N])
            t1.push_back(make_task(L + (long)r1 * ldl + r0, ldl, Linv + (long)r0 * Mp + r0, Mp, nullptr, 0, T, Mp, n1, n0, n0, 1.0, 0.0, 0));
            // X21 = - X22 . T
            t2.push_back(make_task(Linv + (long)r1 * Mp + r1, Mp, T, Mp, nullptr, 0, Linv + (long)r1 * Mp + r0, Mp, n1, n0, n1, -1.0, 0.0, 0));
            next.push_back({r0, n0 + n1});
        }
        if (groups.size() & 1) next.push_back(groups.back());
        for (int pass = 0; pass < 2; ++pass) {
            const std::vector<gemm::Task>& tt = pass ? t2 : t1;
            for (size_t i = 0; i < tt.size(); i += gemm::MAX_TASKS) {
                gemm::TaskBatch b;
                b.n = (int)std::min<size_t>(gemm::MAX_TASKS, tt.size() - i);
                for (int k = 0; k < b.n; ++k) b.t[k] = tt[i + k];
                e = gemm::launch(b, gemm::B_N, s);
                if (e != cudaSuccess) return e;
                if (nlaunch) *nlaunch += 1;
            }
        }
        groups.swap(next);
    }
    return cudaGetLastError();
}
