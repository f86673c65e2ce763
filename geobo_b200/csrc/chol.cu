// Blocked fp64 Cholesky of AkA (inversion.py:100, LAPACK dpotrf) and the blocked forward solve
// V = L^-1 Pt (inversion.py:105,114, LAPACK dtrtrs with 3N right-hand sides).
//
// The matrix is padded to a multiple of NB = 128 with an identity tail, so every panel is full.
// Per panel:  potrf128 (one CTA, shared memory; also emits the inverse of the diagonal block and
// accumulates log det) -> panel solve as a GEMM with that inverse -> trailing update (lower tiles
// only) on the DMMA engine.  The triangular solve with 3N right-hand sides is left-looking: each
// 128-row block of V is one wide GEMM against all previous rows, then a GEMM with the block inverse.
#include <stdlib.h>

#include "common.cuh"
#include "comm.h"

#include <algorithm>
#include <utility>

constexpr int NB = 128;
constexpr int PLD = NB + 1;   // padded leading dimension in shared memory

// ---- potrf128 v2 ---------------------------------------------------------------------------------------------
// Latency-optimised one-CTA Cholesky of a 128 x 128 diagonal block + its inverse.
//  * factor: right-looking, register tiled (256 threads as 16 x 16, thread (ty, tx) owns the cyclic elements
//    (ty + 16 a, tx + 16 b), LOWER triangle only); every thread also tracks the running diagonal of its 8 columns, so
//    the 16 owners of pivot column k compute rsqrt(d_k) locally and ONE barrier per pivot suffices (the scaled column
//    goes through a double-buffered shared array);
//  * inverse: 16 x 16 diagonal blocks by one warp each (forward substitution in registers), then recursive doubling
//    X21 = -X22 (L21 X11) for block sizes 16, 32, 64 with register-tiled shared-memory products.
template <int SZ, int TI, int TJ>
__device__ __forceinline__ void inv_level(double* __restrict__ S, double* __restrict__ T, int tid) {
    // pairs of SZ x SZ diagonal blocks: (o, o + SZ), o = 2 SZ p.  X (strict lower) is stored transposed in the upper triangle
    // of S: X[r][c] at S[c * PLD + r]; the diagonal of X is in xd (passed through T + 64*64).
    constexpr int NPAIR = NB / (2 * SZ);
    constexpr int TPP = 256 / NPAIR;                 // threads per pair
    constexpr int GI = SZ / TI, GJ = SZ / TJ;        // tile grid
    static_assert(GI * GJ == TPP, "tile grid must match the threads of a pair");
    const double* xd = T + 64 * 64;
    const int p = tid / TPP, t = tid % TPP, ti = (t / GJ) * TI, tj = (t % GJ) * TJ;
    const int o1 = 2 * SZ * p, o2 = o1 + SZ;
    double* Tp = T + p * SZ * SZ;
    // T = L21 . X11   (X11 lower triangular incl. diagonal): T[i][j] = sum_{k >= j} L21[i][k] X11[k][j]
    {
        double acc[TI][TJ];
#pragma unroll
        for (int i = 0; i < TI; ++i)
#pragma unroll
            for (int j = 0; j < TJ; ++j) acc[i][j] = 0.0;
        for (int k = tj; k < SZ; ++k) {
            double l[TI], x[TJ];
#pragma unroll
            for (int i = 0; i < TI; ++i) l[i] = S[(o2 + ti + i) * PLD + o1 + k];
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
                const int c = tj + j;
                x[j] = (k > c) ? S[(o1 + c) * PLD + o1 + k] : (k == c ? xd[o1 + c] : 0.0);
            }
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) acc[i][j] = fma(l[i], x[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < TI; ++i)
#pragma unroll
            for (int j = 0; j < TJ; ++j) Tp[(ti + i) * SZ + tj + j] = acc[i][j];
    }
    __syncthreads();
    // X21 = - X22 . T   (X22 lower triangular): X21[i][j] = - sum_{k <= i} X22[i][k] T[k][j]
    {
        double acc[TI][TJ];
#pragma unroll
        for (int i = 0; i < TI; ++i)
#pragma unroll
            for (int j = 0; j < TJ; ++j) acc[i][j] = 0.0;
        for (int k = 0; k < ti + TI; ++k) {
            double x[TI], tt[TJ];
#pragma unroll
            for (int i = 0; i < TI; ++i) {
                const int r = ti + i;
                x[i] = (k < r) ? S[(o2 + k) * PLD + o2 + r] : (k == r ? xd[o2 + r] : 0.0);
            }
#pragma unroll
            for (int j = 0; j < TJ; ++j) tt[j] = Tp[k * SZ + tj + j];
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) acc[i][j] = fma(x[i], tt[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < TI; ++i)
#pragma unroll
            for (int j = 0; j < TJ; ++j) S[(o1 + tj + j) * PLD + o2 + ti + i] = -acc[i][j];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 1) potrf128_kernel(double* __restrict__ Bm, long ldb, int k0, int Mtrue,
                                                          double* __restrict__ linv, double* __restrict__ logdet,
                                                          int* __restrict__ info) {
    extern __shared__ double S[];           // [NB][PLD]: lower = L, strict upper = X^T; then T [64*64] + xd [NB]
    double* T = S + NB * PLD;
    double* xd = T + 64 * 64;
    __shared__ double colk[2][NB];
    __shared__ double dsave[NB];
    __shared__ double red[NB];
    __shared__ int bad;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double* blk = Bm + (long)k0 * ldb + k0;
    double reg[8][8];                       // only b <= a is used (lower triangle of the cyclic tiling)
    double dd[8];                           // running diagonal of columns tx + 16 b
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            const int r = ty + 16 * a, c = tx + 16 * b;
            reg[a][b] = (c <= r) ? blk[(long)r * ldb + c] : 0.0;
        }
#pragma unroll
    for (int b = 0; b < 8; ++b) { const int c = tx + 16 * b; dd[b] = blk[(long)c * ldb + c]; }
    if (tid == 0) bad = 0;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
#pragma unroll 1
        for (int kk = 0; kk < 16; ++kk) {
            const int k = kb * 16 + kk;
            double* ck = colk[k & 1];
            if (tx == kk) {
                const double d = dd[kb];
                const double rs = rsqrt(d);
                if (ty == kk) {
                    if (!(d > 0.0) && bad == 0) bad = k0 + k + 1;
                    dsave[k] = d;
                    xd[k] = rs;                                  // 1 / L_kk = X_kk
                    ck[k] = d * rs;                              // L_kk
                }
#pragma unroll
                for (int a = kb; a < 8; ++a) {
                    const int r = ty + 16 * a;
                    if (r > k) {
                        const double l = reg[a][kb] * rs;
                        ck[r] = l;
                        reg[a][kb] = l;
                    }
                }
            }
            __syncthreads();
            double lr[8], lc[8];
#pragma unroll
            for (int a = kb; a < 8; ++a) lr[a] = ck[ty + 16 * a];
#pragma unroll
            for (int b = kb; b < 8; ++b) lc[b] = ck[tx + 16 * b];
#pragma unroll
            for (int b = kb; b < 8; ++b)
                if (tx + 16 * b > k) dd[b] = fma(-lc[b], lc[b], dd[b]);
#pragma unroll
            for (int a = kb; a < 8; ++a)
#pragma unroll
                for (int b = kb; b <= a; ++b)
                    if (ty + 16 * a > k && tx + 16 * b > k) reg[a][b] = fma(-lr[a], lc[b], reg[a][b]);
        }
    }
    __syncthreads();
    // publish L (lower incl. diagonal) to shared memory and to the matrix
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            const int r = ty + 16 * a, c = tx + 16 * b;
            if (c < r) {
                S[r * PLD + c] = reg[a][b];
                blk[(long)r * ldb + c] = reg[a][b];
            }
        }
    if (tid < NB) {
        const double l = dsave[tid] * xd[tid];
        S[tid * PLD + tid] = l;
        blk[(long)tid * ldb + tid] = l;
        red[tid] = (k0 + tid < Mtrue) ? log(dsave[tid]) : 0.0;   // inversion.py:108: log(diag(L)**2) = log(d_k)
    }
    __syncthreads();
    // ---- inverse, level 0: the eight 16 x 16 diagonal blocks, one warp each, lane j < 16 owns column j
    {
        const int w = tid >> 5, lane = tid & 31, o = 16 * w;
        if (lane < 16) {
            double x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = 0.0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (i == lane) x[i] = xd[o + i];
                else if (i > lane) {
                    double s = 0.0;
#pragma unroll
                    for (int kq = 0; kq < 16; ++kq)
                        if (kq < i) s = fma(S[(o + i) * PLD + o + kq], x[kq], s);     // x[kq] = 0 for kq < lane
                    x[i] = -s * xd[o + i];
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i > lane) S[(o + lane) * PLD + o + i] = x[i];
        }
    }
    __syncthreads();
    inv_level<16, 2, 2>(S, T, tid);
    inv_level<32, 2, 4>(S, T, tid);
    inv_level<64, 4, 4>(S, T, tid);
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int r = e >> 7, c = e & (NB - 1);
        linv[e] = (c < r) ? S[c * PLD + r] : (c == r ? xd[r] : 0.0);
    }
    // log det: fixed-order tree reduction (deterministic)
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

// GEOBO_B200_CHOL_OUTER = panels per outer block.  Default 4 = two-level blocking: the panels of a 512-wide block are
// factored left-looking and the trailing matrix is updated once per block with K = 512, which keeps the DMMA pipe 74 % busy
// (profiles/r2_chol_outer4_cfg3.ncu-rep; four K = 128 updates: 45 %, every C tile read and written four times).
// 1 = plain right-looking.  Read on every call (tests switch it in-process).
static int chol_outer_panels() {
    int outer = 4;
    if (const char* ev = getenv("GEOBO_B200_CHOL_OUTER")) { outer = atoi(ev); if (outer < 1 || outer > 16) outer = 1; }
    return outer;
}

// GEOBO_B200_CHOL_LOOKAHEAD=1 (default 0): the trailing update of an outer block is split into the columns of the next outer
// block (main stream) and the rest (side stream), so that the sequential potrf128 / panel-solve chain of the next block runs
// concurrently with the bulk of the update instead of after it.  Same arithmetic per element, same order of the updates of
// every element (all (b) parts are stream ordered, (a) of block j waits for (b) of block j - 1): results are bit-identical.
struct CholLookahead {
    bool on = false, ready = false, b_pending = false;
    cudaStream_t side = nullptr;
    cudaEvent_t panels = nullptr, b_done = nullptr;
};
// per device (a process may hold contexts on several devices; streams, events and function attributes are per device)
static CholLookahead la_dev[64];
static bool potrf_attr_set[64];

static cudaError_t potrf_attr(int smem) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && potrf_attr_set[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(potrf128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess && dev >= 0 && dev < 64) potrf_attr_set[dev] = true;
    return e;
}

static cudaError_t chol_lookahead_setup(CholLookahead& la) {
    la.on = false;
    la.b_pending = false;
    if (const char* ev = getenv("GEOBO_B200_CHOL_LOOKAHEAD")) la.on = atoi(ev) != 0;
    if (!la.on || la.ready) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&la.side, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&la.panels, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&la.b_done, cudaEventDisableTiming)) != cudaSuccess) return e;
    la.ready = true;
    return cudaSuccess;
}

cudaError_t chol_factor(double* Bm, long ldb, int Mp, int Mtrue, const CholWork& w, cudaStream_t s) {
    const int smem = (NB * PLD + 64 * 64 + NB) * (int)sizeof(double);
    cudaError_t e;
    if ((e = potrf_attr(smem)) != cudaSuccess) return e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    CholLookahead& la = la_dev[dev & 63];
    const int nblk = Mp / NB;
    const int outer = chol_outer_panels();
    if ((e = chol_lookahead_setup(la)) != cudaSuccess) return e;
    for (int ob = 0; ob < nblk; ob += outer) {
        const int o0 = ob * NB, oe = (ob + outer < nblk ? ob + outer : nblk), o1 = oe * NB;
        for (int kb = ob; kb < oe; ++kb) {
            const int k0 = kb * NB;
            double* linv = w.linv + (long)kb * NB * NB;
            if (k0 > o0) {
                // left-looking inside the outer block: A[k0:, kb] -= L[k0:, o0:k0] . L[kb rows, o0:k0]^T
                double* c = Bm + (long)k0 * ldb + k0;
                gemm::TaskBatch b0;
                b0.n = 1;
                b0.t[0] = make_task(Bm + (long)k0 * ldb + o0, ldb, Bm + (long)k0 * ldb + o0, ldb, c, ldb, c, ldb, Mp - k0, NB, k0 - o0, -1.0, 1.0, 0);
                e = gemm::launch(b0, gemm::B_T, s);
                if (e != cudaSuccess) return e;
            }
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
        }
        const int rest = Mp - o1;
        if (rest <= 0) break;
        // trailing update once per outer block: A22 -= L[o1:, o0:o1] . L[o1:, o0:o1]^T  (lower tiles only), K = o1 - o0
        double* lp = Bm + (long)o1 * ldb + o0;
        double* trail = Bm + (long)o1 * ldb + o1;
        gemm::TaskBatch b2;
        b2.n = 1;
        if (!la.on) {
            b2.t[0] = make_task(lp, ldb, lp, ldb, trail, ldb, trail, ldb, rest, rest, o1 - o0, -1.0, 1.0, 1);
            e = gemm::launch(b2, gemm::B_T, s);
            if (e != cudaSuccess) return e;
            continue;
        }
        // look-ahead: (a) the columns of the NEXT outer block on the main stream, so that its panels can start at once;
        // (b) everything to the right of them on the side stream, concurrently with those panels.  Both (a) of this block and
        // (b) of the previous one update the next block's columns, and (b) needs this block's finished panels:
        //   main: wait (b) of the previous block -> (a);   side: wait the panels of this block -> (b)   [(b)s are stream ordered]
        const int wdt = (outer * NB < rest) ? outer * NB : rest;
        if ((e = cudaEventRecord(la.panels, s)) != cudaSuccess) return e;
        if (la.b_pending && (e = cudaStreamWaitEvent(s, la.b_done, 0)) != cudaSuccess) return e;
        b2.t[0] = make_task(lp, ldb, lp, ldb, trail, ldb, trail, ldb, rest, wdt, o1 - o0, -1.0, 1.0, 1);
        if ((e = gemm::launch(b2, gemm::B_T, s)) != cudaSuccess) return e;
        la.b_pending = false;
        if (rest > wdt) {
            if ((e = cudaStreamWaitEvent(la.side, la.panels, 0)) != cudaSuccess) return e;
            double* lp2 = lp + (long)wdt * ldb;
            double* trail2 = trail + (long)wdt * ldb + wdt;
            b2.t[0] = make_task(lp2, ldb, lp2, ldb, trail2, ldb, trail2, ldb, rest - wdt, rest - wdt, o1 - o0, -1.0, 1.0, 1);
            if ((e = gemm::launch(b2, gemm::B_T, la.side)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(la.b_done, la.side)) != cudaSuccess) return e;
            la.b_pending = true;
        }
    }
    if (la.on && la.b_pending) {
        if ((e = cudaStreamWaitEvent(s, la.b_done, 0)) != cudaSuccess) return e;
        la.b_pending = false;
    }
    // PROFILING AID (GEOBO_B200_CHOL_EMULATE="nranks,rank", off by default): after the factorisation, launch once more the trailing
    // updates that rank `rank` of `nranks` issues in chol_factor_dist (block-cyclic column blocks, one batched K = outer x 128 launch
    // per outer block, operands from the finished factor) into a scratch matrix, so that `ncu` -- which cannot be attached to a
    // multi-rank job on this pool -- can capture the kernel with exactly the grid and operand shapes of an 8-GPU run on ONE GPU.
    // The results of these launches are discarded.
    if (const char* ev = getenv("GEOBO_B200_CHOL_EMULATE")) {
        int nr = 0, me = 0;
        if (sscanf(ev, "%d,%d", &nr, &me) == 2 && nr > 1 && me >= 0 && me < nr && outer > 1) {
            double* scratch = nullptr;
            if ((e = cudaMalloc(&scratch, (size_t)Mp * ldb * sizeof(double))) != cudaSuccess) return e;
            cudaMemsetAsync(scratch, 0, (size_t)Mp * ldb * sizeof(double), s);
            for (int ob = 0; ob < nblk; ob += outer) {
                const int o0 = ob * NB, oe = (ob + outer < nblk ? ob + outer : nblk), o1 = oe * NB;
                gemm::TaskBatch b2;
                b2.n = 0;
                for (int jb = oe; jb < nblk; ++jb) {
                    if (jb % nr != me) continue;
                    const int j0 = jb * NB;
                    const double* lj = Bm + (long)j0 * ldb + o0;
                    double* c = scratch + (long)j0 * ldb + j0;
                    b2.t[b2.n++] = make_task(lj, ldb, lj, ldb, c, ldb, c, ldb, Mp - j0, NB, o1 - o0, -1.0, 1.0, 0);
                    if (b2.n == gemm::MAX_TASKS) {
                        if ((e = gemm::launch(b2, gemm::B_T, s)) != cudaSuccess) { cudaFree(scratch); return e; }
                        b2.n = 0;
                    }
                }
                if (b2.n && (e = gemm::launch(b2, gemm::B_T, s)) != cudaSuccess) { cudaFree(scratch); return e; }
            }
            cudaStreamSynchronize(s);
            cudaFree(scratch);
        }
    }
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------- distributed factorisation
// 1-D block-cyclic Cholesky over the ranks of the context: column block kb (128 columns) is owned by
// rank kb % nranks.  Per panel: the owner factors the diagonal block (potrf128) and solves the panel, packs
// [L[k0:, kb] | inverse of the diagonal block] into a contiguous staging buffer, ONE ncclBroadcast over NVLink ships it,
// every rank unpacks it into its replica of L and updates only the trailing column blocks it owns (one batched GEMM
// launch on the fp64 tensor pipe, operand = the packed panel).  L and the block inverses end up replicated, which is what
// the triangular inverse / refinement stages need.  log det and the first non-positive pivot are combined by all-reduce.
__global__ void panel_pack_kernel(const double* __restrict__ Bm, long ldb, int k0, int rows, const double* __restrict__ linv,
                                  double* __restrict__ stage) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long np = (long)rows * NB;
    if (e < np) stage[e] = Bm[(long)(k0 + e / NB) * ldb + k0 + (e % NB)];
    else if (e < np + NB * NB) stage[e] = linv[e - np];
}

__global__ void panel_unpack_kernel(double* __restrict__ Bm, long ldb, int k0, int rows, double* __restrict__ linv,
                                    const double* __restrict__ stage) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long np = (long)rows * NB;
    if (e < np) Bm[(long)(k0 + e / NB) * ldb + k0 + (e % NB)] = stage[e];
    else if (e < np + NB * NB) linv[e - np] = stage[e];
}

// per-panel log det / pivot info -> doubles for the sum all-reduce (only the owner of a panel has non-zero entries)
__global__ void chol_pan_to_double_kernel(const int* __restrict__ paninfo, int nblk, double* __restrict__ pan) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nblk) pan[nblk + k] = (double)paninfo[k];
}

// total log det in panel order (same summation order as the single-GPU factorisation) and the first non-positive pivot
__global__ void chol_pan_finalize_kernel(const double* __restrict__ pan, int nblk, double* __restrict__ logdet, int* __restrict__ info) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    int first = 0;
    for (int k = 0; k < nblk; ++k) {
        s += pan[k];
        const int v = (int)pan[nblk + k];
        if (v > 0 && first == 0) first = v;
    }
    *logdet += s;
    if (first && *info == 0) *info = first;
}

// stage: [Mp * 128 + 128 * 128] doubles, pan: [2 * Mp / 128] doubles, paninfo: [Mp / 128] ints (all device scratch)
int chol_factor_dist(gb_ctx* ctx, double* Bm, long ldb, int Mp, int Mtrue, const CholWork& w, double* stage, double* pan, int* paninfo,
                     cudaStream_t s) {
    const int smem = (NB * PLD + 64 * 64 + NB) * (int)sizeof(double);
    GB_CUDA(ctx, potrf_attr(smem));
    const int nblk = Mp / NB, nr = ctx->nranks, me = ctx->rank;
    GB_CUDA(ctx, cudaMemsetAsync(pan, 0, (size_t)2 * nblk * sizeof(double), s));
    GB_CUDA(ctx, cudaMemsetAsync(paninfo, 0, (size_t)nblk * sizeof(int), s));
    // Two-level blocking (GEOBO_B200_CHOL_OUTER = panels per outer block > 1): inside an outer block the owner of a panel
    // first applies the earlier panels of that block to its own column block (left-looking, operands from its replica of
    // L, which the unpack below keeps complete), and the trailing column blocks are updated once per outer block with
    // K = outer x 128 instead of once per panel with K = 128.  Same flops per rank, fatter GEMMs.
    const int outer = chol_outer_panels();
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB, rows = Mp - k0, owner = kb % nr;
        const int ob = kb / outer * outer, o0 = ob * NB, oe = ob + outer < nblk ? ob + outer : nblk, o1 = oe * NB;
        double* linv = w.linv + (long)kb * NB * NB;
        const long count = (long)rows * NB + NB * NB;
        if (me == owner) {
            if (outer > 1 && k0 > o0) {
                double* c = Bm + (long)k0 * ldb + k0;
                gemm::TaskBatch b0;
                b0.n = 1;
                b0.t[0] = make_task(Bm + (long)k0 * ldb + o0, ldb, Bm + (long)k0 * ldb + o0, ldb, c, ldb, c, ldb, Mp - k0, NB, k0 - o0, -1.0, 1.0, 0);
                GB_CUDA(ctx, gemm::launch(b0, gemm::B_T, s));
            }
            potrf128_kernel<<<1, 256, smem, s>>>(Bm, ldb, k0, Mtrue, linv, pan + kb, paninfo + kb);
            GB_CUDA(ctx, cudaGetLastError());
            if (rows > NB) {
                double* panel = Bm + (long)(k0 + NB) * ldb + k0;
                gemm::TaskBatch b1;
                b1.n = 1;
                b1.t[0] = make_task(panel, ldb, linv, NB, nullptr, 0, panel, ldb, rows - NB, NB, NB, 1.0, 0.0, 0);
                GB_CUDA(ctx, gemm::launch(b1, gemm::B_T, s));
            }
            panel_pack_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(Bm, ldb, k0, rows, linv, stage);
            GB_CUDA(ctx, cudaGetLastError());
        }
        GB_TRY(comm_broadcast_f64(ctx, stage, (size_t)count, owner));
        if (me != owner) {
            panel_unpack_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(Bm, ldb, k0, rows, linv, stage);
            GB_CUDA(ctx, cudaGetLastError());
        }
        gemm::TaskBatch b2;
        b2.n = 0;
        if (outer == 1) {
            // trailing update of the column blocks this rank owns: A[j0:, jb] -= L[j0:, kb] . L[jb rows, kb]^T, operands from the packed panel
            for (int jb = kb + 1; jb < nblk; ++jb) {
                if (jb % nr != me) continue;
                const int j0 = jb * NB;
                const double* pj = stage + (long)(j0 - k0) * NB;          // rows j0.. of the packed panel (ld = NB)
                double* c = Bm + (long)j0 * ldb + j0;
                b2.t[b2.n++] = make_task(pj, NB, pj, NB, c, ldb, c, ldb, Mp - j0, NB, NB, -1.0, 1.0, 0);
                if (b2.n == gemm::MAX_TASKS) {
                    GB_CUDA(ctx, gemm::launch(b2, gemm::B_T, s));
                    b2.n = 0;
                }
            }
        } else if (kb == oe - 1) {
            // last panel of the outer block: A[j0:, jb] -= L[j0:, o0:o1] . L[jb rows, o0:o1]^T for the owned column blocks behind it,
            // operands from this rank's replica of L (K = o1 - o0)
            for (int jb = oe; jb < nblk; ++jb) {
                if (jb % nr != me) continue;
                const int j0 = jb * NB;
                const double* lj = Bm + (long)j0 * ldb + o0;
                double* c = Bm + (long)j0 * ldb + j0;
                b2.t[b2.n++] = make_task(lj, ldb, lj, ldb, c, ldb, c, ldb, Mp - j0, NB, o1 - o0, -1.0, 1.0, 0);
                if (b2.n == gemm::MAX_TASKS) {
                    GB_CUDA(ctx, gemm::launch(b2, gemm::B_T, s));
                    b2.n = 0;
                }
            }
        }
        if (b2.n) GB_CUDA(ctx, gemm::launch(b2, gemm::B_T, s));
    }
    chol_pan_to_double_kernel<<<(nblk + 127) / 128, 128, 0, s>>>(paninfo, nblk, pan);
    GB_CUDA(ctx, cudaGetLastError());
    GB_TRY(comm_allreduce_sum_f64(ctx, pan, (size_t)2 * nblk));
    chol_pan_finalize_kernel<<<1, 32, 0, s>>>(pan, nblk, w.logdet, w.info);
    GB_CUDA(ctx, cudaGetLastError());
    return GB_OK;
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
