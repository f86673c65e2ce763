// Error-free int8 digit-slice GEMM on tcgen05 / TMEM with BOTH operands streamed from pre-tiled digit blocks:
//     C[m, n] = sum_k A[m, k] * B[n, k]        (both K-major)
// used for the two dense products after the Cholesky factorisation of the path:
//   * V = L^-1 . Pt  (inversion.py:114) as  Linv (lower triangular, explicit) x Pt, with the epilogue reducing
//     sum_m V[m, n]^2 per column -- the only thing inversion.py:117,238 needs of V (var = amp - colsumsq(V)) --
//     so V is never written to memory;
//   * AkA = A3 . Pt^T (inversion.py:96) block products (epilogue stores fp64).
// Same arithmetic as ozaki.cu: operands are fixed-point (one exponent per operand row), split into S balanced
// 8-bit digits, digit products accumulate exactly in int32 in TMEM (one accumulator per significance level),
// recombined exactly in int64 and scaled in fp64.
//
// Operand layout (written by the slicing kernels below): [row tile][k step][digit][rows x 32 B in the UMMA
// canonical K-major no-swizzle layout], rows = 128 for A and NT for B, so one pipeline stage of an operand is ONE
// contiguous block fetched by a single bulk async copy (UBLKCP) that completes on the stage's mbarrier.
// CTA = 4 epilogue warps + 1 MMA warp + 1 copy warp, persistent, one CTA per SM.
#include "common.cuh"
#include "umma.cuh"
#include "ozaki.cuh"

using namespace umma;

namespace ozaki {

__device__ __forceinline__ double pow2i(int e) { return __hiloint2double((1023 + e) << 20, 0); }   // 2^e, -1022 <= e <= 1023

// ------------------------------------------------------------------------------------------------ slicing kernels
// exponent per column of a row-major matrix (column absmax over `rows` rows)
__global__ void col_absmax_kernel(const double* __restrict__ X, long rows, long cols, long ld, int* __restrict__ exps) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= cols) return;
    double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
    long r = 0;
    for (; r + 3 < rows; r += 4) {
        m0 = fmax(m0, fabs(X[r * ld + n]));
        m1 = fmax(m1, fabs(X[(r + 1) * ld + n]));
        m2 = fmax(m2, fabs(X[(r + 2) * ld + n]));
        m3 = fmax(m3, fabs(X[(r + 3) * ld + n]));
    }
    for (; r < rows; ++r) m0 = fmax(m0, fabs(X[r * ld + n]));
    exps[n] = scale_exp(fmax(fmax(m0, m1), fmax(m2, m3)));
}

// Transposed slicing of X (rows x ld, `cols` valid columns): column n of X becomes operand ROW n with K = row index.
// One thread per column walks down the column 16 rows at a time (coalesced across the warp), emits S 16-byte digit
// chunks per step into the pre-tiled B layout [n / TR][k step][digit][TR x 32 B], and -- fused, because this is the
// one pass that reads every element of Pt after the Cholesky -- accumulates  mu[n] = sum_m X[m, n] * alpha[m]
// (posterior mean, inversion.py:115 as Pt^T . (L^-T L^-1 y)).
template <int S>
__global__ void __launch_bounds__(128) slice_cols_mean_kernel(const double* __restrict__ X, long rows, long cols, long ld,
                                                              const int* __restrict__ exps, const double* __restrict__ alpha,
                                                              uint8_t* __restrict__ out, int TR, long ksteps,
                                                              double* __restrict__ mu, long ncp, long ncol) {
    __shared__ double al[512];
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = n < cols;
    const long nn = live ? n : 0;
    const int e = exps[nn];
    const long tile = nn / TR;
    const uint32_t rit = (uint32_t)(nn % TR);
    uint8_t* obase = out + tile * ksteps * S * (long)(TR * 32);
    double acc0 = 0.0, acc1 = 0.0;
    for (long m0 = 0; m0 < rows; m0 += 512) {
        __syncthreads();
        for (int q = threadIdx.x; q < 512; q += blockDim.x) al[q] = (m0 + q < rows) ? alpha[m0 + q] : 0.0;
        __syncthreads();
        const int lim = (int)min(512L, rows - m0);      // rows is a multiple of 32
        for (int g = 0; g < lim; g += 16) {
            uint32_t pk[S][4];
#pragma unroll
            for (int q = 0; q < S; ++q) pk[q][0] = pk[q][1] = pk[q][2] = pk[q][3] = 0u;
            double v[16];
#pragma unroll
            for (int b = 0; b < 16; ++b) v[b] = X[(m0 + g + b) * ld + nn];
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                if (b & 1) acc1 = fma(v[b], al[g + b], acc1); else acc0 = fma(v[b], al[g + b], acc0);
                uint8_t d[S];
                digits<S>(ldexp(v[b], -e), d);
#pragma unroll
                for (int q = 0; q < S; ++q) pk[q][b >> 2] |= (uint32_t)d[q] << (8 * (b & 3));
            }
            if (live) {
                const long gg = (m0 + g) >> 4;           // 16-row group
                const long ks = gg >> 1;
                uint8_t* o = obase + ks * S * (long)(TR * 32) + core_offset(rit, (uint32_t)(gg & 1));
#pragma unroll
                for (int q = 0; q < S; ++q)
                    *reinterpret_cast<uint4*>(o + (long)q * (TR * 32)) = make_uint4(pk[q][0], pk[q][1], pk[q][2], pk[q][3]);
            }
        }
    }
    if (live && mu) {
        const long r = n / ncp, col = n % ncp;
        if (col < ncol) mu[r * ncol + col] = acc0 + acc1;
    }
}

// var[r][col] = amp - sum over row tiles of the column sums of squares (inversion.py:117,238; Q10: diag(K) = amp)
__global__ void var_finalize_kernel(const double* __restrict__ partial, int n_mtile, long ldpart, long ncp, long ncol, double amp,
                                    double* __restrict__ var) {
    const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (col >= ncol) return;
    double s = 0.0;
    for (int t = 0; t < n_mtile; ++t) s += partial[(long)t * ldpart + r * ncp + col];    // fixed order: deterministic
    var[r * ncol + col] = amp - s;
}

// ------------------------------------------------------------------------------------------------ GEMM kernel
enum Epi { EPI_STORE = 0, EPI_SUMSQ = 1 };

struct GemmParams {
    const uint8_t* a8;      // [m tile][a_ksteps][S][4096]
    const int* a_exp;       // [M]
    const uint8_t* b8;      // [n tile][b_ksteps][S][NT*32]
    const int* b_exp;       // [N]
    double* C;              // EPI_STORE: [M][ldc]
    double* partial;        // EPI_SUMSQ: [n_mtile][ldc]
    double* scratch;        // EPI_SUMSQ with more than one accumulator flush: [grid][128][NT] running fp64 sums of V
    long ldc;
    int M, N;               // valid rows / columns
    int a_ksteps, b_ksteps; // K steps per tile row of the stored operands
    int a_k0, b_k0;         // first K step of the contraction inside the stored operands
    int ksteps;             // contraction length in K = 32 steps
    int chunk_steps;        // accumulator flush interval (EPI_STORE only; EPI_SUMSQ needs ksteps <= chunk_steps)
    int tri;                // 1: A is lower triangular with 128-row tiles -> tile row mt only needs K steps < 4 (mt + 1)
    int lower;              // 1: skip output tiles strictly above the diagonal
    int n_mtile, n_ntile, gm;   // gm: row tiles per L2 group (tile order: group, column tile, row tile in group)
};

template <int NT>
__device__ __forceinline__ bool tile_coords(const GemmParams& P, long tile, int& mt, int& nt) {
    const int ng = P.n_mtile / P.gm;
    const long full = (long)ng * P.gm * P.n_ntile;
    if (tile < full) {
        const long per = (long)P.gm * P.n_ntile;
        const int g = (int)(tile / per);
        const long r = tile % per;
        nt = (int)(r / P.gm);
        mt = g * P.gm + (int)(r % P.gm);
    } else {
        const int rem = P.n_mtile - ng * P.gm;
        const long r = tile - full;
        nt = (int)(r / rem);
        mt = ng * P.gm + (int)(r % rem);
    }
    return !(P.lower && (long)nt * NT > (long)mt * 128 + 127);
}

template <int S, int NT, int EPI>
__global__ void __launch_bounds__(192, 1) ozaki_gemm_kernel(const __grid_constant__ GemmParams P) {
    constexpr int STAGES = Cfg<S>::STAGES;
    constexpr int A_BYTES = S * 4096, B_SLICE = NT * 32, B_BYTES = S * B_SLICE, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = 512;
    constexpr int G = (256 / NT) < S ? (256 / NT) : S;
    static_assert(S * NT <= TMEM_COLS, "accumulators do not fit in TMEM");
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar, tempty_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ double part[EPI == EPI_SUMSQ ? 4 : 1][EPI == EPI_SUMSQ ? NT : 1];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 4) {
        tmem_alloc(&tmem_base_s, TMEM_COLS);
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            mbar_init(&tfull_bar, 1);
            mbar_init(&tempty_bar, 128);
            fence_barrier_init();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const long ntiles = (long)P.n_mtile * P.n_ntile;

    if (warp == 5) {
        // =============================================================== copy warp (one elected lane)
        if (lane == 0) {
            uint32_t it = 0;
            for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                int mt, nt;
                if (!tile_coords<NT>(P, tile, mt, nt)) continue;
                const int kend = P.tri ? min(P.ksteps, 4 * (mt + 1)) : P.ksteps;
                const uint8_t* a_src = P.a8 + ((size_t)mt * P.a_ksteps + P.a_k0) * A_BYTES;
                const uint8_t* b_src = P.b8 + ((size_t)nt * P.b_ksteps + P.b_k0) * B_BYTES;
                for (int ks = 0; ks < kend; ++ks, ++it) {
                    const int st = (int)(it % STAGES);
                    mbar_wait(&empty_bar[st], ((it / STAGES) & 1) ^ 1);
                    uint8_t* sa = smem + st * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[st], STAGE_BYTES);
                    bulk_g2s(sa, a_src + (size_t)ks * A_BYTES, A_BYTES, &full_bar[st]);
                    bulk_g2s(sa + A_BYTES, b_src + (size_t)ks * B_BYTES, B_BYTES, &full_bar[st]);
                }
            }
        }
    } else if (warp == 4) {
        // =============================================================== MMA issuer: the whole warp runs the (uniform) loop with
        // running ring state (no div / mod / descriptor rebuilds per step), one elected lane issues the tcgen05 instructions
        {
            uint32_t chunk_id = 0, ph = 0;
            int st = 0;
            const uint64_t adesc0 = smem_desc(smem_u32(smem), kLBO, kSBO);
            uint64_t adesc = adesc0;
            uint64_t* fb = full_bar;
            uint64_t* eb = empty_bar;
            for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                int mt, nt;
                if (!tile_coords<NT>(P, tile, mt, nt)) continue;
                const int kend = P.tri ? min(P.ksteps, 4 * (mt + 1)) : P.ksteps;
                for (int k0 = 0; k0 < kend; k0 += P.chunk_steps, ++chunk_id) {
                    const int k1 = min(kend, k0 + P.chunk_steps);
                    mbar_wait(&tempty_bar, (chunk_id & 1) ^ 1);
                    tc_fence_after();
                    for (int ks = k0; ks < k1; ++ks) {
                        mbar_wait(fb, ph);
                        tc_fence_after();
                        const uint32_t acc = ks == k0 ? 0u : 1u;
                        if (elect_one()) {
#pragma unroll
                            for (int qa = 0; qa < S; ++qa) {
#pragma unroll
                                for (int qb0 = 0; qb0 < S - qa; qb0 += G) {
                                    const int g = (S - qa - qb0) < G ? (S - qa - qb0) : G;
                                    mma_i8(tmem_base + (qa + qb0) * NT, adesc + (uint64_t)((qa * 4096) >> 4),
                                           adesc + (uint64_t)((A_BYTES + qb0 * B_SLICE) >> 4), idesc_i8(1, 1, g * NT), qa == 0 ? acc : 1u);
                                }
                            }
                            mma_commit(eb);
                        }
                        ++fb; ++eb; adesc += (uint64_t)(STAGE_BYTES >> 4);
                        if (++st == STAGES) { st = 0; ph ^= 1; fb = full_bar; eb = empty_bar; adesc = adesc0; }
                    }
                    if (elect_one()) mma_commit(&tfull_bar);
                }
            }
        }
    } else {
        // =============================================================== epilogue (warps 0-3 <-> TMEM lanes 32 w .. 32 w + 31)
        uint32_t chunk_id = 0;
        const int row = warp * 32 + lane;
        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int mt, nt;
            if (!tile_coords<NT>(P, tile, mt, nt)) continue;
            const int kend = P.tri ? min(P.ksteps, 4 * (mt + 1)) : P.ksteps;
            const int m = mt * 128 + row, n0t = nt * NT;
            const bool row_ok = m < P.M;
            // result = 2^(eA + eB - 14 - 8 (S-1)) * sum_l acc_l 2^(8 (S-1-l))
            const double rscale = row_ok ? pow2i(P.a_exp[m] - 14 - 8 * (S - 1)) : 0.0;
            for (int k0 = 0; k0 < kend; k0 += P.chunk_steps, ++chunk_id) {
                mbar_wait(&tfull_bar, chunk_id & 1);
                tc_fence_after();
                for (int n0 = 0; n0 < NT; n0 += 8) {
                    uint32_t v[S][8];
#pragma unroll
                    for (int lvl = 0; lvl < S; ++lvl) tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + lvl * NT + n0, v[lvl]);
                    tmem_ld_wait();
                    double val[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        long long acc = (long long)(int)v[0][k];
#pragma unroll
                        for (int lvl = 1; lvl < S; ++lvl) acc = acc * 256 + (long long)(int)v[lvl][k];
                        const int col = n0t + n0 + k;
                        val[k] = col < P.N ? rscale * pow2i(P.b_exp[col]) * (double)acc : 0.0;
                    }
                    if (EPI == EPI_STORE) {
                        if (row_ok) {
                            double* crow = P.C + (long)m * P.ldc + n0t + n0;
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (n0t + n0 + k < P.N) crow[k] = (k0 == 0) ? val[k] : crow[k] + val[k];
                        }
                    } else {
                        const bool last = k0 + P.chunk_steps >= kend;
                        if (kend > P.chunk_steps) {          // V accumulates over several flushes before it is squared
                            double* sc = P.scratch + ((size_t)blockIdx.x * 128 + row) * NT + n0;
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                if (k0 > 0) val[k] += sc[k];
                                if (!last) sc[k] = val[k];
                            }
                        }
                        if (last) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            double sq = val[k] * val[k];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                            if (lane == 0) part[warp][n0 + k] = sq;
                        }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar);        // accumulators drained: the next tile's MMAs may start
            }
            if (EPI == EPI_SUMSQ) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tid < NT && n0t + tid < P.N)
                    P.partial[(long)mt * P.ldc + n0t + tid] = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int S, int NT, int EPI>
static cudaError_t launch_gemm(GemmParams P, int sm_count, cudaStream_t s) {
    constexpr int smem = Cfg<S>::STAGES * (S * 4096 + S * NT * 32);
    cudaError_t e = cudaFuncSetAttribute(ozaki_gemm_kernel<S, NT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    P.n_mtile = (P.M + 127) / 128;
    P.n_ntile = (P.N + NT - 1) / NT;
    // row tiles per group: keep the A digits of one group (<= 48 MB) resident in L2 while the B tiles stream past
    const long a_tile_bytes = (long)P.ksteps * S * 4096;
    long gm = (48L << 20) / (a_tile_bytes > 0 ? a_tile_bytes : 1);
    P.gm = (int)(gm < 1 ? 1 : (gm > P.n_mtile ? P.n_mtile : gm));
    const long ntiles = (long)P.n_mtile * P.n_ntile;
    const int grid = (int)(ntiles < (long)sm_count ? ntiles : (long)sm_count);
    ozaki_gemm_kernel<S, NT, EPI><<<grid, 192, smem, s>>>(P);
    return cudaGetLastError();
}

}  // namespace ozaki

// ------------------------------------------------------------------------------------------------ host entry points
long ozaki_cols_bytes(long cols, long kp, int slices) {
    const long nt = ozaki_tile_n(slices);
    return ((cols + nt - 1) / nt) * (kp / 32) * (long)slices * nt * 32;
}

// Pt (rows x ld fp64) -> per-column exponents + transposed digit blocks (B operand, K = row index), fused with mu = Pt^T alpha
cudaError_t ozaki_slice_cols_mean(const double* X, long rows, long cols, long ld, int slices, int* exps, uint8_t* out,
                                  const double* alpha, double* mu, long ncp, long ncol, cudaStream_t s) {
    ozaki::col_absmax_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, s>>>(X, rows, cols, ld, exps);
    const long ksteps = rows / 32;
    const unsigned grid = (unsigned)((cols + 127) / 128);
    const int TR = ozaki_tile_n(slices);
    switch (slices) {
        case 4: ozaki::slice_cols_mean_kernel<4><<<grid, 128, 0, s>>>(X, rows, cols, ld, exps, alpha, out, TR, ksteps, mu, ncp, ncol); break;
        case 5: ozaki::slice_cols_mean_kernel<5><<<grid, 128, 0, s>>>(X, rows, cols, ld, exps, alpha, out, TR, ksteps, mu, ncp, ncol); break;
        case 6: ozaki::slice_cols_mean_kernel<6><<<grid, 128, 0, s>>>(X, rows, cols, ld, exps, alpha, out, TR, ksteps, mu, ncp, ncol); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// partial[mt][n] = sum over the 128 rows of tile mt of (Linv . Pt)[m, n]^2
long ozaki_colsumsq_scratch_bytes(int Mp, int slices, int sm_count) {
    return Mp > ozaki_chunk() ? (long)sm_count * 128 * ozaki_tile_n(slices) * 8 : 0;
}

cudaError_t ozaki_colsumsq_tri(const uint8_t* a8, const int* a_exp, const uint8_t* b8, const int* b_exp, int Mp, long ncols, int slices,
                               double* partial, double* scratch, int sm_count, cudaStream_t s, long ldpart) {
    ozaki::GemmParams P;
    memset(&P, 0, sizeof P);
    P.a8 = a8; P.a_exp = a_exp; P.b8 = b8; P.b_exp = b_exp; P.partial = partial; P.ldc = ldpart > 0 ? ldpart : ncols;
    P.M = Mp; P.N = (int)ncols;
    P.a_ksteps = P.b_ksteps = P.ksteps = Mp / 32;
    P.chunk_steps = ozaki_chunk() / 32;
    P.scratch = scratch;
    if (P.ksteps > P.chunk_steps && !scratch) return cudaErrorInvalidValue;
    P.tri = 1;
    switch (slices) {
        case 4: return ozaki::launch_gemm<4, ozaki::Cfg<4>::NT, ozaki::EPI_SUMSQ>(P, sm_count, s);
        case 5: return ozaki::launch_gemm<5, ozaki::Cfg<5>::NT, ozaki::EPI_SUMSQ>(P, sm_count, s);
        case 6: return ozaki::launch_gemm<6, ozaki::Cfg<6>::NT, ozaki::EPI_SUMSQ>(P, sm_count, s);
    }
    return cudaErrorInvalidValue;
}

// C[m, n] = sum_k A[m, k] B[n, k] over K steps [a_k0, a_k0 + ksteps) of A and [b_k0, ...) of B; lower: skip tiles above the diagonal.
// The B operand is tiled in rows of ozaki_tile_np(slices) (the sensitivities' digit blocks of the projection are reused).
cudaError_t ozaki_gemm_store(const uint8_t* a8, const int* a_exp, int a_ksteps, int a_k0, const uint8_t* b8, const int* b_exp, int b_ksteps,
                             int b_k0, int ksteps, int M, int N, double* C, long ldc, int lower, int slices, int sm_count, cudaStream_t s) {
    ozaki::GemmParams P;
    memset(&P, 0, sizeof P);
    P.a8 = a8; P.a_exp = a_exp; P.b8 = b8; P.b_exp = b_exp; P.C = C; P.ldc = ldc;
    P.M = M; P.N = N; P.a_ksteps = a_ksteps; P.b_ksteps = b_ksteps; P.a_k0 = a_k0; P.b_k0 = b_k0; P.ksteps = ksteps;
    P.chunk_steps = ozaki_chunk() / 32;
    P.lower = lower;
    switch (slices) {
        case 4: return ozaki::launch_gemm<4, ozaki::Cfg<4>::NTP, ozaki::EPI_STORE>(P, sm_count, s);
        case 5: return ozaki::launch_gemm<5, ozaki::Cfg<5>::NTP, ozaki::EPI_STORE>(P, sm_count, s);
        case 6: return ozaki::launch_gemm<6, ozaki::Cfg<6>::NTP, ozaki::EPI_STORE>(P, sm_count, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t ozaki_var_finalize(const double* partial, int n_mtile, long ldpart, long ncp, long ncol, double amp, double* var, cudaStream_t s, int nr) {
    dim3 grid((unsigned)((ncol + 255) / 256), (unsigned)nr);
    ozaki::var_finalize_kernel<<<grid, 256, 0, s>>>(partial, n_mtile, ldpart, ncp, ncol, amp, var);
    return cudaGetLastError();
}
