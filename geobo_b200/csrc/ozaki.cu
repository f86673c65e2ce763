// Error-free (Ozaki-type) int8 slice products on the tcgen05 tensor cores for the fused covariance
// assembly + projection  Pt = A . K  (the 95 % kernel of the path).
//
// tcgen05 has no fp64 kind and the path is too ill-conditioned for tf32/bf16, so the
// fp64 operands are split into S fixed-point digits of 7 / 8 bits:
//     x = 2^e * sum_q d_q 2^-(7+8q),   d_0 in [-64, 64] (signed), d_q>0 in [0, 255] (unsigned),
// with one exponent per sensor row of A and one per covariance table.  Digit products are accumulated EXACTLY
// in int32 by `tcgen05.mma.kind::i8` (SASS UTCIMMA), one TMEM accumulator per significance level
// l = qa + qb <= S-1 (S(S+1)/2 MMAs per K = 32 step), flushed every CHUNK contraction indices (overflow bound)
// into the fp64 result with an exact int64 recombination.  The only error is the truncation of the operands
// to 7 + 8(S-1) bits (S = 5: 2^-39 relative to the row / table scale) and of levels >= S.
//
// CTA = 13 warps, persistent (one CTA per SM): warps 0-3 epilogue (TMEM -> int64 -> fp64 read-modify-write of
// the Pt tile), warp 4 TMEM allocator + single-thread MMA issuer, warps 5-12 producers: the A digits stream in
// with 16-byte cp.async, the K digits are GENERATED: for 16 consecutive contraction voxels in one z-column the
// 16 digit bytes are one unaligned window of the (symmetric) stationary-covariance byte table, fetched with five
// aligned 32-bit loads + funnel shifts and stored straight into the UMMA canonical (K-major, no-swizzle) layout.
// Stage hand-off is mbarrier based (full: producers -> MMA, empty: tcgen05.commit -> producers).
#include "common.cuh"
#include "umma.cuh"

using namespace umma;

namespace ozaki {

constexpr int NPROD_WARPS = 8, NPROD = NPROD_WARPS * 32;
constexpr int THREADS = (4 + 1 + NPROD_WARPS) * 32;   // 416
constexpr int STAGES = 4;
constexpr int TABLE_PAD = 64;
__host__ __device__ constexpr long plane_stride(long ext) { return (ext + TABLE_PAD + 15) & ~15L; }   // keeps every digit plane 16-byte aligned

// ------------------------------------------------------------------------------------------------ digit extraction
// t in (-1/2, 1/2): digits of t + half an ulp of the last digit (round to nearest overall)
template <int S>
__device__ __forceinline__ void digits(double t, uint8_t (&d)[S]) {
    t += ldexp(1.0, -(7 + 8 * (S - 1) + 1));
    double x = t * 128.0;
    double f = floor(x);
    d[0] = (uint8_t)(int8_t)(int)f;
    double r = x - f;
#pragma unroll
    for (int q = 1; q < S; ++q) {
        x = r * 256.0;
        f = floor(x);
        d[q] = (uint8_t)(int)f;
        r = x - f;
    }
}

// exponent e with |x| <= 2^(e-1)  (so t = x / 2^e lies in [-1/2, 1/2])
__device__ __forceinline__ int scale_exp(double amax) {
    if (!(amax > 0.0) || !isfinite(amax)) return 0;
    int e;
    frexp(amax, &e);      // amax = m 2^e, m in [0.5, 1)
    return e + 1;
}

__global__ void row_absmax_kernel(const double* __restrict__ A, long rows, long cols, long ld, int* __restrict__ exps) {
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    double m = 0.0;
    for (long c = lane; c < cols; c += 32) m = fmax(m, fabs(A[row * ld + c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) exps[row] = scale_exp(m);
}

// A (rows x ld fp64, cols valid) -> S digit planes [q][rows][kp] (bytes; columns >= cols are zero)
template <int S>
__global__ void slice_rows_kernel(const double* __restrict__ A, long rows, long cols, long ld, const int* __restrict__ exps,
                                  uint8_t* __restrict__ out, long kp) {
    const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long row = blockIdx.y;
    if (col >= kp) return;
    uint8_t d[S];
    if (col < cols) {
        digits<S>(ldexp(A[row * ld + col], -exps[row]), d);
    } else {
#pragma unroll
        for (int q = 0; q < S; ++q) d[q] = 0;
    }
#pragma unroll
    for (int q = 0; q < S; ++q) out[((long)q * rows + row) * kp + col] = d[q];
}

__global__ void table_absmax_kernel(const double* __restrict__ tab, long ext, int* __restrict__ exps) {
    __shared__ double red[256];
    const double* t = tab + (long)blockIdx.x * ext;
    double m = 0.0;
    for (long e = threadIdx.x; e < ext; e += blockDim.x) m = fmax(m, fabs(t[e]));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) exps[blockIdx.x] = scale_exp(red[0]);
}

// table (ntab x ext fp64) -> byte planes [tab][q][ext + TABLE_PAD]
template <int S>
__global__ void slice_table_kernel(const double* __restrict__ tab, long ext, const int* __restrict__ exps, uint8_t* __restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tb = blockIdx.y;
    if (e >= plane_stride(ext)) return;
    uint8_t d[S];
    if (e < ext) {
        digits<S>(ldexp(tab[(long)tb * ext + e], -exps[tb]), d);
    } else {
#pragma unroll
        for (int q = 0; q < S; ++q) d[q] = 0;
    }
#pragma unroll
    for (int q = 0; q < S; ++q) out[((long)tb * S + q) * plane_stride(ext) + e] = d[q];
}

// ------------------------------------------------------------------------------------------------ main kernel
struct Params {
    const uint8_t* a8[2];     // [q][Ns][kp] digit planes of A_grav / A_magn
    const int* a_exp[2];      // [Ns]
    const uint8_t* t8;        // [9][S][plane_stride(ext)] digit planes of the covariance tables (only c = 0, 1 are used)
    const int* t_exp;         // [9]
    const int* L;             // [kp] extended-lattice ids
    double* Pt;               // [Mp][ldp]
    long ext, C0, kp, ldp, ncp;
    int Ns, ncol, c0, chunk;  // chunk: contraction indices per accumulator flush (multiple of 32)
    int n_stile, n_itile;
};

template <int S, int NT>
__global__ void __launch_bounds__(THREADS, 1) ozaki_project_kernel(const __grid_constant__ Params P) {
    constexpr int A_BYTES = S * 4096, B_SLICE = NT * 32, STAGE_BYTES = A_BYTES + S * B_SLICE;
    constexpr int TMEM_COLS = 512;
    static_assert(S * NT <= TMEM_COLS, "accumulators do not fit in TMEM");
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar, tempty_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 4) {
        tmem_alloc(&tmem_base_s, TMEM_COLS);
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], NPROD); mbar_init(&empty_bar[s], 1); }
            mbar_init(&tfull_bar, 1);
            mbar_init(&tempty_bar, 128);
            fence_barrier_init();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const long tiles_per_task = (long)P.n_stile * P.n_itile;
    const long ntiles = 6 * tiles_per_task;
    const int nchunk = (int)((P.kp + P.chunk - 1) / P.chunk);

    if (warp >= 5) {
        // =============================================================== producers
        const int pt = tid - 5 * 32;                 // 0 .. 255
        const bool gen = pt < 2 * NT;                // one (column n, k-half) unit per thread
        const int gn = pt >> 1, gkh = pt & 1;
        uint32_t it = 0;
        int prev_stage = -1;
        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int task = (int)(tile / tiles_per_task);            // c * 3 + r
            const long rem = tile % tiles_per_task;
            const int stile = (int)(rem / P.n_itile), itile = (int)(rem % P.n_itile);
            const int c = task / 3;
            const int s0 = stile * 128, i0 = itile * NT;
            const uint8_t* a8 = P.a8[c];
            const uint8_t* t8 = P.t8 + (long)task * S * plane_stride(P.ext);
            long li = 0;
            if (gen) li = (long)P.L[P.c0 + min(i0 + gn, P.ncol - 1)];
            for (int ch = 0; ch < nchunk; ++ch) {
                const long jbeg = (long)ch * P.chunk, jend = min(P.kp, jbeg + P.chunk);
                for (long j0 = jbeg; j0 < jend; j0 += 32, ++it) {
                    const int st = it % STAGES;
                    mbar_wait(&empty_bar[st], ((it / STAGES) & 1) ^ 1);
                    uint8_t* sa = smem + st * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    // ---- A digits: S x 128 rows x 2 halves of 16 bytes, cp.async
                    for (int e = pt; e < S * 256; e += NPROD) {
                        const int q = e >> 8, r2 = e & 255, row = r2 >> 1, kh = r2 & 1;
                        const int gr = min(s0 + row, P.Ns - 1);
                        const uint8_t* src = a8 + ((long)q * P.Ns + gr) * P.kp + j0 + kh * 16;
                        const uint32_t dst = smem_u32(sa + q * 4096 + core_offset(row, kh));
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
                    }
                    asm volatile("cp.async.commit_group;");
                    // ---- K digits: one unaligned 16-byte window of the byte table per (column, half, digit)
                    if (gen) {
                        const long off = P.C0 + (long)P.L[j0 + gkh * 16] - li;      // symmetric table: index L(j) - L(i) + C0
                        const uint32_t sh = (uint32_t)(off & 3) * 8;
                        const uint8_t* base = t8 + (off & ~3L);
#pragma unroll
                        for (int q = 0; q < S; ++q) {
                            const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (long)q * plane_stride(P.ext));
                            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3), w4 = __ldg(w + 4);
                            uint4 v;
                            v.x = __funnelshift_r(w0, w1, sh);
                            v.y = __funnelshift_r(w1, w2, sh);
                            v.z = __funnelshift_r(w2, w3, sh);
                            v.w = __funnelshift_r(w3, w4, sh);
                            *reinterpret_cast<uint4*>(sb + q * B_SLICE + core_offset(gn, gkh)) = v;
                        }
                    }
                    // ---- publish the PREVIOUS stage (its cp.async group has had a whole step to land)
                    if (prev_stage >= 0) {
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                        fence_proxy_async_smem();
                        mbar_arrive(&full_bar[prev_stage]);
                    }
                    prev_stage = st;
                }
            }
        }
        if (prev_stage >= 0) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            fence_proxy_async_smem();
            mbar_arrive(&full_bar[prev_stage]);
        }
    } else if (warp == 4) {
        // =============================================================== MMA issuer (one thread)
        if (lane == 0) {
            uint32_t it = 0, chunk_id = 0;
            for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int ch = 0; ch < nchunk; ++ch, ++chunk_id) {
                    const long jbeg = (long)ch * P.chunk, jend = min(P.kp, jbeg + P.chunk);
                    mbar_wait(&tempty_bar, (chunk_id & 1) ^ 1);     // epilogue has drained the accumulators
                    tc_fence_after();
                    bool first = true;
                    for (long j0 = jbeg; j0 < jend; j0 += 32, ++it) {
                        const int st = it % STAGES;
                        mbar_wait(&full_bar[st], (it / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + st * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
                        for (int lvl = 0; lvl < S; ++lvl) {
#pragma unroll
                            for (int qa = 0; qa <= lvl; ++qa) {
                                const int qb = lvl - qa;
                                const uint64_t ad = smem_desc(sa + qa * 4096, kLBO, kSBO);
                                const uint64_t bd = smem_desc(sb + qb * B_SLICE, kLBO, kSBO);
                                mma_i8(tmem_base + lvl * NT, ad, bd, idesc_i8(qa == 0, qb == 0, NT), (first && qa == 0) ? 0u : 1u);
                            }
                        }
                        first = false;
                        mma_commit(&empty_bar[st]);       // stage reusable once these MMAs have read it
                    }
                    mma_commit(&tfull_bar);               // accumulators of this chunk complete
                }
            }
        }
    } else {
        // =============================================================== epilogue (warps 0-3 <-> TMEM lanes 32 w .. 32 w + 31)
        uint32_t chunk_id = 0;
        const int row = warp * 32 + lane;
        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int task = (int)(tile / tiles_per_task);
            const long rem = tile % tiles_per_task;
            const int stile = (int)(rem / P.n_itile), itile = (int)(rem % P.n_itile);
            const int c = task / 3, r = task % 3;
            const int s = stile * 128 + row, i0 = itile * NT;
            const bool row_ok = s < P.Ns;
            // result = 2^(eA + eK - 14 - 8 (S-1)) * sum_l acc_l 2^(8 (S-1-l))
            const double scale = row_ok ? ldexp(1.0, P.a_exp[c][s] + P.t_exp[task] - 14 - 8 * (S - 1)) : 0.0;
            double* prow = P.Pt + ((long)c * P.Ns + (row_ok ? s : 0)) * P.ldp + (long)r * P.ncp + i0;
            for (int ch = 0; ch < nchunk; ++ch, ++chunk_id) {
                mbar_wait(&tfull_bar, chunk_id & 1);
                tc_fence_after();
                for (int n0 = 0; n0 < NT; n0 += 8) {
                    uint32_t v[S][8];
#pragma unroll
                    for (int lvl = 0; lvl < S; ++lvl) tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + lvl * NT + n0, v[lvl]);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            long long acc = (long long)(int)v[0][k];
#pragma unroll
                            for (int lvl = 1; lvl < S; ++lvl) acc = acc * 256 + (long long)(int)v[lvl][k];
                            const int col = i0 + n0 + k;
                            if (col < P.ncol) {
                                const double add = scale * (double)acc;
                                prow[n0 + k] = (ch == 0) ? add : prow[n0 + k] + add;
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int S, int NT>
static cudaError_t launch_one(const Params& P, int sm_count, cudaStream_t s) {
    constexpr int smem = STAGES * (S * 4096 + S * NT * 32);
    cudaError_t e = cudaFuncSetAttribute(ozaki_project_kernel<S, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    Params q = P;
    q.n_itile = (P.ncol + NT - 1) / NT;
    const long ntiles = 6L * q.n_stile * q.n_itile;
    const int grid = (int)(ntiles < (long)sm_count ? ntiles : (long)sm_count);
    ozaki_project_kernel<S, NT><<<grid, THREADS, smem, s>>>(q);
    return cudaGetLastError();
}

}  // namespace ozaki

// ------------------------------------------------------------------------------------------------ host entry points
int ozaki_tile_n(int slices) { return slices == 4 ? 128 : slices == 5 ? 96 : slices == 6 ? 80 : 0; }

cudaError_t ozaki_slice_sens(const double* A, long rows, long cols, long ld, int slices, int* exps, uint8_t* out, long kp, cudaStream_t s) {
    ozaki::row_absmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(A, rows, cols, ld, exps);
    dim3 grid((unsigned)((kp + 255) / 256), (unsigned)rows);
    switch (slices) {
        case 4: ozaki::slice_rows_kernel<4><<<grid, 256, 0, s>>>(A, rows, cols, ld, exps, out, kp); break;
        case 5: ozaki::slice_rows_kernel<5><<<grid, 256, 0, s>>>(A, rows, cols, ld, exps, out, kp); break;
        case 6: ozaki::slice_rows_kernel<6><<<grid, 256, 0, s>>>(A, rows, cols, ld, exps, out, kp); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t ozaki_slice_tables(const double* tables, long ext, int slices, int* exps, uint8_t* out, cudaStream_t s) {
    ozaki::table_absmax_kernel<<<9, 256, 0, s>>>(tables, ext, exps);
    dim3 grid((unsigned)((ozaki::plane_stride(ext) + 255) / 256), 9);
    switch (slices) {
        case 4: ozaki::slice_table_kernel<4><<<grid, 256, 0, s>>>(tables, ext, exps, out); break;
        case 5: ozaki::slice_table_kernel<5><<<grid, 256, 0, s>>>(tables, ext, exps, out); break;
        case 6: ozaki::slice_table_kernel<6><<<grid, 256, 0, s>>>(tables, ext, exps, out); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

long ozaki_table_bytes(long ext, int slices) { return 9L * slices * ozaki::plane_stride(ext); }

cudaError_t ozaki_project(const OzakiArgs& a, int slices, int sm_count, cudaStream_t s) {
    ozaki::Params P;
    P.a8[0] = a.a8[0]; P.a8[1] = a.a8[1]; P.a_exp[0] = a.a_exp[0]; P.a_exp[1] = a.a_exp[1];
    P.t8 = a.t8; P.t_exp = a.t_exp; P.L = a.L; P.Pt = a.Pt;
    P.ext = a.ext; P.C0 = a.C0; P.kp = a.kp; P.ldp = a.ldp; P.ncp = a.ncp;
    P.Ns = a.Ns; P.ncol = a.ncol; P.c0 = a.c0;
    P.chunk = slices <= 5 ? 8192 : 4096;       // int32 overflow bound per significance level (see header comment)
    P.n_stile = (a.Ns + 127) / 128;
    P.n_itile = 0;
    switch (slices) {
        case 4: return ozaki::launch_one<4, 128>(P, sm_count, s);
        case 5: return ozaki::launch_one<5, 96>(P, sm_count, s);
        case 6: return ozaki::launch_one<6, 80>(P, sm_count, s);
    }
    return cudaErrorInvalidValue;
}
