// fp64 GEMM engine on the DMMA tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4 on sm_100a).
//
// tcgen05 has no fp64 kind, so every fp64 contraction of the path runs here:
//   * the fused covariance-assembly + projection  Pt = A . K   (B operand GENERATED in shared
//     memory from the stationary-covariance table -- the 3N x 3N matrix never exists in HBM),
//   * AkA = A . Pt^T, the Cholesky panel / trailing updates, and the blocked triangular solve.
//
// CTA tile 128 x 128 x BK, 8 warps (2 x 4), warp tile 64 x 32 = 8 x 4 DMMA fragments,
// multi-stage cp.async (LDGSTS) pipeline, 16-byte copies for stored operands and 8-byte
// gathers straight from the L2-resident table for the generated operand.
#include <stdlib.h>

#include "common.cuh"

namespace gemm {

constexpr int BM = 128, BN = 128, THREADS = 256;
constexpr int LDBN_S = BN + 4;   // 132 doubles: conflict-free 8-byte fragment loads in the N-contiguous layouts

template <int BK>
struct Cfg {
    static constexpr int LDA_S = BK + 4;   // (BK + 4) % 16 == 4: conflict-free 8-byte fragment loads
    static constexpr int A_STAGE = BM * LDA_S;
    static constexpr int B_STAGE = (BN * LDA_S > BK * LDBN_S) ? BN * LDA_S : BK * LDBN_S;
    static constexpr int smem_bytes(int stages) { return stages * (A_STAGE + B_STAGE) * (int)sizeof(double); }
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// LATE = 1: the loads of the slab that enters the ring are issued after the first k4 step of the
// current slab, so the tensor pipe restarts right after the barrier.
template <int MODE, int BK, int STAGES, int LATE>
__global__ void __launch_bounds__(THREADS, 1) gemm_f64_kernel(const __grid_constant__ TaskBatch batch) {
    using C = Cfg<BK>;
    constexpr int LDA_S = C::LDA_S, A_STAGE = C::A_STAGE, B_STAGE = C::B_STAGE;
    constexpr int CH = BK / 2;                      // 16-byte chunks per operand row
    constexpr int NA = BM * CH / THREADS;           // A chunks per thread (4 for BK=16, 8 for BK=32)
    constexpr int NBN = BK * (BN / 2) / THREADS;    // B_N chunks per thread
    constexpr int NG = BK * BN / THREADS;           // B_GEN elements per thread
    const Task& T = batch.t[blockIdx.z];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= T.M || n0 >= T.N) return;
    if (T.lower && n0 > m0 + BM - 1) return;

    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int nk = T.K / BK;

    // ---- per-thread copy coordinates
    const double* a_src[NA];
    int a_dst[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) {
        const int c = tid + THREADS * q;
        const int row = c / CH, kc = c % CH;
        const int gr = min(m0 + row, T.M - 1);
        a_src[q] = T.A + (long)gr * T.lda + kc * 2;
        a_dst[q] = row * LDA_S + kc * 2;
    }
    constexpr int NB = (MODE == B_T) ? NA : (MODE == B_N ? NBN : 1);
    const double* b_src[NB];
    int b_dst[NB];
    int li = 0;                 // B_GEN: lattice id of this thread's output column
    int lj_next[NG];            // B_GEN: lattice ids of the contraction rows this thread gathers (next slab to issue)
    const int gen_n = tid & (BN - 1), gen_j = tid >> 7;   // 2 j-rows per pass
    if (MODE == B_T) {
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int c = tid + THREADS * q;
            const int row = c / CH, kc = c % CH;
            const int gr = min(n0 + row, T.N - 1);
            b_src[q] = T.B + (long)gr * T.ldb + kc * 2;
            b_dst[q] = row * LDA_S + kc * 2;
        }
    } else if (MODE == B_N) {
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int c = tid + THREADS * q;
            const int krow = c / (BN / 2), nc = c % (BN / 2);
            const long gn = min((long)n0 + nc * 2, T.ldb - 2);
            b_src[q] = T.B + (long)krow * T.ldb + gn;
            b_dst[q] = krow * LDBN_S + nc * 2;
        }
    } else {
        li = T.Lrow[min(n0 + gen_n, T.N - 1)];
#pragma unroll
        for (int q = 0; q < NG; ++q) lj_next[q] = T.Lcol[min(gen_j + 2 * q, T.K - 1)];
    }

    auto issue = [&](int kt, int stage) {
        double* as = As + stage * A_STAGE;
        double* bs = Bs + stage * B_STAGE;
        const long koff = (long)kt * BK;
#pragma unroll
        for (int q = 0; q < NA; ++q) cp_async16(as + a_dst[q], a_src[q] + koff);
        if (MODE == B_T) {
#pragma unroll
            for (int q = 0; q < NB; ++q) cp_async16(bs + b_dst[q], b_src[q] + koff);
        } else if (MODE == B_N) {
#pragma unroll
            for (int q = 0; q < NB; ++q) cp_async16(bs + b_dst[q], b_src[q] + koff * T.ldb);
        } else {
#pragma unroll
            for (int q = 0; q < NG; ++q) cp_async8(bs + (gen_j + 2 * q) * LDBN_S + gen_n, T.B + (li - lj_next[q]));
            const int kn = (kt + 1) * BK;   // prefetch the lattice ids of the slab issued next
#pragma unroll
            for (int q = 0; q < NG; ++q) lj_next[q] = T.Lcol[min(kn + gen_j + 2 * q, T.K - 1)];
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) issue(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int kn = kt + STAGES - 1;
        if (!LATE) {
            if (kn < nk) issue(kn, kn % STAGES);
            cp_async_commit();
        }
        const double* as = As + (kt % STAGES) * A_STAGE + (wm * 64 + g) * LDA_S + t4;
        const double* bs = Bs + (kt % STAGES) * B_STAGE;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            double a[8], b[4];
#pragma unroll
            for (int mf = 0; mf < 8; ++mf) a[mf] = as[mf * 8 * LDA_S + k4 * 4];
            if (MODE == B_T) {
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) b[nf] = bs[(wn * 32 + nf * 8 + g) * LDA_S + k4 * 4 + t4];
            } else {
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) b[nf] = bs[(k4 * 4 + t4) * LDBN_S + wn * 32 + nf * 8 + g];
            }
#pragma unroll
            for (int mf = 0; mf < 8; ++mf)
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) dmma(acc[mf][nf][0], acc[mf][nf][1], a[mf], b[nf]);
            if (LATE && k4 == 0) {
                if (kn < nk) issue(kn, kn % STAGES);
                cp_async_commit();
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: C = alpha * acc + beta * Cin
    const double alpha = T.alpha, beta = T.beta;
    const bool vec_ok = ((T.ldc & 1) == 0) && ((((uintptr_t)T.C) & 15) == 0) &&
                        (beta == 0.0 || (((T.ldcin & 1) == 0) && ((((uintptr_t)T.Cin) & 15) == 0)));
#pragma unroll
    for (int mf = 0; mf < 8; ++mf) {
        const int m = m0 + wm * 64 + mf * 8 + g;
        if (m >= T.M) continue;
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) {
            const int n = n0 + wn * 32 + nf * 8 + 2 * t4;
            if (n >= T.N) continue;
            double v0 = alpha * acc[mf][nf][0], v1 = alpha * acc[mf][nf][1];
            double* cp = T.C + (long)m * T.ldc + n;
            if (n + 1 < T.N && vec_ok) {
                if (beta != 0.0) {
                    const double2 ci = *reinterpret_cast<const double2*>(T.Cin + (long)m * T.ldcin + n);
                    v0 += beta * ci.x;
                    v1 += beta * ci.y;
                }
                *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
            } else {
                if (beta != 0.0) v0 += beta * T.Cin[(long)m * T.ldcin + n];
                cp[0] = v0;
                if (n + 1 < T.N) {
                    if (beta != 0.0) v1 += beta * T.Cin[(long)m * T.ldcin + n + 1];
                    cp[1] = v1;
                }
            }
        }
    }
}

// ---- variant table (GEOBO_B200_GEMM="BK,STAGES,LATE" selects one at context creation; default below)
typedef void (*kernel_fn)(const TaskBatch);
struct Variant {
    int bk, stages, late, smem;
    kernel_fn fn[3];
};
#define GB_VARIANT(BK, ST, LT)                                                                               \
    {BK, ST, LT, Cfg<BK>::smem_bytes(ST),                                                                    \
     {gemm_f64_kernel<B_T, BK, ST, LT>, gemm_f64_kernel<B_N, BK, ST, LT>, gemm_f64_kernel<B_GEN, BK, ST, LT>}}
static const Variant kVariants[] = {
    GB_VARIANT(16, 3, 0), GB_VARIANT(16, 4, 0), GB_VARIANT(16, 3, 1), GB_VARIANT(16, 4, 1),
    GB_VARIANT(32, 3, 0), GB_VARIANT(32, 3, 1), GB_VARIANT(32, 2, 0),
};
static int g_variant = 5;   // BK = 32, 3 stages, late issue: fastest in the round-1 sweep (profiles/README.md)

cudaError_t init() {
    const char* env = getenv("GEOBO_B200_GEMM");
    if (env) {
        int bk = 0, st = 0, lt = 0;
        if (sscanf(env, "%d,%d,%d", &bk, &st, &lt) == 3) {
            for (size_t i = 0; i < sizeof(kVariants) / sizeof(kVariants[0]); ++i)
                if (kVariants[i].bk == bk && kVariants[i].stages == st && kVariants[i].late == lt) g_variant = (int)i;
        }
    }
    const Variant& v = kVariants[g_variant];
    for (int m = 0; m < 3; ++m) {
        cudaError_t e = cudaFuncSetAttribute(v.fn[m], cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int tile_k() { return kVariants[g_variant].bk; }

cudaError_t launch(const TaskBatch& batch, BMode mode, cudaStream_t stream) {
    const Variant& v = kVariants[g_variant];
    int maxM = 0, maxN = 0;
    for (int i = 0; i < batch.n; ++i) {
        if (batch.t[i].K % v.bk != 0) return cudaErrorInvalidValue;
        maxM = max(maxM, batch.t[i].M);
        maxN = max(maxN, batch.t[i].N);
    }
    if (batch.n == 0 || maxM == 0 || maxN == 0) return cudaSuccess;
    dim3 grid((maxN + BN - 1) / BN, (maxM + BM - 1) / BM, batch.n);
    if (grid.y > 65535) return cudaErrorInvalidValue;
    v.fn[(int)mode]<<<grid, THREADS, v.smem, stream>>>(batch);
    return cudaGetLastError();
}

}  // namespace gemm
