// fp64 GEMM engine on the DMMA tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4 on sm_100a).
//
// tcgen05 has no fp64 kind, so every fp64 contraction of the path runs here:
//   * the fused covariance-assembly + projection  Pt = A . K   (B operand GENERATED in shared
//     memory from the stationary-covariance table -- the 3N x 3N matrix never exists in HBM),
//   * AkA = A . Pt^T, the Cholesky panel / trailing updates, and the blocked triangular solve.
//
// CTA tile 128 x 128 x 16, 8 warps (2 x 4), warp tile 64 x 32 = 8 x 4 DMMA fragments,
// 3-stage cp.async (LDGSTS) pipeline, 16-byte copies for stored operands and 8-byte
// gathers straight from the L2-resident table for the generated operand.
#include "common.cuh"

namespace gemm {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, STAGES = 3;
constexpr int LDA_S = BK + 4;    // 20 doubles: conflict-free 8-byte fragment loads
constexpr int LDBN_S = BN + 4;   // 132 doubles
constexpr int A_STAGE = BM * LDA_S;                // doubles
constexpr int B_STAGE = (BN * LDA_S > BK * LDBN_S) ? BN * LDA_S : BK * LDBN_S;
constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(double);

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

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) gemm_f64_kernel(const __grid_constant__ TaskBatch batch) {
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
    // A (and B_T): 1024 16-byte chunks, 4 per thread: row = c / 8, kc = c % 8
    const double* a_src[4];
    int a_dst[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int c = tid + THREADS * q;
        int row = c >> 3, kc = c & 7;
        int gr = min(m0 + row, T.M - 1);
        a_src[q] = T.A + (long)gr * T.lda + kc * 2;
        a_dst[q] = row * LDA_S + kc * 2;
    }
    const double* b_src[4];
    int b_dst[4];
    int li = 0;           // B_GEN: lattice id of this thread's output column
    int lj_next[8];       // B_GEN: lattice ids of the 8 contraction rows this thread gathers (next slab to issue)
    const int gen_n = tid & (BN - 1), gen_j = tid >> 7;   // 2 j-rows per pass, 8 passes
    if (MODE == B_T) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int c = tid + THREADS * q;
            int row = c >> 3, kc = c & 7;
            int gr = min(n0 + row, T.N - 1);
            b_src[q] = T.B + (long)gr * T.ldb + kc * 2;
            b_dst[q] = row * LDA_S + kc * 2;
        }
    } else if (MODE == B_N) {
        // tile [16][128]: 1024 chunks: krow = c / 64, nc = c % 64
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int c = tid + THREADS * q;
            int krow = c >> 6, nc = c & 63;
            long gn = min((long)n0 + nc * 2, T.ldb - 2);
            b_src[q] = T.B + (long)krow * T.ldb + gn;
            b_dst[q] = krow * LDBN_S + nc * 2;
        }
    } else {
        li = T.Lrow[min(n0 + gen_n, T.N - 1)];
#pragma unroll
        for (int q = 0; q < 8; ++q) lj_next[q] = T.Lcol[min(gen_j + 2 * q, T.K - 1)];
    }

    auto issue = [&](int kt, int stage) {
        double* as = As + stage * A_STAGE;
        double* bs = Bs + stage * B_STAGE;
        const long koff = (long)kt * BK;
#pragma unroll
        for (int q = 0; q < 4; ++q) cp_async16(as + a_dst[q], a_src[q] + koff);
        if (MODE == B_T) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cp_async16(bs + b_dst[q], b_src[q] + koff);
        } else if (MODE == B_N) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cp_async16(bs + b_dst[q], b_src[q] + koff * T.ldb);
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                cp_async8(bs + (gen_j + 2 * q) * LDBN_S + gen_n, T.B + (li - lj_next[q]));
            // prefetch lattice ids of the slab issued next
            const int kn = (kt + 1) * BK;
#pragma unroll
            for (int q = 0; q < 8; ++q) lj_next[q] = T.Lcol[min(kn + gen_j + 2 * q, T.K - 1)];
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // ---- prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) issue(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int kn = kt + STAGES - 1;
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

cudaError_t init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(gemm_f64_kernel<B_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_f64_kernel<B_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_f64_kernel<B_GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    return e;
}

cudaError_t launch(const TaskBatch& batch, BMode mode, cudaStream_t stream) {
    int maxM = 0, maxN = 0;
    for (int i = 0; i < batch.n; ++i) {
        if (batch.t[i].K % BK != 0) return cudaErrorInvalidValue;
        maxM = max(maxM, batch.t[i].M);
        maxN = max(maxN, batch.t[i].N);
    }
    if (batch.n == 0 || maxM == 0 || maxN == 0) return cudaSuccess;
    dim3 grid((maxN + BN - 1) / BN, (maxM + BM - 1) / BM, batch.n);
    if (grid.y > 65535) return cudaErrorInvalidValue;
    switch (mode) {
        case B_T: gemm_f64_kernel<B_T><<<grid, THREADS, SMEM_BYTES, stream>>>(batch); break;
        case B_N: gemm_f64_kernel<B_N><<<grid, THREADS, SMEM_BYTES, stream>>>(batch); break;
        case B_GEN: gemm_f64_kernel<B_GEN><<<grid, THREADS, SMEM_BYTES, stream>>>(batch); break;
    }
    return cudaGetLastError();
}

}  // namespace gemm
