// Hardware check of the tcgen05 conventions in geobo_b200/csrc/umma.cuh: K-major no-swizzle smem descriptors
// (LBO/SBO), kind::i8 instruction descriptor (signed/unsigned operands, N = 96 / 128), accumulate flag, TMEM
// lane/column addressing with tcgen05.ld.32x32b, tcgen05.commit -> mbarrier.  One CTA, exact int32 comparison.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/umma_test tools/umma_test.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../geobo_b200/csrc/umma.cuh"

using namespace umma;

// A: [128][K] bytes, B: [N][K] bytes (K-major), D: [128][2*N] int32 (two accumulators: D0 = A.B^T, D1 = 2 * A.B^T via accumulate)
__global__ void __launch_bounds__(160, 1) umma_test_kernel(const uint8_t* A, const uint8_t* B, int* D, int N, int nk, int a_signed, int b_signed) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = nk * 32;
    uint8_t* sA = smem;                       // nk blocks of 4096 B
    uint8_t* sB = smem + nk * 4096;           // nk blocks of N*32 B
    if (warp == 4) {
        tmem_alloc(&tmem_base_s, 512);
        if (tid == 128) { mbar_init(&bar, 1); fence_barrier_init(); }
    } else {
        for (int e = tid; e < 128 * nk * 2; e += 128) {          // 16-byte pieces of A
            const int row = e / (nk * 2), p = e % (nk * 2), ks = p >> 1, kh = p & 1;
            *reinterpret_cast<uint4*>(sA + ks * 4096 + core_offset(row, kh)) = *reinterpret_cast<const uint4*>(A + (size_t)row * K + ks * 32 + kh * 16);
        }
        for (int e = tid; e < N * nk * 2; e += 128) {
            const int row = e / (nk * 2), p = e % (nk * 2), ks = p >> 1, kh = p & 1;
            *reinterpret_cast<uint4*>(sB + ks * N * 32 + core_offset(row, kh)) = *reinterpret_cast<const uint4*>(B + (size_t)row * K + ks * 32 + kh * 16);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 128) {
        const uint32_t idesc = idesc_i8(a_signed, b_signed, N);
        for (int pass = 0; pass < 3; ++pass) {        // pass 0 -> D0; pass 1, 2 -> D1 (second pass accumulates on top of the first)
            const uint32_t d = tmem_base + (pass == 0 ? 0 : N);
            for (int ks = 0; ks < nk; ++ks) {
                const uint64_t ad = smem_desc(smem_u32(sA + ks * 4096), kLBO, kSBO);
                const uint64_t bd = smem_desc(smem_u32(sB + ks * N * 32), kLBO, kSBO);
                mma_i8(d, ad, bd, idesc, (ks > 0 || pass == 2) ? 1u : 0u);
            }
        }
        mma_commit(&bar);
    }
    if (warp < 4) {
        mbar_wait(&bar, 0);
        tc_fence_after();
        const int row = warp * 32 + (tid & 31);
        for (int c0 = 0; c0 < 2 * N; c0 += 8) {
            uint32_t v[8];
            tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int q = 0; q < 8; ++q) D[(size_t)row * 2 * N + c0 + q] = (int)v[q];
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// A from tensor memory (tcgen05.st by the four warps that own the lanes), B from shared memory; D0 = A.B^T over nk K-steps
__global__ void __launch_bounds__(160, 1) umma_ts_test_kernel(const uint8_t* A, const uint8_t* B, int* D, int N, int nk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = nk * 32;
    uint8_t* sB = smem;
    if (warp == 4) {
        tmem_alloc(&tmem_base_s, 512);
        if (tid == 128) { mbar_init(&bar, 1); fence_barrier_init(); }
    } else {
        for (int e = tid; e < N * nk * 2; e += 128) {
            const int row = e / (nk * 2), p = e % (nk * 2), ks = p >> 1, kh = p & 1;
            *reinterpret_cast<uint4*>(sB + ks * N * 32 + core_offset(row, kh)) = *reinterpret_cast<const uint4*>(B + (size_t)row * K + ks * 32 + kh * 16);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_col = 256;                      // A operand columns: 8 per K step, after the accumulator
    if (warp < 4) {
        const int row = warp * 32 + (tid & 31);
        for (int ks = 0; ks < nk; ++ks) {
            uint32_t v[8];
            for (int q = 0; q < 8; ++q) v[q] = *reinterpret_cast<const uint32_t*>(A + (size_t)row * K + ks * 32 + q * 4);
            tmem_st8(tmem_base + ((uint32_t)(warp * 32) << 16) + a_col + ks * 8, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 128) {
        const uint32_t idesc = idesc_i8(1, 1, N);
        for (int ks = 0; ks < nk; ++ks) {
            const uint64_t bd = smem_desc(smem_u32(sB + ks * N * 32), kLBO, kSBO);
            mma_i8_ts(tmem_base, tmem_base + a_col + ks * 8, bd, idesc, ks > 0 ? 1u : 0u);
        }
        mma_commit(&bar);
    }
    if (warp < 4) {
        mbar_wait(&bar, 0);
        tc_fence_after();
        const int row = warp * 32 + (tid & 31);
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[8];
            tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int q = 0; q < 8; ++q) D[(size_t)row * N + c0 + q] = (int)v[q];
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

static int run_ts_tests() {
    int fails = 0;
    for (int N : {80, 160, 240}) {
        const int nk = 4, K = nk * 32;
        std::vector<uint8_t> A(128 * K), B(N * K);
        srand(4321 + N);
        for (auto& x : A) x = rand() & 255;
        for (auto& x : B) x = rand() & 255;
        uint8_t *dA, *dB; int* dD;
        cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, 128 * N * 4);
        cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xff, 128 * N * 4);
        const int smem = nk * N * 32;
        cudaFuncSetAttribute(umma_ts_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        umma_ts_test_kernel<<<1, 160, smem>>>(dA, dB, dD, N, nk);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("TS N=%d CUDA error: %s\n", N, cudaGetErrorString(e)); return 100; }
        std::vector<int> D(128 * N);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        long bad = 0; int first = -1;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            long ref = 0;
            for (int k = 0; k < K; ++k) ref += (long)(int)(int8_t)A[m * K + k] * (int)(int8_t)B[n * K + k];
            if (D[m * N + n] != (int)ref) { if (first < 0) first = m * N + n; ++bad; }
        }
        printf("TS (A in TMEM) N=%3d : %s (%ld mismatches)\n", N, bad ? "FAIL" : "ok", bad);
        if (bad) { int m = first / N, n = first % N; printf("   first mismatch at m=%d n=%d: got %d\n", m, n, D[m * N + n]); }
        fails += bad != 0;
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return fails;
}

int main() {
    int fails = 0;
    for (int N : {96, 128}) for (int as = 0; as < 2; ++as) for (int bs = 0; bs < 2; ++bs) {
        const int nk = 4, K = nk * 32;
        std::vector<uint8_t> A(128 * K), B(N * K);
        srand(1234 + N + as * 2 + bs);
        for (auto& x : A) x = rand() & 255;
        for (auto& x : B) x = rand() & 255;
        uint8_t *dA, *dB; int* dD;
        cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, 128 * 2 * N * 4);
        cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xff, 128 * 2 * N * 4);
        const int smem = nk * 4096 + nk * N * 32;
        cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        umma_test_kernel<<<1, 160, smem>>>(dA, dB, dD, N, nk, as, bs);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d as=%d bs=%d CUDA error: %s\n", N, as, bs, cudaGetErrorString(e)); return 2; }
        std::vector<int> D(128 * 2 * N);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        long bad = 0; int first = -1;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            long ref = 0;
            for (int k = 0; k < K; ++k) {
                const int a = as ? (int)(int8_t)A[m * K + k] : (int)A[m * K + k];
                const int b = bs ? (int)(int8_t)B[n * K + k] : (int)B[n * K + k];
                ref += (long)a * b;
            }
            if (D[m * 2 * N + n] != (int)ref || D[m * 2 * N + N + n] != (int)(2 * ref)) { if (first < 0) first = m * N + n; ++bad; }
        }
        printf("N=%3d a_signed=%d b_signed=%d : %s (%ld mismatches%s)\n", N, as, bs, bad ? "FAIL" : "ok", bad, bad ? "" : "");
        if (bad) { int m = first / N, n = first % N; printf("   first mismatch at m=%d n=%d: got %d / %d\n", m, n, D[m * 2 * N + n], D[m * 2 * N + N + n]); }
        fails += bad != 0;
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    fails += run_ts_tests();
    printf(fails ? "UMMA_TEST_FAIL\n" : "UMMA_TEST_OK\n");
    return fails ? 1 : 0;
}
