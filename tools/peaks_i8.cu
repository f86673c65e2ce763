// int8 tcgen05 (UTCIMMA) issue-rate microbenchmark: one CTA per SM, one thread issues back-to-back
// tcgen05.mma.kind::i8 128 x N x 32 from fixed shared-memory tiles (K-major, no swizzle) into TMEM.
// Prints clocks per instruction and the resulting dense int8 TOP/s for several N and for the digit-product
// instruction mixes of ozaki.cu (S = 4, 5, 6).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/peaks_i8 tools/peaks_i8.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../geobo_b200/csrc/umma.cuh"
using namespace umma;

// mode 0: all instructions N = n0, alternating between two A tiles;  mode S (4,5,6): the ozaki instruction mix
__global__ void __launch_bounds__(160, 1) i8_rate_kernel(int n0, int mode, int iters, long long* clocks) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (6 * 4096 + 512 * 32) / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x01020304u * (e % 61 + 1);
    fence_proxy_async_smem();
    if (warp == 4) {
        tmem_alloc(&tmem_base_s, 512);
        if (tid == 128) { mbar_init(&bar, 1); fence_barrier_init(); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 128) {
        const uint32_t sa = smem_u32(smem), sb = sa + 6 * 4096;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (mode == 0) {
                mma_i8(tmem_base, smem_desc(sa + (i & 3) * 4096, kLBO, kSBO), smem_desc(sb, kLBO, kSBO), idesc_i8(1, 1, n0), 1u);
            } else if (mode < 0) {     // rotate over -mode independent accumulators
                const int nrot = -mode;
                mma_i8(tmem_base + (i % nrot) * n0, smem_desc(sa + (i & 3) * 4096, kLBO, kSBO), smem_desc(sb, kLBO, kSBO), idesc_i8(1, 1, n0), 1u);
            } else if (mode >= 14 && mode <= 16) {   // ozaki mix, reordered so that consecutive instructions touch disjoint levels where possible
                const int S = mode - 10, NT = S == 4 ? 128 : S == 5 ? 96 : 80, G = 256 / NT;
                for (int half = 0; half < 2; ++half)
                    for (int qa = 0; qa < S; ++qa) {
                        int cnt = 0;
                        for (int qb0 = 0; qb0 < S - qa; qb0 += G, ++cnt) {
                            if ((cnt & 1) != half) continue;
                            const int g = (S - qa - qb0) < G ? (S - qa - qb0) : G;
                            mma_i8(tmem_base + (qa + qb0) * NT, smem_desc(sa + qa * 4096, kLBO, kSBO), smem_desc(sb + qb0 * NT * 32, kLBO, kSBO),
                                   idesc_i8(1, 1, g * NT), 1u);
                        }
                    }
            } else {
                const int S = mode, NT = S == 4 ? 128 : S == 5 ? 96 : 80, G = 256 / NT;
                for (int qa = 0; qa < S; ++qa)
                    for (int qb0 = 0; qb0 < S - qa; qb0 += G) {
                        const int g = (S - qa - qb0) < G ? (S - qa - qb0) : G;
                        mma_i8(tmem_base + (qa + qb0) * NT, smem_desc(sa + qa * 4096, kLBO, kSBO), smem_desc(sb + qb0 * NT * 32, kLBO, kSBO),
                               idesc_i8(1, 1, g * NT), 1u);
                    }
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        clocks[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// Fully unrolled variant: descriptors are loop invariant, NROT independent accumulators, 8 instructions per loop trip.
template <int N, int NROT>
__global__ void __launch_bounds__(160, 1) i8_rate_unrolled(int iters, long long* clocks) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (6 * 4096 + 512 * 32) / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x01020304u * (e % 61 + 1);
    fence_proxy_async_smem();
    if (warp == 4) {
        tmem_alloc(&tmem_base_s, 512);
        if (tid == 128) { mbar_init(&bar, 1); fence_barrier_init(); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 128) {
        const uint32_t sa = smem_u32(smem), sb = sa + 6 * 4096;
        const uint64_t ad0 = smem_desc(sa, kLBO, kSBO), ad1 = smem_desc(sa + 4096, kLBO, kSBO), bd = smem_desc(sb, kLBO, kSBO);
        constexpr uint32_t idesc = idesc_i8(1, 1, N);
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) mma_i8(tmem_base + (u % NROT) * N, (u & 1) ? ad1 : ad0, bd, idesc, 1u);
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        clocks[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// A operand from tensor memory (TS form): the digit-product mix of the v3 projection kernel, S slices, tile 128 x NT,
// one A digit times up to 256 / NT B digit planes per instruction; accumulators at columns 0.., A digits at column 480..
template <int S, int NT, int COMMITS>
__global__ void __launch_bounds__(160, 1) i8_rate_ts_mix(int iters, long long* clocks) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, dummy[8];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (6 * 4096 + 512 * 32) / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x01020304u * (e % 61 + 1);
    fence_proxy_async_smem();
    if (warp == 4) {
        tmem_alloc(&tmem_base_s, 512);
        if (tid == 128) { mbar_init(&bar, 1); for (int q = 0; q < 8; ++q) mbar_init(&dummy[q], 1); fence_barrier_init(); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    constexpr int G = (256 / NT) < S ? (256 / NT) : S;
    if (tid == 128) {
        const uint32_t sb = smem_u32(smem) + 6 * 4096;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int qa = 0; qa < S; ++qa)
#pragma unroll
                for (int qb0 = 0; qb0 < S - qa; qb0 += G) {
                    const int g = (S - qa - qb0) < G ? (S - qa - qb0) : G;
                    mma_i8_ts(tmem_base + (qa + qb0) * NT, tmem_base + 512 - 8 * S + qa * 8, smem_desc(sb + qb0 * NT * 32, kLBO, kSBO),
                              idesc_i8(1, 1, g * NT), 1u);
                }
#pragma unroll
            for (int q = 0; q < COMMITS; ++q) mma_commit(&dummy[(i + q) & 7]);     // nobody waits on these: cost of the commit itself
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        clocks[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

template <int S, int NT, int COMMITS>
static void run_ts_mix(int sms, long long* dclk) {
    const int smem = 6 * 4096 + 512 * 32, iters = 20000;
    cudaFuncSetAttribute(i8_rate_ts_mix<S, NT, COMMITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    i8_rate_ts_mix<S, NT, COMMITS><<<sms, 160, smem>>>(100, dclk);
    cudaEventRecord(e0);
    i8_rate_ts_mix<S, NT, COMMITS><<<sms, 160, smem>>>(iters, dclk);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); exit(1); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> clk(sms);
    cudaMemcpy(clk.data(), dclk, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double cavg = 0; for (auto c : clk) cavg += (double)c; cavg /= sms;
    const double macs = 128.0 * NT * 32 * (S * (S + 1) / 2);
    printf("  {\"ts_mix\": true, \"commits_per_step\": %d, \"S\": %d, \"NT\": %d, \"clk_per_kstep\": %.1f, \"clk_per_column\": %.2f, \"int8_tops\": %.1f, \"macs_per_clk_per_sm\": %.0f},\n",
           COMMITS, S, NT, cavg / iters, cavg / iters / NT, 2.0 * macs * iters * sms / (ms * 1e-3) / 1e12, macs / (cavg / iters));
}

template <int N, int NROT>
static void run_unrolled(int sms, long long* dclk, bool last) {
    const int smem = 6 * 4096 + 512 * 32, iters = 200000;
    cudaFuncSetAttribute(i8_rate_unrolled<N, NROT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    i8_rate_unrolled<N, NROT><<<sms, 160, smem>>>(1024, dclk);
    cudaEventRecord(e0);
    i8_rate_unrolled<N, NROT><<<sms, 160, smem>>>(iters, dclk);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); exit(1); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> clk(sms);
    cudaMemcpy(clk.data(), dclk, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double cavg = 0; for (auto c : clk) cavg += (double)c; cavg /= sms;
    const double macs = 128.0 * N * 32;
    printf("  {\"unrolled\": true, \"n\": %d, \"accumulators\": %d, \"clk_per_mma\": %.1f, \"ms\": %.3f, \"int8_tops\": %.1f, \"macs_per_clk_per_sm\": %.0f}%s\n",
           N, NROT, cavg / iters, ms, 2.0 * macs * iters * sms / (ms * 1e-3) / 1e12, macs / (cavg / iters), last ? "" : ",");
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    long long* dclk;
    cudaMalloc(&dclk, sms * sizeof(long long));
    const int smem = 6 * 4096 + 512 * 32;
    cudaFuncSetAttribute(i8_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("{\"sms\": %d, \"clock_khz\": %d, \"runs\": [\n", sms, khz);
    run_ts_mix<5, 80, 0>(sms, dclk);
    run_ts_mix<5, 80, 1>(sms, dclk);
    run_ts_mix<5, 80, 2>(sms, dclk);
    run_ts_mix<5, 48, 0>(sms, dclk);
    run_ts_mix<5, 48, 1>(sms, dclk);
    run_ts_mix<5, 48, 2>(sms, dclk);
    run_ts_mix<6, 64, 0>(sms, dclk);
    run_ts_mix<4, 112, 0>(sms, dclk);
    run_unrolled<256, 1>(sms, dclk, false);
    run_unrolled<256, 2>(sms, dclk, false);
    run_unrolled<240, 2>(sms, dclk, false);
    run_unrolled<192, 2>(sms, dclk, false);
    run_unrolled<160, 3>(sms, dclk, false);
    run_unrolled<128, 1>(sms, dclk, false);
    run_unrolled<128, 4>(sms, dclk, false);
    run_unrolled<96, 5>(sms, dclk, false);
    run_unrolled<80, 1>(sms, dclk, false);
    run_unrolled<80, 6>(sms, dclk, false);
    run_unrolled<64, 8>(sms, dclk, false);
    run_unrolled<32, 8>(sms, dclk, false);
    const int cases[][2] = {{256, 0}, {0, 4}, {0, 5}, {0, 6}};
    const int ncase = sizeof(cases) / sizeof(cases[0]);
    for (int ci = 0; ci < ncase; ++ci) {
        const int n0 = cases[ci][0], mode = cases[ci][1];
        const int iters = mode ? 20000 : 100000;
        i8_rate_kernel<<<sms, 160, smem>>>(n0, mode, 1000, dclk);   // warm-up
        cudaEventRecord(e0);
        i8_rate_kernel<<<sms, 160, smem>>>(n0, mode, iters, dclk);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); return 1; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> clk(sms);
        cudaMemcpy(clk.data(), dclk, sms * sizeof(long long), cudaMemcpyDeviceToHost);
        double cavg = 0; for (auto c : clk) cavg += (double)c; cavg /= sms;
        double macs_per_iter;
        if (mode <= 0) macs_per_iter = 128.0 * n0 * 32;
        else { const int S = mode > 10 ? mode - 10 : mode, NT = S == 4 ? 128 : S == 5 ? 96 : 80; macs_per_iter = 128.0 * NT * 32 * (S * (S + 1) / 2); }
        const double tops = 2.0 * macs_per_iter * iters * sms / (ms * 1e-3) / 1e12;
        printf("  {\"n\": %d, \"mix_S\": %d, \"clk_per_iter\": %.1f, \"ms\": %.3f, \"int8_tops\": %.1f, \"macs_per_clk_per_sm\": %.0f}%s\n",
               n0, mode, cavg / iters, ms, tops, macs_per_iter / (cavg / iters), ci + 1 < ncase ? "," : "");
    }
    printf("]}\n");
    return 0;
}
