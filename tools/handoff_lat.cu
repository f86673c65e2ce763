// Latency microbenchmarks of the hand-offs used by the tcgen05 kernels (one CTA):
//  (a) tcgen05.commit -> mbarrier phase completion observed by the issuing thread (no MMAs pending / after one MMA)
//  (b) ping-pong between two warps through two mbarriers (arrive -> try_wait wake-up), per one-way hop
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/handoff_lat tools/handoff_lat.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include "../geobo_b200/csrc/umma.cuh"
using namespace umma;

__global__ void __launch_bounds__(160, 1) lat_kernel(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 4) {
        tmem_alloc(&tmem_base_s, 512);
        if (tid == 128) { for (int q = 0; q < 4; ++q) mbar_init(&bar[q], 1); fence_barrier_init(); }
    }
    for (int e = tid; e < 16384 / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x01010101u;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int N = 64;
    if (tid == 128) {
        // (a0) commit with nothing pending
        long long acc = 0;
        for (int i = 0; i < N; ++i) {
            long long t0 = clock64();
            mma_commit(&bar[0]);
            mbar_wait(&bar[0], i & 1);
            acc += clock64() - t0;
        }
        out[0] = acc / N;
        // (a1) one 128x240x32 MMA then commit
        acc = 0;
        const uint64_t ad = smem_desc(smem_u32(smem), kLBO, kSBO), bd = smem_desc(smem_u32(smem) + 4096, kLBO, kSBO);
        for (int i = 0; i < N; ++i) {
            long long t0 = clock64();
            mma_i8(tmem_base, ad, bd, idesc_i8(1, 1, 240), 0u);
            mma_commit(&bar[1]);
            mbar_wait(&bar[1], i & 1);
            acc += clock64() - t0;
        }
        out[1] = acc / N;
        // (a2) eight MMAs then commit
        acc = 0;
        for (int i = 0; i < N; ++i) {
            long long t0 = clock64();
            for (int q = 0; q < 8; ++q) mma_i8(tmem_base, ad, bd, idesc_i8(1, 1, 240), 1u);
            mma_commit(&bar[1]);
            mbar_wait(&bar[1], i & 1);
            acc += clock64() - t0;
        }
        out[2] = acc / N;
    }
    __syncthreads();
    // (b) ping-pong warp 0 lane 0 <-> warp 1 lane 0
    if (tid == 0) {
        long long t0 = clock64();
        for (int i = 0; i < N; ++i) {
            mbar_arrive(&bar[2]);
            mbar_wait(&bar[3], i & 1);
        }
        out[3] = (clock64() - t0) / (2 * N);
    } else if (tid == 32) {
        for (int i = 0; i < N; ++i) {
            mbar_wait(&bar[2], i & 1);
            mbar_arrive(&bar[3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 8 * sizeof(long long));
    cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    lat_kernel<<<1, 160, 16384>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[8]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("{\"commit_only_clk\": %lld, \"mma240_commit_clk\": %lld, \"mma240x8_commit_clk\": %lld, \"mbarrier_hop_clk\": %lld}\n", h[0], h[1], h[2], h[3]);
    return 0;
}
