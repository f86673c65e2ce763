// Measures the fp64 denominators that MEASURED_PEAKS.json does not carry:
//   * DMMA.8x8x4 issue-bound peak (register-resident mma.sync.m8n8k4.f64 loop)
//   * DFMA peak (vector fp64 pipe)
//   * cuBLAS DGEMM 8192^3 (library reference for the fp64 roofline; test/measurement only, never on the product path)
//   * device-to-device copy bandwidth
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/peaks tools/peaks.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void dmma_peak(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_peak(double* out, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1e-9;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, p.multiProcessorCount);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double* out;
    CK(cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double)));
    float ms;
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int threads = warps * 32, blocks = p.multiProcessorCount * 2, iters = 20000;
        dmma_peak<<<blocks, threads>>>(out, 100);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        dmma_peak<<<blocks, threads>>>(out, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = (double)blocks * warps * iters * 16 * 512.0;
        printf(", \"dmma_tflops_w%d\": %.2f", warps * 2, flops / ms / 1e9);
    }
    {
        const int threads = 512, blocks = p.multiProcessorCount * 4, iters = 20000;
        dfma_peak<<<blocks, threads>>>(out, 100);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        dfma_peak<<<blocks, threads>>>(out, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"dfma_tflops\": %.2f", (double)blocks * threads * iters * 16 * 2.0 / ms / 1e9);
    }
    {
        const int n = 8192;
        double *A, *B, *C;
        CK(cudaMalloc(&A, (size_t)n * n * 8));
        CK(cudaMalloc(&B, (size_t)n * n * 8));
        CK(cudaMalloc(&C, (size_t)n * n * 8));
        CK(cudaMemset(A, 0, (size_t)n * n * 8));
        CK(cudaMemset(B, 0, (size_t)n * n * 8));
        cublasHandle_t h;
        cublasCreate(&h);
        const double one = 1.0, zero = 0.0;
        cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0));
            cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf(", \"cublas_dgemm_8192_tflops\": %.2f", 2.0 * n * n * (double)n / best / 1e9);
        // sustained: 3 s back to back
        int reps = (int)(3000.0f / best) + 1;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"cublas_dgemm_8192_tflops_sustained\": %.2f", 2.0 * n * n * (double)n * reps / ms / 1e9);
        // copy bandwidth
        size_t bytes = (size_t)n * n * 8;
        CK(cudaMemcpy(C, A, bytes, cudaMemcpyDeviceToDevice));
        best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0));
            CK(cudaMemcpyAsync(C, A, bytes, cudaMemcpyDeviceToDevice));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf(", \"d2d_copy_gbs\": %.1f", 2.0 * bytes / best / 1e6);
    }
    printf("}\n");
    return 0;
}
