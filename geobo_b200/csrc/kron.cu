// Kronecker-structured products with the squared-exponential covariance blocks (opt-in,
// gb_hyper.structure = GB_STRUCTURE_KRON): Pt = A3 . K and z = K . w as three Toeplitz mode products per block instead of
// the dense N x N contraction.  The per-thread arithmetic lives in kron.cuh (also compiled for the host by the CPU tests).
//
//   kron_factors_kernel : 9 blocks x 3 axes factor lines from the stationary covariance tables
//   kron_y_kernel       : y mode for the three property blocks of one data block, rows [s0, s0 + chunk) -> scratch T
//   kron_zx_kernel      : z and x modes of one x-z plane in shared memory -> rows of Pt (or of z, accumulating over c)
// Both compute kernels are fp64 FMA kernels whose operands come from L2 / shared memory; the scratch T of a row chunk is sized
// to stay L2 resident (GEOBO_B200_KRON_SCRATCH_MB, default 64), so HBM sees A once and Pt once.
#include "common.cuh"
#include "kron.cuh"

__global__ void kron_factors_kernel(const double* __restrict__ tables, long ext, long C0, KronGeom g, double* __restrict__ kf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int axis = blockIdx.y, b = blockIdx.z;
    if (i < g.FL) kf[((long)b * 3 + axis) * g.FL + i] = kron_factor(tables + (long)b * ext + C0, g, axis, i);
}

// grid (plane tiles, groups of KRON_JT output y-rows, rows of the chunk)
__global__ void __launch_bounds__(KRON_YTHREADS) kron_y_kernel(KronGeom g, const double* __restrict__ A, long lda,
                                                               const double* __restrict__ kf, int blk0, double* __restrict__ T,
                                                               long r_stride) {
    extern __shared__ double kron_smem[];
    for (int i = threadIdx.x; i < 3 * g.FL; i += KRON_YTHREADS) kron_smem[i] = kf[((long)(blk0 + i / g.FL) * 3 + 0) * g.FL + i % g.FL];
    __syncthreads();
    kron_y_thread(g, A + (long)blockIdx.z * lda, kron_smem, (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x,
                  T + (long)blockIdx.z * g.nyl * g.XZ, r_stride);
}

// grid (local y-rows, rows of the chunk, 3 property blocks)
__global__ void __launch_bounds__(KRON_ZXTHREADS) kron_zx_kernel(KronGeom g, const double* __restrict__ T, long r_stride_T,
                                                                 const double* __restrict__ kf, int blk0, double* __restrict__ out,
                                                                 long ldo, long r_stride_out, int accumulate) {
    extern __shared__ double kron_smem[];
    double* pin = kron_smem;
    double* ptmp = pin + (long)g.xN * g.zs;
    double* fx = ptmp + (long)g.xN * g.zs;
    double* fz = fx + g.FL;
    const int jl = blockIdx.x, sl = blockIdx.y, r = blockIdx.z;
    const double* kfb = kf + (long)(blk0 + r) * 3 * g.FL;
    kron_zx_load(g, T + r * r_stride_T + ((long)sl * g.nyl + jl) * g.XZ, kfb + g.FL, kfb + 2 * g.FL, (int)threadIdx.x, KRON_ZXTHREADS,
                 pin, fx, fz);
    __syncthreads();
    kron_z_phase(g, pin, fz, (int)threadIdx.x, KRON_ZXTHREADS, ptmp);
    __syncthreads();
    kron_x_phase(g, ptmp, fx, g.jy0 + jl, (int)threadIdx.x, KRON_ZXTHREADS, out + (long)sl * ldo + r * r_stride_out, accumulate);
}

static size_t kron_zx_smem(const KronGeom& g) { return ((size_t)2 * g.xN * g.zs + 2 * (size_t)g.FL) * sizeof(double); }

long kron_factor_doubles(const KronGeom& g) { return 9L * 3 * g.FL; }

long kron_scratch_doubles(const KronGeom& g, long rows) {
    long mb = 64;
    if (const char* e = getenv("GEOBO_B200_KRON_SCRATCH_MB")) mb = atol(e) > 0 ? atol(e) : mb;
    const long per_row = 3L * g.nyl * g.XZ;
    long chunk = (mb << 20) / (long)sizeof(double) / per_row;
    if (chunk < 1) chunk = 1;
    if (chunk > rows) chunk = rows;
    return chunk * per_row;
}

int kron_supported(const KronGeom& g, char* why, size_t len) {
    if (kron_zx_smem(g) > 227 * 1024) {
        snprintf(why, len, "structure = kron: one x-z plane (%d x %d) does not fit the %d KB of shared memory of one block", g.xN, g.zN, 227);
        return 0;
    }
    return 1;
}

cudaError_t kron_build_factors(const double* tables, long ext, long C0, const KronGeom& g, double* kf, cudaStream_t s) {
    dim3 grid((unsigned)((g.FL + 127) / 128), 3, 9);
    kron_factors_kernel<<<grid, 128, 0, s>>>(tables, ext, C0, g, kf);
    return cudaGetLastError();
}

// out[s][r * r_stride_out + (j - c0)] (+)= sum_i A[s][i] * K_(blk0 + r)[i][j]   for rows s < nrows, r = 0..2, j in [c0, c1)
cudaError_t kron_apply(const KronGeom& g, const double* kf, int blk0, const double* A, long lda, long nrows, double* T, long T_doubles,
                       double* out, long ldo, long r_stride_out, int accumulate, cudaStream_t s, long* nlaunch, int nr) {
    const long per_row = 3L * g.nyl * g.XZ;
    long chunk = T_doubles / per_row;
    if (chunk < 1) return cudaErrorInvalidValue;
    if (chunk > 32768) chunk = 32768;
    const size_t smem_zx = kron_zx_smem(g), smem_y = (size_t)3 * g.FL * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(kron_zx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_zx);
    if (e != cudaSuccess) return e;
    const unsigned qtiles = (unsigned)((g.XZ + KRON_YTHREADS * KRON_YQ - 1) / (KRON_YTHREADS * KRON_YQ));
    const unsigned jgroups = (unsigned)((g.nyl + KRON_JT - 1) / KRON_JT);
    for (long s0 = 0; s0 < nrows; s0 += chunk) {
        const long n = nrows - s0 < chunk ? nrows - s0 : chunk;
        const long r_stride = n * g.nyl * g.XZ;
        kron_y_kernel<<<dim3(qtiles, jgroups, (unsigned)n), KRON_YTHREADS, smem_y, s>>>(g, A + s0 * lda, lda, kf, blk0, T, r_stride);
        kron_zx_kernel<<<dim3((unsigned)g.nyl, (unsigned)n, (unsigned)nr), KRON_ZXTHREADS, smem_zx, s>>>(g, T, r_stride, kf, blk0, out + s0 * ldo, ldo,
                                                                                                 r_stride_out, accumulate);
        if (nlaunch) *nlaunch += 2;
    }
    return cudaGetLastError();
}
