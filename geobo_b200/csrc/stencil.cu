// Compact-support (stencil) products with the 'sparse' covariance blocks (opt-in,
// gb_hyper.structure = GB_STRUCTURE_COMPACT): Pt = A3 . K and z = K . w as a 3-D tap sum over the offsets that can lie inside
// the kernel's support, taps read from the stationary covariance tables.  Per-thread arithmetic: stencil.cuh (also compiled
// for the host by the CPU tests).  No shared memory: the taps are warp-uniform loads (one broadcast per tap and block), the
// inputs are L1 / L2 hits of one row of A (N doubles) shared by the whole block.
#include "common.cuh"
#include "stencil.cuh"

// grid (tiles of STENCIL_THREADS work items of one x-z plane, local y-rows, rows of the chunk)
__global__ void __launch_bounds__(STENCIL_THREADS) stencil_kernel(StencilGeom g, const double* __restrict__ A, long lda,
                                                                  const double* __restrict__ tab0, double* __restrict__ out, long ldo,
                                                                  long r_stride_out, int accumulate, int nr) {
    stencil_item(g, A + (long)blockIdx.z * lda, tab0, g.jy0 + (int)blockIdx.y, (long)blockIdx.x * STENCIL_THREADS + threadIdx.x,
                 out + (long)blockIdx.z * ldo, r_stride_out, accumulate, nr);
}

// out[s][r * r_stride_out + (j - c0)] (+)= sum_i A[s][i] * K_(blk0 + r)[i][j]   for rows s < nrows, r = 0..2, j in [c0, c1);
// tab0 = tables + blk0 * ext + C0
cudaError_t stencil_apply(const StencilGeom& g, const double* tab0, const double* A, long lda, long nrows, double* out, long ldo,
                          long r_stride_out, int accumulate, cudaStream_t s, long* nlaunch, int nr) {
    const int nstrip = (g.zN + STENCIL_W - 1) / STENCIL_W;
    const unsigned tiles = (unsigned)(((long)g.xN * nstrip + STENCIL_THREADS - 1) / STENCIL_THREADS);
    const long chunk = 32768;
    for (long s0 = 0; s0 < nrows; s0 += chunk) {
        const long n = nrows - s0 < chunk ? nrows - s0 : chunk;
        stencil_kernel<<<dim3(tiles, (unsigned)g.nyl, (unsigned)n), STENCIL_THREADS, 0, s>>>(g, A + s0 * lda, lda, tab0, out + s0 * ldo, ldo,
                                                                                               r_stride_out, accumulate, nr);
        if (nlaunch) *nlaunch += 1;
    }
    return cudaGetLastError();
}
