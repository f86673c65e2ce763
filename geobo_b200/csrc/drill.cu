// Drill-core samples -> voxel cube: the reference's align_drill (geobo/run_geobo.py:132-159)
// is a Python triple loop over all voxels with an O(samples) mask each.  Here one thread owns one voxel and the samples
// stream through shared memory in tiles (every thread of the block reads the same sample: one broadcast per load), in
// sample order, so the per-voxel sum is the plain left-to-right sum of the selected samples.
#include "common.cuh"
#include "drill.cuh"

constexpr int DRILL_TILE = 1024;
constexpr int DRILL_THREADS = 256;

__global__ void __launch_bounds__(DRILL_THREADS) align_drill_kernel(const double* __restrict__ voxelpos, long N,
                                                                    const double* __restrict__ coord, const double* __restrict__ data,
                                                                    long ns, double dx, double dy, double dz, double* __restrict__ out) {
    __shared__ double sx[DRILL_TILE], sy[DRILL_TILE], sz[DRILL_TILE], sv[DRILL_TILE];
    const long v = (long)blockIdx.x * DRILL_THREADS + threadIdx.x;
    const bool live = v < N;
    double vx = 0.0, vy = 0.0, vz = 0.0;
    if (live) { vx = voxelpos[v]; vy = voxelpos[N + v]; vz = voxelpos[2 * N + v]; }
    DrillAcc acc;
    acc.sum = 0.0;
    acc.cnt = 0;
    for (long t0 = 0; t0 < ns; t0 += DRILL_TILE) {
        const long n = ns - t0 < DRILL_TILE ? ns - t0 : DRILL_TILE;
        __syncthreads();
        for (long i = threadIdx.x; i < n; i += DRILL_THREADS) {
            sx[i] = coord[3 * (t0 + i)];
            sy[i] = coord[3 * (t0 + i) + 1];
            sz[i] = coord[3 * (t0 + i) + 2];
            sv[i] = data[t0 + i];
        }
        __syncthreads();
        if (live) drill_accumulate(acc, vx, vy, vz, sx, sy, sz, sv, n, dx, dy, dz);
    }
    if (live) out[v] = drill_finish(acc);
}

extern "C" int gb_align_drill(gb_ctx* ctx, const double* voxelpos, int64_t n_vox, const double* coord, const double* data, int64_t ns,
                              const double voxsize[3], double* out) {
    if (!ctx || !voxelpos || !voxsize || !out || n_vox < 1 || ns < 0 || (ns > 0 && (!coord || !data)))
        return gb_fail(ctx, GB_ERR_ARG, "gb_align_drill: bad argument");
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> vp, cd, dv, o;
    GB_CUDA(ctx, vp.alloc((size_t)3 * n_vox));
    GB_CUDA(ctx, cd.alloc((size_t)3 * ns));
    GB_CUDA(ctx, dv.alloc((size_t)ns));
    GB_CUDA(ctx, o.alloc((size_t)n_vox));
    GB_CUDA(ctx, cudaMemcpyAsync(vp.p, voxelpos, (size_t)3 * n_vox * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (ns > 0) {
        GB_CUDA(ctx, cudaMemcpyAsync(cd.p, coord, (size_t)3 * ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        GB_CUDA(ctx, cudaMemcpyAsync(dv.p, data, (size_t)ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    const unsigned blocks = (unsigned)((n_vox + DRILL_THREADS - 1) / DRILL_THREADS);
    align_drill_kernel<<<blocks, DRILL_THREADS, 0, ctx->stream>>>(vp.p, (long)n_vox, cd.p, dv.p, (long)ns, voxsize[0], voxsize[1], voxsize[2], o.p);
    GB_CUDA(ctx, cudaGetLastError());
    GB_CUDA(ctx, cudaMemcpyAsync(out, o.p, (size_t)n_vox * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}
