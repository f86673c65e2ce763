// Bayesian-optimisation acquisition sweeps on the reconstructed cubes:
// the reference evaluates futility_vertical / futility_drill (geobo/run_geobo.py:175-235) one point at a time from
// scipy.optimize.shgo; the objective is piecewise constant in the voxel indices, so it is evaluated here for EVERY
// voxel column (vertical) or for a whole batch of (x0, y0, azimuth, dip) candidates (ray-marched gathers) in one launch.
#include "common.cuh"

// out[a * n1 + b] = sum_z rec[a, b, :] + kappa * sqrt(sum_z var[a, b, :]) - beta * sum_z costs[a, b, :]
// for interior columns 0 < a < n0-1, 0 < b < n1-1 (run_geobo.py:192-194), -inf elsewhere (:195-198).  One warp per column.
__global__ void __launch_bounds__(256) acq_vertical_kernel(const double* __restrict__ rec, const double* __restrict__ var,
                                                           const double* __restrict__ costs, long n0, long n1, long n2,
                                                           double kappa, double beta, double* __restrict__ out) {
    const long col = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (col >= n0 * n1) return;
    const long a = col / n1, b = col % n1;
    double sr = 0.0, sv = 0.0, sc = 0.0;
    for (long z = lane; z < n2; z += 32) {
        sr += rec[col * n2 + z];
        sv += var[col * n2 + z];
        if (costs) sc += costs[col * n2 + z];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
        sc += __shfl_xor_sync(0xffffffffu, sc, o);
    }
    if (lane == 0) {
        const bool interior = a > 0 && a < n0 - 1 && b > 0 && b < n1 - 1;
        out[col] = interior ? sr + kappa * sqrt(sv) - beta * sc : -INFINITY;
    }
}

// futility_drill (run_geobo.py:203-235) for a batch of candidates; one thread per candidate.
// out[i] = +utility (the reference returns its negative), 0 when the ray leaves the cube (the reference's IndexError ->
// `except: funct = 0.`); negative voxel indices wrap like NumPy fancy indexing does (index -k = n - k).
__global__ void acq_drill_kernel(const double* __restrict__ rec, const double* __restrict__ var, const double* __restrict__ costs,
                                 long n0, long n1, long n2, double vx, double vy, double vz, double zmax, double length, int nstep,
                                 const double* __restrict__ params, long n, double kappa, double beta, double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x0 = params[4 * i], y0 = params[4 * i + 1], az = params[4 * i + 2], dip = params[4 * i + 3];
    if (!(isfinite(x0) && isfinite(y0) && isfinite(az) && isfinite(dip))) { out[i] = 0.0; return; }
    const double phi = az * M_PI / 180.0, theta = (180.0 - dip) * M_PI / 180.0;      // :225
    const double st = sin(theta), ct = cos(theta), cp = cos(phi), sp = sin(phi);
    const double step = nstep > 1 ? length / (double)(nstep - 1) : 0.0;               // np.linspace(0, length, nstep), :218
    double sr = 0.0, sv = 0.0, sc = 0.0;
    bool ok = true;
    for (int k = 0; k < nstep; ++k) {
        const double r = (k == nstep - 1 && nstep > 1) ? length : k * step;
        const double x = x0 + r * st * cp, y = y0 + r * st * sp, z = zmax + r * ct;    // utils.spherical2cartes
        const double qx = x / vx, qy = y / vy, qz = -z / vz;                          // :226-228, astype(int) truncates toward zero
        if (!(fabs(qx) < 2e9 && fabs(qy) < 2e9 && fabs(qz) < 2e9)) { ok = false; break; }
        long ix = (long)qx, iy = (long)qy, iz = (long)qz;
        if (ix < 0) ix += n0;
        if (iy < 0) iy += n1;
        if (iz < 0) iz += n2;
        if (ix < 0 || ix >= n0 || iy < 0 || iy >= n1 || iz < 0 || iz >= n2) { ok = false; break; }
        const long idx = (ix * n1 + iy) * n2 + iz;
        sr += rec[idx];
        sv += var[idx];
        if (costs) sc += costs[idx];
    }
    out[i] = ok ? sr + kappa * sqrt(sv) - beta * sc : 0.0;                             // :229-231
}

extern "C" int gb_acquisition_vertical(gb_ctx* ctx, const double* rec, const double* var, const double* costs, const int64_t shape[3],
                                       double kappa, double beta, double* out) {
    if (!ctx || !rec || !var || !shape || !out || shape[0] < 1 || shape[1] < 1 || shape[2] < 1)
        return gb_fail(ctx, GB_ERR_ARG, "gb_acquisition_vertical: bad argument");
    const long n0 = shape[0], n1 = shape[1], n2 = shape[2], nvox = n0 * n1 * n2;
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> r, v, c, o;
    GB_CUDA(ctx, r.alloc(nvox));
    GB_CUDA(ctx, v.alloc(nvox));
    GB_CUDA(ctx, o.alloc(n0 * n1));
    GB_CUDA(ctx, cudaMemcpyAsync(r.p, rec, nvox * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(v.p, var, nvox * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (costs) {
        GB_CUDA(ctx, c.alloc(nvox));
        GB_CUDA(ctx, cudaMemcpyAsync(c.p, costs, nvox * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    acq_vertical_kernel<<<(unsigned)((n0 * n1 + 7) / 8), 256, 0, ctx->stream>>>(r.p, v.p, costs ? c.p : nullptr, n0, n1, n2, kappa, beta, o.p);
    GB_CUDA(ctx, cudaGetLastError());
    GB_CUDA(ctx, cudaMemcpyAsync(out, o.p, n0 * n1 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

extern "C" int gb_acquisition_drill(gb_ctx* ctx, const double* rec, const double* var, const double* costs, const int64_t shape[3],
                                    const double voxsize[3], double zmax, double length, const double* params, int64_t n,
                                    double kappa, double beta, double* out) {
    if (!ctx || !rec || !var || !shape || !voxsize || !params || !out || n < 1 || shape[0] < 1 || shape[1] < 1 || shape[2] < 1)
        return gb_fail(ctx, GB_ERR_ARG, "gb_acquisition_drill: bad argument");
    const long n0 = shape[0], n1 = shape[1], n2 = shape[2], nvox = n0 * n1 * n2;
    double vmin = voxsize[0] < voxsize[1] ? voxsize[0] : voxsize[1];
    if (voxsize[2] < vmin) vmin = voxsize[2];
    if (!(vmin > 0.0) || !(length > 0.0)) return gb_fail(ctx, GB_ERR_ARG, "gb_acquisition_drill: voxel sizes and length must be positive");
    const int nstep = (int)(2.0 * length / vmin);                                     // run_geobo.py:217
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> r, v, c, pr, o;
    GB_CUDA(ctx, r.alloc(nvox));
    GB_CUDA(ctx, v.alloc(nvox));
    GB_CUDA(ctx, pr.alloc(4 * n));
    GB_CUDA(ctx, o.alloc(n));
    GB_CUDA(ctx, cudaMemcpyAsync(r.p, rec, nvox * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(v.p, var, nvox * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(pr.p, params, 4 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (costs) {
        GB_CUDA(ctx, c.alloc(nvox));
        GB_CUDA(ctx, cudaMemcpyAsync(c.p, costs, nvox * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    acq_drill_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(r.p, v.p, costs ? c.p : nullptr, n0, n1, n2, voxsize[0], voxsize[1],
                                                                           voxsize[2], zmax, length, nstep, pr.p, n, kappa, beta, o.p);
    GB_CUDA(ctx, cudaGetLastError());
    GB_CUDA(ctx, cudaMemcpyAsync(out, o.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}
