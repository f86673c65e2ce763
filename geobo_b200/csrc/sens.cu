// Forward-model sensitivity matrices (geobo/sensormodel.py:29-133) on the device, fp64.
//
// One CTA per (sensor, chunk of voxel rows iy).  The CTA evaluates the prism-corner potential eZ on
// two consecutive y-planes of the edge lattice in shared memory (each corner is evaluated once per
// sensor instead of eight times), takes the 8-corner alternating difference with the reference's own
// association order, and writes the voxel row coalesced.  fp64 is mandatory: the 1e6 m edge padding
// (sensormodel.py:64-68) makes the corner differences catastrophically cancelling.
#include "common.cuh"
#include "formulas.cuh"

template <int KIND>
__global__ void __launch_bounds__(256) a_sens_kernel(const double* __restrict__ edges, const double* __restrict__ loc,
                                                     int xN, int yN, int zN, int rows_per_cta, double bx, double by,
                                                     double bz, double mul, double div, double* __restrict__ out, long ld,
                                                     int iy_begin, int iy_end) {
    // voxel rows [iy_begin, iy_end) only (a column chunk of the matrix: the voxel index is y-major); out holds those columns
    extern __shared__ double planes[];   // [2][(xN+1)*(zN+1)]
    const int n = blockIdx.x;
    const int iy0 = iy_begin + blockIdx.y * rows_per_cta;
    const int iy1 = min(iy_end, iy0 + rows_per_cta);
    if (iy0 >= iy_end) return;
    const int px = xN + 1, pz = zN + 1, plane = px * pz;
    const long nedge = (long)(yN + 1) * plane;
    const double* xE = edges;
    const double* yE = edges + nedge;
    const double* zE = edges + 2 * nedge;
    const double lx = loc[3 * n + 0], ly = loc[3 * n + 1], lz = loc[3 * n + 2];
    double* orow = out + (long)n * ld - (long)iy_begin * xN * zN;

    for (int j = iy0; j <= iy1; ++j) {
        double* cur = planes + (j & 1) * plane;
        for (int q = threadIdx.x; q < plane; q += blockDim.x) {
            const long e = (long)j * plane + q;
            double x0 = __dsub_rn(xE[e], lx), y0 = __dsub_rn(yE[e], ly);
            const double z0 = __dsub_rn(zE[e], lz);
            if (j == 0) { x0 = __dsub_rn(x0, GB_ALONG_WAY); y0 = __dsub_rn(y0, GB_ALONG_WAY); }    // :65-66
            if (j == yN) { x0 = __dadd_rn(x0, GB_ALONG_WAY); y0 = __dadd_rn(y0, GB_ALONG_WAY); }   // :67-68
            cur[q] = KIND == GB_SENS_GRAV ? grav_corner(x0, y0, z0) : magn_corner(x0, y0, z0, bx, by, bz);
        }
        __syncthreads();
        if (j > iy0) {
            const int iy = j - 1;
            const double* hi = cur;                               // eZ[i+1, ., .]
            const double* lo = planes + ((j - 1) & 1) * plane;    // eZ[i,   ., .]
            for (int v = threadIdx.x; v < xN * zN; v += blockDim.x) {
                const int ix = v / zN, iz = v - ix * zN;
                const int c00 = ix * pz + iz, c01 = c00 + 1, c10 = c00 + pz, c11 = c10 + 1;   // (x, z) corner offsets
                // sensormodel.py:83-84, same association order as the Python expression
                const double up = __dadd_rn(__dsub_rn(__dsub_rn(hi[c11], hi[c10]), hi[c01]), hi[c00]);
                const double dn = __dadd_rn(__dsub_rn(__dsub_rn(lo[c11], lo[c10]), lo[c01]), lo[c00]);
                const double s = -__dsub_rn(up, dn);
                orow[((long)iy * xN + ix) * zN + iz] = __ddiv_rn(__dmul_rn(mul, s), div);   // :88-91
            }
        }
        __syncthreads();
    }
}

cudaError_t launch_a_sens(int kind, const double B[3], const double* loc, int64_t nsens, const double* edges,
                          const int64_t n[3], double mul, double div, double* out, int64_t ld, int sm_count, cudaStream_t s) {
    return launch_a_sens_range(kind, B, loc, nsens, edges, n, mul, div, out, ld, 0, (int)n[1], sm_count, s);
}

// columns of voxel rows [iy_begin, iy_end) only: out[sensor * ld + (j - iy_begin * xN * zN)]
cudaError_t launch_a_sens_range(int kind, const double B[3], const double* loc, int64_t nsens, const double* edges,
                                const int64_t n[3], double mul, double div, double* out, int64_t ld, int iy_begin, int iy_end,
                                int sm_count, cudaStream_t s) {
    const int xN = (int)n[0], zN = (int)n[2];
    const int yN = iy_end - iy_begin;          // rows of this launch
    if (iy_begin < 0 || iy_end > (int)n[1] || yN < 1) return cudaErrorInvalidValue;
    const size_t smem = (size_t)2 * (xN + 1) * (zN + 1) * sizeof(double);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    // enough CTAs to fill the machine a few times over; each extra chunk re-evaluates one plane
    int64_t want = (8L * sm_count + nsens - 1) / nsens;
    if (want < 1) want = 1;
    int chunks = (int)(want < yN ? want : yN);
    int rows = (yN + chunks - 1) / chunks;
    chunks = (yN + rows - 1) / rows;
    dim3 grid((unsigned)nsens, (unsigned)chunks);
    cudaError_t e;
    if (kind == GB_SENS_GRAV) {
        e = cudaFuncSetAttribute(a_sens_kernel<GB_SENS_GRAV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        a_sens_kernel<GB_SENS_GRAV><<<grid, 256, smem, s>>>(edges, loc, xN, (int)n[1], zN, rows, B[0], B[1], B[2], mul, div, out, ld, iy_begin, iy_end);
    } else {
        e = cudaFuncSetAttribute(a_sens_kernel<GB_SENS_MAGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        a_sens_kernel<GB_SENS_MAGN><<<grid, 256, smem, s>>>(edges, loc, xN, (int)n[1], zN, rows, B[0], B[1], B[2], mul, div, out, ld, iy_begin, iy_end);
    }
    return cudaGetLastError();
}

// y = A x, one warp per row (simcube.py:149-150)
__global__ void gemv_kernel(const double* __restrict__ A, long rows, long cols, long ld, const double* __restrict__ x,
                            double* __restrict__ y) {
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (long c = lane; c < cols; c += 32) acc = fma(A[row * ld + c], x[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[row] = acc;
}

cudaError_t launch_gemv(const double* A, int64_t rows, int64_t cols, int64_t ld, const double* x, double* y, cudaStream_t s) {
    gemv_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(A, rows, cols, ld, x, y);
    return cudaGetLastError();
}

// sensormodel.grav_func / magn_func (sensormodel.py:96-133) applied elementwise to coordinate arrays.
__global__ void corner_func_kernel(int kind, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                   long count, double bx, double by, double bz, double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[i] = kind == GB_SENS_GRAV ? grav_corner(x[i], y[i], z[i]) : magn_corner(x[i], y[i], z[i], bx, by, bz);
}

cudaError_t launch_corner_func(int kind, const double* x, const double* y, const double* z, int64_t count, const double B[3],
                               double* out, cudaStream_t s) {
    corner_func_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(kind, x, y, z, count, B[0], B[1], B[2], out);
    return cudaGetLastError();
}

// sensormodel.A_drill (sensormodel.py:136-153): one-hot rows by exact coordinate equality.
__global__ void a_drill_kernel(const double* __restrict__ loc, long nd, const double* __restrict__ vp, long N, double* __restrict__ out) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long i = blockIdx.y;
    if (j >= N) return;
    const bool hit = vp[j] == loc[i] && vp[N + j] == loc[nd + i] && vp[2 * N + j] == loc[2 * nd + i];
    out[i * N + j] = hit ? 1.0 : 0.0;
}

cudaError_t launch_a_drill(const double* loc, int64_t nd, const double* vp, int64_t N, double* out, cudaStream_t s) {
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)nd);
    a_drill_kernel<<<grid, 256, 0, s>>>(loc, nd, vp, N, out);
    return cudaGetLastError();
}
