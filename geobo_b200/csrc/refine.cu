// fp64 matrix-free application of the data-space operator  (A3 K A3^T + Sigma)  and of  K A3^T :
// the pieces of the iterative refinement of  alpha = (A K A^T + Sigma)^-1 y  and of the posterior mean
//     mu = K A3^T alpha                      (inversion.py:96-115: mu = V^T u = K A3^T (A K A^T + Sigma)^-1 y)
// used by the int8 tensor-core path: the factor L comes from an AkA whose operands were rounded to S digits, so it is
// used as a preconditioner and the residual  y - (A3 K A3^T + Sigma) alpha  is evaluated here with the fp64 sensitivities
// and the fp64 stationary-covariance tables -- no Pt, no digits.  Cost per application: 9 N^2 table FMAs (the
// covariance block matrix times one vector) + two passes over A.
#include "common.cuh"

// partial[part][c][j] = sum_{s in part} A_c[s][j] * alpha[c*Ns + s]      (A3^T alpha for the two survey blocks)
__global__ void __launch_bounds__(256) at_gemv_partial_kernel(const double* __restrict__ A0, const double* __restrict__ A1, long Ns,
                                                              long N, long lda, const double* __restrict__ alpha, int nsplit,
                                                              double* __restrict__ partial, long Kp) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int part = blockIdx.y, c = blockIdx.z;
    if (j >= N) return;
    const double* A = c ? A1 : A0;
    const long per = (Ns + nsplit - 1) / nsplit;
    const long s0 = part * per, s1 = min(Ns, s0 + per);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    long s = s0;
    for (; s + 3 < s1; s += 4) {
        a0 = fma(A[s * lda + j], alpha[c * Ns + s], a0);
        a1 = fma(A[(s + 1) * lda + j], alpha[c * Ns + s + 1], a1);
        a2 = fma(A[(s + 2) * lda + j], alpha[c * Ns + s + 2], a2);
        a3 = fma(A[(s + 3) * lda + j], alpha[c * Ns + s + 3], a3);
    }
    for (; s < s1; ++s) a0 = fma(A[s * lda + j], alpha[c * Ns + s], a0);
    partial[((long)part * 2 + c) * Kp + j] = (a0 + a1) + (a2 + a3);
}

// w[c][j] = sum_part partial (c = 0, 1; fixed order), w[2][j] = 0 (the drill block is scattered afterwards)
__global__ void w_reduce_kernel(const double* __restrict__ partial, int nsplit, long Kp, long N, double* __restrict__ w) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Kp) return;
    for (int c = 0; c < 2; ++c) {
        double s = 0.0;
        if (j < N)
            for (int p = 0; p < nsplit; ++p) s += partial[((long)p * 2 + c) * Kp + j];
        w[c * Kp + j] = s;
    }
    w[2 * Kp + j] = 0.0;
}

// A_drill^T alpha_drill: one-hot rows (sensormodel.py:136-153), drilled voxels are distinct
__global__ void w_drill_kernel(const double* __restrict__ alpha_drill, const int64_t* __restrict__ drill, long nd, double* __restrict__ w2) {
    const long d = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < nd) w2[drill[d]] = alpha_drill[d];
}

// z[r][i] = sum_c sum_j K[(r,i),(c,j)] w[c][j]  for this rank's voxels (16-voxel z segments; needs zN % 16 == 0 and
// c0 % 16 == 0).  K comes from the stationary tables: K[(r,i),(c,j)] = tab[c*3+r][C0 + L(j) - L(i)].
// Block = 16 consecutive output segments x 4 z-quads x 4 groups of the contraction-column sweep.  A thread owns 4
// consecutive outputs and slides a 4-entry window of the table along the contraction column (1 table load + 1/3 w load
// per 4 FMAs).  The 16 segments of a block visit the contraction columns in a ROTATED order (cj = sweep index + own
// column) so that at any moment they all need the SAME lattice offset, i.e. the same table stretch: the table loads of a
// warp coalesce to one stretch and stay in L1 instead of each output column streaming all nine tables from L2.
// NB = 3, or 2 without drill data: the drill data block c = 2 has zero weights and the drill property block r = 2 is NaN in the
// reference, so only the four survey blocks are swept (4/9 of the table FMAs).
template <int NB>
__global__ void __launch_bounds__(256) kw_kernel(const double* __restrict__ tables, long ext, long C0, int xN, int yN, int zN,
                                                 const double* __restrict__ w, long Kp, long c0, long nseg, double* __restrict__ z,
                                                 long ncp) {
    __shared__ double red[4][64][13];
    const int tid = threadIdx.x, quad = tid & 3, cl = (tid >> 2) & 15, dg = tid >> 6;
    const long seg = (long)blockIdx.x * 16 + cl;
    const bool live = seg < nseg;
    const long i_base = c0 + (live ? seg : 0) * 16;
    const int ci = (int)(i_base / zN), iz0 = (int)(i_base % zN) + 4 * quad;
    const int lci = (ci / xN) * (2 * xN - 1) + ci % xN;
    const int ncolumns = xN * yN, zs = 2 * zN - 1;
    // sweep range of this block (blockIdx.y of gridDim.y slices, for small shards) split over the 4 thread groups
    const int slice = (ncolumns + gridDim.y - 1) / gridDim.y, sl0 = blockIdx.y * slice, sl1 = min(ncolumns, sl0 + slice);
    const int per = (max(sl1 - sl0, 0) + 3) / 4, sw0 = sl0 + dg * per, sw1 = min(sl1, sw0 + per);
    z += (long)blockIdx.y * 3 * ncp;
    double acc[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0.0;
    if (NB == 3) {
    for (int sw = sw0; sw < sw1; ++sw) {
        int cj = sw + ci;
        if (cj >= ncolumns) cj -= ncolumns;
        const int lcj = (cj / xN) * (2 * xN - 1) + cj % xN;
        const long base = C0 + (long)(lcj - lci) * zs - iz0;      // table offset of (jz = 0, this thread's first output)
        const double* wc = w + (long)cj * zN;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double* wv = wc + (long)c * Kp;
            const double* t0 = tables + (long)(c * 3 + 0) * ext + base;
            const double* t1 = tables + (long)(c * 3 + 1) * ext + base;
            const double* t2 = tables + (long)(c * 3 + 2) * ext + base;
            // window: a_k = table[jz - k]
            double a1 = __ldg(t0 - 1), a2 = __ldg(t0 - 2), a3 = __ldg(t0 - 3);
            double b1 = __ldg(t1 - 1), b2 = __ldg(t1 - 2), b3 = __ldg(t1 - 3);
            double d1 = __ldg(t2 - 1), d2 = __ldg(t2 - 2), d3 = __ldg(t2 - 3);
#pragma unroll 4
            for (int jz = 0; jz < zN; ++jz) {
                const double x = __ldg(wv + jz);
                const double a0 = __ldg(t0 + jz), b0 = __ldg(t1 + jz), d0 = __ldg(t2 + jz);
                acc[0][0] = fma(a0, x, acc[0][0]); acc[0][1] = fma(a1, x, acc[0][1]); acc[0][2] = fma(a2, x, acc[0][2]); acc[0][3] = fma(a3, x, acc[0][3]);
                acc[1][0] = fma(b0, x, acc[1][0]); acc[1][1] = fma(b1, x, acc[1][1]); acc[1][2] = fma(b2, x, acc[1][2]); acc[1][3] = fma(b3, x, acc[1][3]);
                acc[2][0] = fma(d0, x, acc[2][0]); acc[2][1] = fma(d1, x, acc[2][1]); acc[2][2] = fma(d2, x, acc[2][2]); acc[2][3] = fma(d3, x, acc[2][3]);
                a3 = a2; a2 = a1; a1 = a0;
                b3 = b2; b2 = b1; b1 = b0;
                d3 = d2; d2 = d1; d1 = d0;
            }
        }
    }
    } else {
    for (int sw = sw0; sw < sw1; ++sw) {
        int cj = sw + ci;
        if (cj >= ncolumns) cj -= ncolumns;
        const int lcj = (cj / xN) * (2 * xN - 1) + cj % xN;
        const long base = C0 + (long)(lcj - lci) * zs - iz0;
        const double* wc = w + (long)cj * zN;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const double* wv = wc + (long)c * Kp;
            const double* t0 = tables + (long)(c * 3 + 0) * ext + base;
            const double* t1 = tables + (long)(c * 3 + 1) * ext + base;
            double a1 = __ldg(t0 - 1), a2 = __ldg(t0 - 2), a3 = __ldg(t0 - 3);
            double b1 = __ldg(t1 - 1), b2 = __ldg(t1 - 2), b3 = __ldg(t1 - 3);
#pragma unroll 4
            for (int jz = 0; jz < zN; ++jz) {
                const double x = __ldg(wv + jz);
                const double a0 = __ldg(t0 + jz), b0 = __ldg(t1 + jz);
                acc[0][0] = fma(a0, x, acc[0][0]); acc[0][1] = fma(a1, x, acc[0][1]); acc[0][2] = fma(a2, x, acc[0][2]); acc[0][3] = fma(a3, x, acc[0][3]);
                acc[1][0] = fma(b0, x, acc[1][0]); acc[1][1] = fma(b1, x, acc[1][1]); acc[1][2] = fma(b2, x, acc[1][2]); acc[1][3] = fma(b3, x, acc[1][3]);
                a3 = a2; a2 = a1; a1 = a0;
                b3 = b2; b2 = b1; b1 = b0;
            }
        }
    }
    }
#pragma unroll
    for (int r = 0; r < NB; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) red[dg][tid & 63][r * 4 + q] = acc[r][q];
    __syncthreads();
    for (int o = tid; o < 64 * 4 * NB; o += 256) {
        const int lq = o / (4 * NB), rq = o % (4 * NB), r = rq >> 2, q = rq & 3;
        const long sg = (long)blockIdx.x * 16 + (lq >> 2);
        if (sg < nseg)      // fixed order over the four sweep groups: deterministic
            z[(long)r * ncp + sg * 16 + (lq & 3) * 4 + q] = (red[0][lq][rq] + red[1][lq][rq]) + (red[2][lq][rq] + red[3][lq][rq]);
    }
}

// t[c*Ns + s] = sum_{j < ncol} A_c[s][c0 + j] z[c][j]   (warp per row)
__global__ void __launch_bounds__(256) a_gemv_kernel(const double* __restrict__ A0, const double* __restrict__ A1, long Ns, long lda, long c0,
                                                     long ncol, const double* __restrict__ z, long ncp, double* __restrict__ t,
                                                     long zoff, int accumulate) {
    // c0: first column inside A's rows, zoff: first entry of z[c] (they differ when A is a column chunk of the matrix)
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, c = blockIdx.y;
    if (row >= Ns) return;
    const double* a = (c ? A1 : A0) + row * lda + c0;
    const double* zc = z + (long)c * ncp + zoff;
    double s0 = 0.0, s1 = 0.0;
    long j = lane;
    for (; j + 32 < ncol; j += 64) {
        s0 = fma(a[j], zc[j], s0);
        s1 = fma(a[j + 32], zc[j + 32], s1);
    }
    if (j < ncol) s0 = fma(a[j], zc[j], s0);
    double s = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) t[c * Ns + row] = accumulate ? t[c * Ns + row] + s : s;
}

// drill rows of A3 z: z[2][drill - c0] when the drilled voxel is in this rank's shard, else 0 (other ranks add it)
__global__ void t_drill_kernel(const double* __restrict__ z2, const int64_t* __restrict__ drill, long nd, long c0, long c1,
                               double* __restrict__ t_drill) {
    const long d = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= nd) return;
    const long j = drill[d];
    t_drill[d] = (j >= c0 && j < c1) ? z2[j - c0] : 0.0;
}

// r = y - t - sigma^2 alpha  (pad rows: 0)
__global__ void residual_kernel(const double* __restrict__ y, const double* __restrict__ t, const double* __restrict__ alpha, long Ns,
                                long M, long Mp, double s0, double s1, double s2, double* __restrict__ r) {
    const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mp) return;
    if (m >= M) { r[m] = 0.0; return; }
    const double sg = m < Ns ? s0 : (m < 2 * Ns ? s1 : s2);
    r[m] = y[m] - t[m] - sg * sg * alpha[m];
}

// v[m] = sum_{k <= m} Linv[m][k] x[k * xs]   (warp per row)
__global__ void __launch_bounds__(256) linv_gemv_kernel(const double* __restrict__ Linv, long n, const double* __restrict__ x, int xs,
                                                        double* __restrict__ v) {
    const long m = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= n) return;
    const double* row = Linv + m * n;
    double s = 0.0;
    for (long k = lane; k <= m; k += 32) s = fma(row[k], x[k * xs], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) v[m] = s;
}

// out[k] (+)= sum_{m >= k} Linv[m][k] x[m * xs]   (32 columns x 8 row groups per block, fixed-order reduction)
__global__ void __launch_bounds__(256) linv_t_gemv_kernel(const double* __restrict__ Linv, long n, const double* __restrict__ x, int xs,
                                                          double* __restrict__ out, int accumulate) {
    __shared__ double red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long k = (long)blockIdx.x * 32 + tx;
    double s = 0.0;
    if (k < n)
        for (long m = (long)blockIdx.x * 32 + ty; m < n; m += 8)
            if (m >= k) s = fma(Linv[m * n + k], x[m * xs], s);
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && k < n) {
        double tot = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) tot += red[q][tx];
        out[k] = accumulate ? out[k] + tot : tot;
    }
}

// z[e] = sum over the sweep slices (fixed order)
__global__ void kw_reduce_kernel(double* __restrict__ z, long count, int nslice) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    double s = z[e];
    for (int k = 1; k < nslice; ++k) s += z[(long)k * count + e];
    z[e] = s;
}

// out[0] = sum_m a[m] b[m]  (one block, fixed order)
__global__ void dot_kernel(const double* __restrict__ a, const double* __restrict__ b, long n, double* __restrict__ out) {
    __shared__ double red[256];
    double acc = 0.0;
    for (long m = threadIdx.x; m < n; m += blockDim.x) acc = fma(a[m], b[m], acc);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

__global__ void scatter_mu_kernel(const double* __restrict__ z, long ncp, long ncol, double* __restrict__ mu) {
    const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (col < ncol) mu[r * ncol + col] = z[(long)r * ncp + col];
}

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t refine_at_alpha(const RefineArgs& a, const double* alpha, double* w, cudaStream_t s) {
    dim3 grid((unsigned)((a.N + 255) / 256), (unsigned)a.nsplit, 2);
    at_gemv_partial_kernel<<<grid, 256, 0, s>>>(a.A[0], a.A[1], a.Ns, a.N, a.lda, alpha, a.nsplit, a.partial, a.Kp);
    w_reduce_kernel<<<(unsigned)((a.Kp + 255) / 256), 256, 0, s>>>(a.partial, a.nsplit, a.Kp, a.N, w);
    if (a.nd) w_drill_kernel<<<(unsigned)((a.nd + 255) / 256), 256, 0, s>>>(alpha + 2 * a.Ns, a.drill, a.nd, w + 2 * a.Kp);
    return cudaGetLastError();
}

// sweep slices per output block so that small voxel-column shards (multi-GPU) still fill the SMs; z needs room for
// refine_kw_slices(ncol) copies of [3][ncp]
int refine_kw_slices(long ncol) {
    const long nbx = (ncol / 16 + 15) / 16;
    long k = (296 + nbx - 1) / (nbx > 0 ? nbx : 1);
    return (int)(k < 1 ? 1 : (k > 8 ? 8 : k));
}

cudaError_t refine_kw(const RefineArgs& a, const double* w, double* z, cudaStream_t s) {
    const long nseg = a.ncol / 16;
    const unsigned nbx = (unsigned)((nseg + 15) / 16);
    const int nslice = refine_kw_slices(a.ncol);
    dim3 grid(nbx, (unsigned)nslice);
    if (a.nprop == 2) kw_kernel<2><<<grid, 256, 0, s>>>(a.tables, a.ext, a.C0, a.n[0], a.n[1], a.n[2], w, a.Kp, a.c0, nseg, z, a.ncp);
    else kw_kernel<3><<<grid, 256, 0, s>>>(a.tables, a.ext, a.C0, a.n[0], a.n[1], a.n[2], w, a.Kp, a.c0, nseg, z, a.ncp);
    if (nslice > 1) kw_reduce_kernel<<<(unsigned)((3 * a.ncp + 255) / 256), 256, 0, s>>>(z, 3 * a.ncp, nslice);
    return cudaGetLastError();
}

cudaError_t refine_a_z(const RefineArgs& a, const double* z, double* t, cudaStream_t s) {
    dim3 grid((unsigned)((a.Ns + 7) / 8), 2);
    a_gemv_kernel<<<grid, 256, 0, s>>>(a.A[0], a.A[1], a.Ns, a.lda, a.c0, a.ncol, z, a.ncp, t, 0, 0);
    return refine_a_z_drill(a, z, t, s);
}

cudaError_t refine_a_z_drill(const RefineArgs& a, const double* z, double* t, cudaStream_t s) {
    if (a.nd) t_drill_kernel<<<(unsigned)((a.nd + 255) / 256), 256, 0, s>>>(z + 2 * a.ncp, a.drill, a.nd, a.c0, a.c0 + a.ncol, t + 2 * a.Ns);
    return cudaGetLastError();
}

cudaError_t refine_a_z_chunk(const RefineArgs& a, const double* A0, const double* A1, long ld, long j0, long ja, long jb, const double* z,
                             double* t, int accumulate, cudaStream_t s) {
    if (jb <= ja) return cudaSuccess;
    dim3 grid((unsigned)((a.Ns + 7) / 8), 2);
    a_gemv_kernel<<<grid, 256, 0, s>>>(A0, A1, a.Ns, ld, ja - j0, jb - ja, z, a.ncp, t, ja - a.c0, accumulate);
    return cudaGetLastError();
}

cudaError_t refine_at_alpha_chunk(const RefineArgs& a, const double* A0, const double* A1, long ld, long j0, long ncols, const double* alpha,
                                  cudaStream_t s) {
    dim3 grid((unsigned)((ncols + 255) / 256), (unsigned)a.nsplit, 2);
    at_gemv_partial_kernel<<<grid, 256, 0, s>>>(A0, A1, a.Ns, ncols, ld, alpha, a.nsplit, a.partial + j0, a.Kp);
    return cudaGetLastError();
}

// w[2][drill] = alpha_drill (after an all-reduce of the two survey blocks of w, which must not sum the replicated drill block)
cudaError_t refine_at_alpha_finish_drill(const RefineArgs& a, const double* alpha, double* w, cudaStream_t s) {
    if (a.nd) w_drill_kernel<<<(unsigned)((a.nd + 255) / 256), 256, 0, s>>>(alpha + 2 * a.Ns, a.drill, a.nd, w + 2 * a.Kp);
    return cudaGetLastError();
}

cudaError_t refine_at_alpha_finish(const RefineArgs& a, const double* alpha, double* w, cudaStream_t s) {
    w_reduce_kernel<<<(unsigned)((a.Kp + 255) / 256), 256, 0, s>>>(a.partial, a.nsplit, a.Kp, a.N, w);
    if (a.nd) w_drill_kernel<<<(unsigned)((a.nd + 255) / 256), 256, 0, s>>>(alpha + 2 * a.Ns, a.drill, a.nd, w + 2 * a.Kp);
    return cudaGetLastError();
}

cudaError_t refine_residual(const double* y, const double* t, const double* alpha, long Ns, long M, long Mp, const double sigma[3],
                            double* r, cudaStream_t s) {
    residual_kernel<<<(unsigned)((Mp + 255) / 256), 256, 0, s>>>(y, t, alpha, Ns, M, Mp, sigma[0], sigma[1], sigma[2], r);
    return cudaGetLastError();
}

// out (+)= Linv^T (Linv x)
cudaError_t refine_apply_inverse(const double* Linv, long Mp, const double* x, double* tmp, double* out, int accumulate, cudaStream_t s) {
    linv_gemv_kernel<<<(unsigned)((Mp + 7) / 8), 256, 0, s>>>(Linv, Mp, x, 1, tmp);
    linv_t_gemv_kernel<<<(unsigned)((Mp + 31) / 32), 256, 0, s>>>(Linv, Mp, tmp, 1, out, accumulate);
    return cudaGetLastError();
}

cudaError_t refine_linv_t(const double* Linv, long Mp, const double* x, int xs, double* out, cudaStream_t s) {
    linv_t_gemv_kernel<<<(unsigned)((Mp + 31) / 32), 256, 0, s>>>(Linv, Mp, x, xs, out, 0);
    return cudaGetLastError();
}

cudaError_t refine_dot(const double* a, const double* b, long n, double* out, cudaStream_t s) {
    dot_kernel<<<1, 256, 0, s>>>(a, b, n, out);
    return cudaGetLastError();
}

cudaError_t refine_scatter_mu(const double* z, long ncp, long ncol, double* mu, cudaStream_t s, int nr) {
    dim3 grid((unsigned)((ncol + 255) / 256), (unsigned)nr);
    scatter_mu_kernel<<<grid, 256, 0, s>>>(z, ncp, ncol, mu);
    return cudaGetLastError();
}
