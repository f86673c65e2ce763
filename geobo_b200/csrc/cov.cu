// GP covariance functions (geobo/kernels.py) on the device, fp64, formula-for-formula.
//
//  * cov_value():            the nine blocks of kernels.create_cov (kernels.py:158-195)
//  * cov_tables_kernel:      stationary tables over the EXTENDED difference lattice
//                            (2yN-1)(2xN-1)(2zN-1): K[(c,j),(r,i)] = tab[c][r][L(i) - L(j) + C0],
//                            L(v) = (vy (2xN-1) + vx)(2zN-1) + vz -- one integer subtraction per
//                            covariance element inside the fused projection GEMM.
//  * create_cov_dense_kernel: kernels.create_cov for an arbitrary user D2 (API compatibility).
//  * create_cov_grid_kernel:  the same 3N x 3N matrix straight from the grid spec, staged in shared
//                            memory and written with bulk async copies (UBLKCP) -- HBM-write bound.
#include "common.cuh"
#include "formulas.cuh"

// ------------------------------------------------------------------------------------ tables
__global__ void cov_tables_kernel(CovParams P, int xN, int yN, int zN, double sx, double sy, double sz, double* tables, int* nonfinite) {
    const int EX = 2 * xN - 1, EY = 2 * yN - 1, EZ = 2 * zN - 1;
    const long ext = (long)EX * EY * EZ;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ext) return;
    const int ez = (int)(e % EZ);
    const long t = e / EZ;
    const int ex = (int)(t % EX), ey = (int)(t / EX);
    const double D2 = lattice_d2(ex - (xN - 1), ey - (yN - 1), ez - (zN - 1), sx, sy, sz);
    const int cr = blockIdx.y;   // c * 3 + r : row-block c (data side), column-block r (voxel side)
    const double v = cov_value(P, cr / 3, cr % 3, D2);
    tables[(long)cr * ext + e] = v;
    // a NaN / inf anywhere in kcov poisons the reference's dense Asens3 . kcov . Asens3^T (0 * NaN), so its Cholesky always fails
    // (inversion.py:98-104) -- also when the block-structured products here would never touch the value (e.g. nd = 0)
    if (nonfinite && !isfinite(v)) *nonfinite = 1;
}

cudaError_t launch_cov_tables(const CovParams& cp, const int64_t n[3], const double vox[3], double* tables, cudaStream_t s, int* nonfinite) {
    const long ext = (2 * n[0] - 1) * (2 * n[1] - 1) * (2 * n[2] - 1);
    dim3 grid((unsigned)((ext + 255) / 256), 9);
    cov_tables_kernel<<<grid, 256, 0, s>>>(cp, (int)n[0], (int)n[1], (int)n[2], vox[0], vox[1], vox[2], tables, nonfinite);
    return cudaGetLastError();
}

// L(v) for every voxel (flat order (iy*xN+ix)*zN+iz); entries >= N repeat voxel 0 (zero-padded operand columns)
__global__ void lattice_ids_kernel(int xN, int yN, int zN, int* L, long n_padded) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_padded) return;
    const long N = (long)xN * yN * zN;
    const long vv = v < N ? v : 0;
    const int iz = (int)(vv % zN);
    const long t = vv / zN;
    const int ix = (int)(t % xN), iy = (int)(t / xN);
    L[v] = (iy * (2 * xN - 1) + ix) * (2 * zN - 1) + iz;
}

cudaError_t launch_lattice_ids(const int64_t n[3], int* L, int64_t n_padded, cudaStream_t s) {
    lattice_ids_kernel<<<(unsigned)((n_padded + 255) / 256), 256, 0, s>>>((int)n[0], (int)n[1], (int)n[2], L, n_padded);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ dense create_cov(D2, ...)
// out[(r*n + i) * 3n + c*n + j] = block(r, c)(D2[i, j]); one thread per D2 element, 9 coalesced stores.
__global__ void create_cov_dense_kernel(CovParams P, const double* __restrict__ D2, long n, double* __restrict__ out) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long i = blockIdx.y;
    if (j >= n) return;
    const double d2 = D2[i * n + j];
    const long ld = 3 * n;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) out[(r * n + i) * ld + c * n + j] = cov_value(P, r, c, d2);
}

cudaError_t launch_create_cov_dense(const CovParams& cp, const double* D2, int64_t n, double* out, cudaStream_t s) {
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
    create_cov_dense_kernel<<<grid, 256, 0, s>>>(cp, D2, n, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ dense assembly from the grid spec
// One CTA per (row i of one row-block r): the row's voxel coordinates are decoded once, the 3N
// entries are produced in shared memory in chunks and written with cp.async.bulk (TMA 1-D bulk
// store, SASS UBLKCP) so the SM never stalls on the store path.
constexpr int ASM_CHUNK = 2048;   // doubles per staged chunk (16 KB), double buffered

__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, unsigned bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(256) create_cov_grid_kernel(CovParams P, int xN, int yN, int zN, double sx, double sy,
                                                              double sz, double* __restrict__ out) {
    __shared__ __align__(128) double stage[2][ASM_CHUNK];
    const long N = (long)xN * yN * zN;
    const long row = blockIdx.x;            // 0 .. 3N-1
    const int r = (int)(row / N);
    const long i = row % N;
    const int iz = (int)(i % zN);
    const long ti = i / zN;
    const int ix = (int)(ti % xN), iy = (int)(ti / xN);
    double* orow = out + row * 3 * N;
    const long total = 3 * N;
    int buf = 0;
    for (long base = 0; base < total; base += ASM_CHUNK, buf ^= 1) {
        const int len = (int)min((long)ASM_CHUNK, total - base);
        // the bulk store that last read this buffer (two chunks ago) must have finished reading smem
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        for (int q = threadIdx.x; q < len; q += blockDim.x) {
            const long col = base + q;
            const int c = (int)(col / N);
            const long j = col % N;
            const int jz = (int)(j % zN);
            const long tj = j / zN;
            const int jx = (int)(tj % xN), jy = (int)(tj / xN);
            stage[buf][q] = cov_value(P, r, c, lattice_d2(jx - ix, jy - iy, jz - iz, sx, sy, sz));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned bytes = (unsigned)len * 8u;
            if ((bytes & 15u) == 0 && ((((uintptr_t)(orow + base)) & 15) == 0)) {
                bulk_store(orow + base, stage[buf], bytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            } else {
                for (int q = 0; q < len; ++q) orow[base + q] = stage[buf][q];
            }
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Same matrix assembled from the stationary-covariance tables: on the regular grid every block of create_cov depends only
// on the lattice offset, so the 9 block functions are evaluated once per offset (cov_tables_kernel, 9 x 8N values instead
// of 9 N^2 transcendental evaluations -- bit-identical, the same cov_value / lattice_d2 are used) and the dense 3N x 3N
// matrix is a gather  out[(r,i),(c,j)] = tab[r*3+c][C0 + L(j) - L(i)].  One CTA owns the zN rows of one z-column ci of
// one row block r: for a z-column cj of the contraction side the zN x zN output block is Toeplitz in ONE contiguous
// table stretch of 2 zN - 1 values, so a chunk of CW columns needs CW (2 zN - 1) table loads for CW zN^2 outputs.  The
// zN row segments of the chunk are built in shared memory and written with one bulk async store each (UBLKCP), double
// buffered.  This turns the assembly from fp64-transcendental bound (0.17 of the HBM roofline) into HBM-write bound.
__global__ void __launch_bounds__(256) create_cov_grid_gather_kernel(const double* __restrict__ tables, long ext, long C0, int xN,
                                                                     int yN, int zN, int CW, double* __restrict__ out) {
    extern __shared__ __align__(128) double sm[];
    const int zs = 2 * zN - 1, ncolumns = xN * yN, seg = CW * zN;      // seg: doubles per row segment of one chunk
    double* T = sm;                              // [CW][zs]
    double* stage = sm + ((CW * zs + 15) & ~15); // [2][zN][seg]
    const long N = (long)ncolumns * zN;
    const int r = blockIdx.y, ci = blockIdx.x;
    const int lci = (ci / xN) * (2 * xN - 1) + ci % xN;
    const int nchunk = (ncolumns + CW - 1) / CW;
    int buf = 0;
    for (int c = 0; c < 3; ++c) {
        const double* tab = tables + (long)(r * 3 + c) * ext + C0 - (zN - 1);
        for (int ch = 0; ch < nchunk; ++ch, buf ^= 1) {
            const int cj0 = ch * CW, ncj = min(CW, ncolumns - cj0);
            // the bulk stores that last read this stage buffer (two chunks ago) must have finished reading shared memory
            // (bulk groups are tracked per issuing thread: every one of the zN issuers waits for its own)
            if (threadIdx.x < zN) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
            for (int e = threadIdx.x; e < ncj * zs; e += blockDim.x) {
                const int cl = e / zs, k = e - cl * zs, cj = cj0 + cl;
                const int lcj = (cj / xN) * (2 * xN - 1) + cj % xN;
                T[cl * zs + k] = __ldg(tab + (long)(lcj - lci) * zs + k);      // table offsets jz - iz = k - (zN - 1)
            }
            __syncthreads();
            double* st = stage + (long)buf * zN * seg;
            for (int e = threadIdx.x; e < ncj * zN; e += blockDim.x) {
                const int cl = e / zN, jz = e - cl * zN;
                const double* t = T + cl * zs + jz + (zN - 1);
                for (int iz = 0; iz < zN; ++iz) st[iz * seg + e] = t[-iz];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x < zN) {
                const int iz = threadIdx.x;
                double* dst = out + ((long)r * N + (long)ci * zN + iz) * 3 * N + (long)c * N + (long)cj0 * zN;
                const unsigned bytes = (unsigned)(ncj * zN) * 8u;
                if ((bytes & 15u) == 0 && (((uintptr_t)dst) & 15) == 0) {
                    bulk_store(dst, st + iz * seg, bytes);
                } else {
                    for (int q = 0; q < ncj * zN; ++q) dst[q] = st[iz * seg + q];
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (threadIdx.x < zN) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// direct = true: evaluate the covariance function per element (reference order of work); false: tables + gather
cudaError_t launch_create_cov_grid(const CovParams& cp, const int64_t n[3], const double vox[3], double* out, cudaStream_t s,
                                   double* tables, int* L) {
    const long N = n[0] * n[1] * n[2];
    if (!tables || !L) {
        create_cov_grid_kernel<<<(unsigned)(3 * N), 256, 0, s>>>(cp, (int)n[0], (int)n[1], (int)n[2], vox[0], vox[1], vox[2], out);
        return cudaGetLastError();
    }
    const long ext = (2 * n[0] - 1) * (2 * n[1] - 1) * (2 * n[2] - 1);
    const long C0 = ((n[1] - 1) * (2 * n[0] - 1) + (n[0] - 1)) * (2 * n[2] - 1) + (n[2] - 1);
    cudaError_t e = launch_cov_tables(cp, n, vox, tables, s);
    if (e != cudaSuccess) return e;
    const int zN = (int)n[2], ncolumns = (int)(n[0] * n[1]);
    if (zN > 256) return cudaErrorInvalidValue;
    int CW = 4096 / (zN * zN);
    if (CW < 1) CW = 1;
    if (CW > ncolumns) CW = ncolumns;
    const int smem = (((CW * (2 * zN - 1) + 15) & ~15) + 2 * zN * CW * zN) * (int)sizeof(double);
    e = cudaFuncSetAttribute(create_cov_grid_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)ncolumns, 3);
    create_cov_grid_gather_kernel<<<grid, 256, smem, s>>>(tables, ext, C0, (int)n[0], (int)n[1], zN, CW, out);
    (void)L;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ calcGridPoints3D / calcDistanceMatrix
__global__ void grid_points_kernel(int xN, int yN, int zN, double sx, double sy, double sz, double* out) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long N = (long)xN * yN * zN;
    if (v >= N) return;
    const int iz = (int)(v % zN);
    const long t = v / zN;
    const int ix = (int)(t % xN), iy = (int)(t / xN);
    out[3 * v + 0] = __dmul_rn((double)(ix + 1), sx);   // kernels.py:37-39: arange(1, n+1) * scale
    out[3 * v + 1] = __dmul_rn((double)(iy + 1), sy);
    out[3 * v + 2] = __dmul_rn((double)(iz + 1), sz);
}

cudaError_t launch_grid_points(const int64_t lpix[3], const double sc[3], double* out, cudaStream_t s) {
    const long N = lpix[0] * lpix[1] * lpix[2];
    grid_points_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>((int)lpix[0], (int)lpix[1], (int)lpix[2], sc[0], sc[1], sc[2], out);
    return cudaGetLastError();
}

__global__ void sqdist_kernel(const double* __restrict__ pts, long n, int dim, double* __restrict__ out) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long i = blockIdx.y;
    if (j >= n) return;
    double acc = 0.0;
    for (int d = 0; d < dim; ++d) {           // kernels.py:46,54-58: sum over dimensions in order
        const double delta = __dsub_rn(pts[j * dim + d], pts[i * dim + d]);
        acc = __dadd_rn(acc, __dmul_rn(delta, delta));
    }
    out[i * n + j] = acc;
}

cudaError_t launch_sqdist(const double* pts, int64_t n, int dim, double* out, cudaStream_t s) {
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
    sqdist_kernel<<<grid, 256, 0, s>>>(pts, n, dim, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ elementwise API kernels
// kernels.gpkernel / gpkernel2 / gpkernel_sparse / gpkernel_sparse2 / gpkernel_matern32 / gpkernel_matern32_2
// (kernels.py:81-156) applied to an arbitrary array of squared distances.
__global__ void cov_function_kernel(int kernel_id, int cross, const double* __restrict__ D2, long count, double l1, double l2,
                                    double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double d2 = D2[i];
    double v;
    if (!cross)
        v = kernel_id == GB_KERNEL_EXP ? k_exp_same(d2, l1) : kernel_id == GB_KERNEL_MATERN32 ? k_matern_same(d2, l1) : k_sparse_same(d2, l1);
    else
        v = kernel_id == GB_KERNEL_EXP ? k_exp_cross(d2, l1, l2)
            : kernel_id == GB_KERNEL_MATERN32 ? k_matern_cross(d2, l1, l2) : k_sparse_cross(d2, l1, l2);
    out[i] = v;
}

cudaError_t launch_cov_function(int kernel_id, int cross, const double* D2, int64_t count, double l1, double l2, double* out,
                                cudaStream_t s) {
    cov_function_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(kernel_id, cross, D2, count, l1, l2, out);
    return cudaGetLastError();
}
