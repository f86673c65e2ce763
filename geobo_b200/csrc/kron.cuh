// Separable (Kronecker) form of the squared-exponential covariance blocks -- SURVEY.md section 8(f) row 3.
//
// On the regular voxel grid every block of kernels.create_cov with fkernel = 'exp' (geobo/kernels.py:81-99, :158-195) is
//     K_cr[(iy,ix,iz),(jy,jx,jz)] = t0 * fy(jy-iy) * fx(jx-ix) * fz(jz-iz),        exp(-(dx^2+dy^2+dz^2)/s) = product of three
// so a product with K_cr is three small Toeplitz mode products instead of one N x N contraction:
//     (A K_cr)[s, (jy,jx,jz)] = sum_iy fy'(jy-iy) sum_ix fx(jx-ix) sum_iz fz(jz-iz) A[s, (iy,ix,iz)]
// 2 N (xN + yN + zN) flops per row instead of 2 N^2.  The factor lines are read off the stationary covariance tables
// (the table values along the three axes through the zero offset), so every kernel family / weight / amplitude rule of
// create_cov stays in ONE place (cov.cu); fy' carries the 1 / t0^2 that makes the product of the three lines equal K.
//
// This header holds the per-thread arithmetic of the two kernels of kron.cu, written against a thread id so that the
// CPU suite can compile it with g++ and run the exact index arithmetic in loops (tests/host_harness/kron_host.cpp).
// It is an OPT-IN fast path (gb_hyper.structure = GB_STRUCTURE_KRON); the dense contraction stays the default and the
// roofline numbers of DESIGN.md section 4 are those of the dense path.
#pragma once
#ifndef GB_HD
#if defined(__CUDACC__)
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif
#endif

constexpr int KRON_JT = 8;          // y-mode kernel: output y-rows per thread
constexpr int KRON_YQ = 2;          // y-mode kernel: voxels of the x-z plane per thread
constexpr int KRON_YTHREADS = 128;  // y-mode kernel: threads per block (one block covers KRON_YTHREADS * KRON_YQ plane voxels)
constexpr int KRON_W = 4;           // z/x-mode kernel: outputs per thread along the transformed axis (sliding factor window)
constexpr int KRON_ZXTHREADS = 256;
constexpr int KRON_PAD = 8;         // zero entries behind every factor line (windows of ragged tiles read past 2n-2)

struct KronGeom {
    int xN, yN, zN;
    long XZ;          // xN * zN: voxels of one y-row (voxel index = (iy * xN + ix) * zN + iz, kernels.py:27-42)
    int jy0, nyl;     // first y-row and number of y-rows touched by this rank's voxel-column shard
    long c0, c1;      // the shard [c0, c1)
    int FL;           // doubles per factor line: 2 * max(xN, yN, zN) - 1 + KRON_PAD
    int zs;           // row stride of the shared-memory planes of the z/x kernel (odd: conflict-free column walks)
};

GB_HD KronGeom kron_geom(long xN, long yN, long zN, long c0, long c1) {
    KronGeom g;
    g.xN = (int)xN; g.yN = (int)yN; g.zN = (int)zN;
    g.XZ = xN * zN;
    g.c0 = c0; g.c1 = c1;
    g.jy0 = (int)(c0 / g.XZ);
    g.nyl = (int)((c1 - 1) / g.XZ) - g.jy0 + 1;
    long m = xN > yN ? xN : yN;
    if (zN > m) m = zN;
    g.FL = (int)(2 * m - 1 + KRON_PAD);
    g.zs = (int)(zN | 1);
    return g;
}

// Factor line entry: block b of the 9 covariance tables, axis 0 = y (scaled by 1 / t0^2), 1 = x, 2 = z; i = offset + n - 1.
// tab points at the zero offset of table b (tables + b * ext + C0).
GB_HD double kron_factor(const double* tab, const KronGeom& g, int axis, int i) {
    const int n = axis == 0 ? g.yN : axis == 1 ? g.xN : g.zN;
    if (i > 2 * n - 2) return 0.0;
    const long stride = axis == 0 ? (long)(2 * g.xN - 1) * (2 * g.zN - 1) : axis == 1 ? (long)(2 * g.zN - 1) : 1L;
    const double v = tab[(long)(i - (n - 1)) * stride];
    if (axis != 0) return v;
    const double t0 = tab[0];
    return t0 != 0.0 ? v / t0 / t0 : 0.0;      // a zero cross weight (kernels.py:181) zeroes the whole block
}

// ---------------------------------------------------------------------------------------------------------- y mode
// T[r][jy - jy0][q] = sum_iy fy_r(jy - iy) * Arow[iy * XZ + q]   for the three property blocks r of one data block,
// q = voxel of the x-z plane.  sf: [3][FL] y factor lines (shared memory on the device).  One thread: KRON_YQ plane
// voxels x KRON_JT output rows x 3 blocks; every input value is loaded once and used 3 * KRON_JT times.
GB_HD void kron_y_thread(const KronGeom& g, const double* Arow, const double* sf, int qtile, int jgroup, int tid, double* T,
                         long r_stride) {
    const long q0 = (long)qtile * (KRON_YTHREADS * KRON_YQ) + tid;
    if (q0 >= g.XZ) return;
    const int jb = g.jy0 + jgroup * KRON_JT;
    bool live[KRON_YQ];
    double acc[3][KRON_JT][KRON_YQ];
#pragma unroll
    for (int u = 0; u < KRON_YQ; ++u) live[u] = q0 + (long)u * KRON_YTHREADS < g.XZ;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int t = 0; t < KRON_JT; ++t)
#pragma unroll
            for (int u = 0; u < KRON_YQ; ++u) acc[r][t][u] = 0.0;
    for (int iy = 0; iy < g.yN; ++iy) {
        double a[KRON_YQ];
#pragma unroll
        for (int u = 0; u < KRON_YQ; ++u) a[u] = live[u] ? Arow[(long)iy * g.XZ + q0 + (long)u * KRON_YTHREADS] : 0.0;
        const int base = jb - iy + g.yN - 1;        // >= 0; rows past yN - 1 read the zero padding of the line
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int t = 0; t < KRON_JT; ++t) {
                const double f = sf[r * g.FL + base + t];
#pragma unroll
                for (int u = 0; u < KRON_YQ; ++u) acc[r][t][u] += f * a[u];
            }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int t = 0; t < KRON_JT; ++t) {
            const int jl = jb + t - g.jy0;
            if (jl >= g.nyl) continue;
#pragma unroll
            for (int u = 0; u < KRON_YQ; ++u)
                if (live[u]) T[r * r_stride + (long)jl * g.XZ + q0 + (long)u * KRON_YTHREADS] = acc[r][t][u];
        }
}

// ---------------------------------------------------------------------------------------------------------- z and x modes
// One block owns one x-z plane (xN x zN values of one y-row): phase A stages it, phase B applies the z mode, phase C the
// x mode and stores the result into the caller's row (the voxel columns of the rank's shard only).
// Shared memory: pin[xN][zs], ptmp[xN][zs], fz[FL], fx[FL].

// phase A: plane and the two factor lines -> shared memory
GB_HD void kron_zx_load(const KronGeom& g, const double* plane, const double* kf_x, const double* kf_z, int tid, int nthreads,
                        double* pin, double* fx, double* fz) {
    for (long i = tid; i < g.XZ; i += nthreads) pin[(i / g.zN) * g.zs + (i % g.zN)] = plane[i];
    for (int i = tid; i < g.FL; i += nthreads) {
        fx[i] = kf_x[i];
        fz[i] = kf_z[i];
    }
}

// One strip of KRON_W consecutive outputs o0 .. o0+3 along an axis of length n:  out[o] = sum_i f(o - i) * in[i * istride].
// The factor window slides by one entry per input index (one new factor load per KRON_W multiply-adds).
GB_HD void kron_strip(const double* in, long istride, const double* f, int n, int o0, double acc[KRON_W]) {
    double w[KRON_W];
#pragma unroll
    for (int t = 0; t < KRON_W; ++t) {
        w[t] = f[o0 + t + n - 1];
        acc[t] = 0.0;
    }
    for (int i = 0; i < n; ++i) {
        const double a = in[(long)i * istride];
#pragma unroll
        for (int t = 0; t < KRON_W; ++t) acc[t] += w[t] * a;
        if (i + 1 < n) {
#pragma unroll
            for (int t = KRON_W - 1; t > 0; --t) w[t] = w[t - 1];
            w[0] = f[o0 - (i + 1) + n - 1];
        }
    }
}

// phase B: ptmp[ix][jz] = sum_iz fz(jz - iz) * pin[ix][iz].  Work item = (strip of 4 jz, ix), ix fastest across lanes.
GB_HD void kron_z_phase(const KronGeom& g, const double* pin, const double* fz, int tid, int nthreads, double* ptmp) {
    const int nstrip = (g.zN + KRON_W - 1) / KRON_W;
    for (int w = tid; w < nstrip * g.xN; w += nthreads) {
        const int ix = w % g.xN, jz0 = (w / g.xN) * KRON_W;
        double acc[KRON_W];
        kron_strip(pin + (long)ix * g.zs, 1, fz, g.zN, jz0, acc);
#pragma unroll
        for (int t = 0; t < KRON_W; ++t)
            if (jz0 + t < g.zN) ptmp[(long)ix * g.zs + jz0 + t] = acc[t];
    }
}

// phase C: out[(jy, jx, jz)] = sum_ix fx(jx - ix) * ptmp[ix][jz]  -> row[voxel column - c0] for columns inside the shard.
// Work item = (strip of 4 jx, jz), jz fastest across lanes (coalesced stores, conflict-free plane reads).
GB_HD void kron_x_phase(const KronGeom& g, const double* ptmp, const double* fx, int jy, int tid, int nthreads, double* row,
                        int accumulate) {
    const int nstrip = (g.xN + KRON_W - 1) / KRON_W;
    for (int w = tid; w < nstrip * g.zN; w += nthreads) {
        const int jz = w % g.zN, jx0 = (w / g.zN) * KRON_W;
        double acc[KRON_W];
        kron_strip(ptmp + jz, g.zs, fx, g.xN, jx0, acc);
#pragma unroll
        for (int t = 0; t < KRON_W; ++t) {
            if (jx0 + t >= g.xN) continue;
            const long j = (long)jy * g.XZ + (long)(jx0 + t) * g.zN + jz;
            if (j < g.c0 || j >= g.c1) continue;
            if (accumulate) row[j - g.c0] += acc[t];
            else row[j - g.c0] = acc[t];
        }
    }
}
