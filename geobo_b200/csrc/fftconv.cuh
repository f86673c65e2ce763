// Block-Toeplitz (FFT) form of the stationary covariance blocks -- SURVEY.md section 8(f) row 3, any kernel family.
//
// On the voxel grid every block of kernels.create_cov (geobo/kernels.py:158-195) depends only on the offset between two
// voxels, K_cr[i, j] = tab_cr[L(j) - L(i)], so a product with K_cr is a 3-D linear convolution with the (even) table,
// evaluated exactly -- up to fp64 rounding -- by circulant embedding: zero-pad every axis to P >= 2n - 1, transform,
// multiply with the (real) spectrum of the wrapped table, transform back, keep the first n outputs.  Two real rows ride on
// one complex transform (row 2b in the real part, row 2b+1 in the imaginary part; a real spectrum keeps them apart).
//     cost per row: O(P^3 log P)  instead of 2 N^2.
// Lengths are powers of two; transforms are radix-2 decimation-in-time in shared memory, LPB lines per block.
//
// This header holds the per-thread arithmetic of the kernels of fftconv.cu, written against a thread id so that the CPU
// suite can compile it with g++ and run the exact index arithmetic phase by phase (tests/host_harness/fftconv_host.cpp).
// OPT-IN (gb_hyper.structure = GB_STRUCTURE_FFT); the dense contraction stays the default.
#pragma once
#ifndef GB_HD
#if defined(__CUDACC__)
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif
#endif

constexpr int FFT_LPB = 16;        // lines per block (adjacent lines of the strided passes are adjacent in memory: 256 B segments)
constexpr int FFT_THREADS = 256;
constexpr int FFT_MAXP = 512;

struct alignas(16) cplx {
    double re, im;
};

GB_HD int fft_pow2_at_least(long n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}
GB_HD int fft_log2(int p) {
    int l = 0;
    while ((1 << l) < p) ++l;
    return l;
}
GB_HD int fft_bitrev(int k, int logP) {
    int r = 0;
    for (int b = 0; b < logP; ++b) r |= ((k >> b) & 1) << (logP - 1 - b);
    return r;
}

struct FftGeom {
    int xN, yN, zN;
    int Px, Py, Pz;   // padded lengths: powers of two >= 2n - 1
    long XZ, P3;      // xN * zN, Py * Px * Pz
    int jy0, nyl;     // y-rows touched by the rank's voxel-column shard [c0, c1)
    long c0, c1;
};

GB_HD FftGeom fft_geom(long xN, long yN, long zN, long c0, long c1) {
    FftGeom g;
    g.xN = (int)xN; g.yN = (int)yN; g.zN = (int)zN;
    g.Px = fft_pow2_at_least(2 * xN - 1); g.Py = fft_pow2_at_least(2 * yN - 1); g.Pz = fft_pow2_at_least(2 * zN - 1);
    g.XZ = xN * zN;
    g.P3 = (long)g.Px * g.Py * g.Pz;
    g.c0 = c0; g.c1 = c1;
    g.jy0 = (int)(c0 / g.XZ);
    g.nyl = (int)((c1 - 1) / g.XZ) - g.jy0 + 1;
    return g;
}

// One batched 1-D pass: line l = o * inner + i (i < inner) starts at o * in_outer + i (input) / o * out_outer + i (output),
// element k sits k * stride further.  Inputs k >= nvalid are zeros (not read); outputs [out_first, out_first + nkeep) are
// stored at element index k - out_first.
struct FftPass {
    int P, logP;
    long stride, inner, nlines, in_outer, out_outer;
    int nvalid, out_first, nkeep;
    int inverse;        // conjugated twiddles
    int line_fastest;   // loads / stores: 1 = adjacent threads take adjacent lines (strided passes), 0 = adjacent elements
};

GB_HD void fft_split(const FftPass& q, long idx, int& li, int& k) {
    if (q.line_fastest) { li = (int)(idx % FFT_LPB); k = (int)(idx / FFT_LPB); }
    else { k = (int)(idx % q.P); li = (int)(idx / q.P); }
}

// generic load: smem[li][bitrev(k)] = in[line][k] (* spectrum[k * stride + i] when mul != nullptr)
GB_HD void fft_load(const FftPass& q, const cplx* in, const double* mul, long line0, int tid, int nthreads, cplx* sm) {
    for (long idx = tid; idx < (long)FFT_LPB * q.P; idx += nthreads) {
        int li, k;
        fft_split(q, idx, li, k);
        const long line = line0 + li;
        cplx v;
        v.re = 0.0; v.im = 0.0;
        if (line < q.nlines && k < q.nvalid) {
            const long o = line / q.inner, i = line % q.inner;
            v = in[o * q.in_outer + i + (long)k * q.stride];
            if (mul) {
                const double m = mul[(long)k * q.stride + i];
                v.re *= m; v.im *= m;
            }
        }
        sm[li * q.P + fft_bitrev(k, q.logP)] = v;
    }
}

// stage s = 1 .. logP of the decimation-in-time butterflies; tw[k] = exp(-2 pi i k / P), k < P / 2
GB_HD void fft_stage(const FftPass& q, int s, const cplx* tw, int tid, int nthreads, cplx* sm) {
    const int halfP = q.P >> 1, half = 1 << (s - 1);
    for (int j = tid; j < FFT_LPB * halfP; j += nthreads) {
        const int li = j / halfP, b = j % halfP;
        const int grp = b >> (s - 1), pos = b & (half - 1);
        const int i0 = li * q.P + (grp << s) + pos, i1 = i0 + half;
        cplx w = tw[pos << (q.logP - s)];
        if (q.inverse) w.im = -w.im;
        const cplx a = sm[i0], c = sm[i1];
        cplx t;
        t.re = w.re * c.re - w.im * c.im;
        t.im = w.re * c.im + w.im * c.re;
        cplx lo, hi;
        lo.re = a.re + t.re; lo.im = a.im + t.im;
        hi.re = a.re - t.re; hi.im = a.im - t.im;
        sm[i0] = lo;
        sm[i1] = hi;
    }
}

GB_HD void fft_store(const FftPass& q, cplx* out, long line0, int tid, int nthreads, const cplx* sm) {
    for (long idx = tid; idx < (long)FFT_LPB * q.P; idx += nthreads) {
        int li, k;
        fft_split(q, idx, li, k);
        const long line = line0 + li;
        if (line >= q.nlines || k < q.out_first || k >= q.out_first + q.nkeep) continue;
        const long o = line / q.inner, i = line % q.inner;
        out[o * q.out_outer + i + (long)(k - q.out_first) * q.stride] = sm[li * q.P + k];
    }
}

// forward z pass, loader: line l = (b * yN + y) * xN + x takes the zN values of voxel column (y, x) of rows 2b (real part) and
// 2b + 1 (imaginary part; zero when the row count is odd) of A
GB_HD void fft_load_rows(const FftGeom& g, const FftPass& q, const double* A, long lda, long nrows, long line0, int tid, int nthreads, cplx* sm) {
    const long cols = (long)g.yN * g.xN;
    for (long idx = tid; idx < (long)FFT_LPB * q.P; idx += nthreads) {
        const int k = (int)(idx % q.P), li = (int)(idx / q.P);
        const long line = line0 + li;
        cplx v;
        v.re = 0.0; v.im = 0.0;
        if (line < q.nlines && k < g.zN) {
            const long b = line / cols, v0 = (line % cols) * g.zN;
            v.re = A[(2 * b) * lda + v0 + k];
            if (2 * b + 1 < nrows) v.im = A[(2 * b + 1) * lda + v0 + k];
        }
        sm[li * q.P + fft_bitrev(k, q.logP)] = v;
    }
}

// inverse z pass, storer: line l = (b * nyl + jyl) * xN + jx holds voxel column (jy0 + jyl, jx); outputs k < zN inside the
// rank's shard go to rows 2b / 2b + 1 of `out` (row stride ldo) at [j - c0], scaled by 1 / (Py Px Pz)
GB_HD void fft_store_rows(const FftGeom& g, const FftPass& q, double* out, long ldo, long nrows, int accumulate, long line0, int tid,
                          int nthreads, const cplx* sm) {
    const long cols = (long)g.nyl * g.xN;
    const double scale = 1.0 / (double)g.P3;
    for (long idx = tid; idx < (long)FFT_LPB * q.P; idx += nthreads) {
        const int k = (int)(idx % q.P), li = (int)(idx / q.P);
        const long line = line0 + li;
        if (line >= q.nlines || k >= g.zN) continue;
        const long b = line / cols, rem = line % cols;
        const long j = (g.jy0 + rem / g.xN) * g.XZ + (rem % g.xN) * g.zN + k;
        if (j < g.c0 || j >= g.c1) continue;
        const cplx v = sm[li * q.P + k];
        double* o0 = out + (2 * b) * ldo + (j - g.c0);
        *o0 = accumulate ? *o0 + v.re * scale : v.re * scale;
        if (2 * b + 1 < nrows) {
            double* o1 = out + (2 * b + 1) * ldo + (j - g.c0);
            *o1 = accumulate ? *o1 + v.im * scale : v.im * scale;
        }
    }
}

// wrapped table of one block: element (ky, kx, kz) of the P3 lattice = tab0[offset (dy, dx, dz)] with d = k for k <= n - 1,
// d = k - P for k >= P - (n - 1), zero in between.  tab0: zero offset of the block's stationary table.
GB_HD cplx fft_wrapped_tap(const FftGeom& g, const double* tab0, long e) {
    const int kz = (int)(e % g.Pz), kx = (int)((e / g.Pz) % g.Px), ky = (int)(e / ((long)g.Pz * g.Px));
    cplx v;
    v.re = 0.0; v.im = 0.0;
    const int dz = kz <= g.zN - 1 ? kz : (kz >= g.Pz - (g.zN - 1) ? kz - g.Pz : g.Pz);
    const int dx = kx <= g.xN - 1 ? kx : (kx >= g.Px - (g.xN - 1) ? kx - g.Px : g.Px);
    const int dy = ky <= g.yN - 1 ? ky : (ky >= g.Py - (g.yN - 1) ? ky - g.Py : g.Py);
    if (dz == g.Pz || dx == g.Px || dy == g.Py) return v;
    v.re = tab0[((long)dy * (2 * g.xN - 1) + dx) * (2 * g.zN - 1) + dz];
    return v;
}

// the passes of one product (shared by the launcher and the host harness)
GB_HD FftPass fft_pass(int P, long stride, long inner, long nlines, long in_outer, long out_outer, int nvalid, int out_first, int nkeep,
                       int inverse, int line_fastest) {
    FftPass q;
    q.P = P; q.logP = fft_log2(P);
    q.stride = stride; q.inner = inner; q.nlines = nlines; q.in_outer = in_outer; q.out_outer = out_outer;
    q.nvalid = nvalid; q.out_first = out_first; q.nkeep = nkeep; q.inverse = inverse; q.line_fastest = line_fastest;
    return q;
}
// B = complex row pairs in flight.  Buffers: X, Y [B * P3], Z [B * nyl * xN * Pz].
GB_HD FftPass fft_pass_fwd_z(const FftGeom& g, long B) {      // A rows -> X as [B][yN][xN][Pz]
    return fft_pass(g.Pz, 1, 1, B * g.yN * g.xN, 0, g.Pz, g.zN, 0, g.Pz, 0, 0);
}
GB_HD FftPass fft_pass_fwd_x(const FftGeom& g, long B) {      // X [B][yN][xN][Pz] -> Y [B][yN][Px][Pz]
    return fft_pass(g.Px, g.Pz, g.Pz, B * g.yN * g.Pz, (long)g.xN * g.Pz, (long)g.Px * g.Pz, g.xN, 0, g.Px, 0, 1);
}
GB_HD FftPass fft_pass_fwd_y(const FftGeom& g, long B) {      // Y [B][yN][Px][Pz] -> X [B][Py][Px][Pz]
    const long plane = (long)g.Px * g.Pz;
    return fft_pass(g.Py, plane, plane, B * plane, g.yN * plane, g.Py * plane, g.yN, 0, g.Py, 0, 1);
}
GB_HD FftPass fft_pass_inv_y(const FftGeom& g, long B) {      // X * spectrum -> Y [B][nyl][Px][Pz]  (rows jy0 .. jy0 + nyl - 1)
    const long plane = (long)g.Px * g.Pz;
    return fft_pass(g.Py, plane, plane, B * plane, g.Py * plane, g.nyl * plane, g.Py, g.jy0, g.nyl, 1, 1);
}
GB_HD FftPass fft_pass_inv_x(const FftGeom& g, long B) {      // Y [B][nyl][Px][Pz] -> Z [B][nyl][xN][Pz]
    return fft_pass(g.Px, g.Pz, g.Pz, B * g.nyl * g.Pz, (long)g.Px * g.Pz, (long)g.xN * g.Pz, g.Px, 0, g.xN, 1, 1);
}
GB_HD FftPass fft_pass_inv_z(const FftGeom& g, long B) {      // Z [B][nyl][xN][Pz] -> rows of the output
    return fft_pass(g.Pz, 1, 1, B * g.nyl * g.xN, g.Pz, 0, g.Pz, 0, g.zN, 1, 0);
}
