// Compact-support form of the 'sparse' covariance blocks -- SURVEY.md section 8(f) row 3 ("tile culling for the compact
// kernel", done per tap instead of per tile).
//
// The Melkumyan kernels (geobo/kernels.py:101-138; the reference's default kernelfunc) vanish for d >= gamma (same
// property) and d > (l1 + l2) / 2 (cross), so on the voxel grid a product with a block K_cr is a 3-D stencil over the
// offsets |dy| <= ry, |dx| <= rx, |dz| <= rz that can lie inside the support:
//     (A K_cr)[s, (jy,jx,jz)] = sum_{dy,dx,dz} tab_cr[dy,dx,dz] * A[s, (jy-dy, jx-dx, jz-dz)]
// with the taps read from the stationary covariance tables (cov.cu) -- (2ry+1)(2rx+1)(2rz+1) multiply-adds per output
// instead of N.  Taps inside the window but outside the support are exact zeros in the table and are skipped.
//
// Per-thread arithmetic of stencil.cu, written against an item id so that the CPU suite can compile it with g++
// (tests/host_harness/stencil_host.cpp).  OPT-IN (gb_hyper.structure = GB_STRUCTURE_COMPACT); the dense contraction stays
// the default.
#pragma once
#ifndef GB_HD
#if defined(__CUDACC__)
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif
#endif

constexpr int STENCIL_W = 4;          // consecutive z outputs per thread (sliding input window)
constexpr int STENCIL_THREADS = 256;

struct StencilGeom {
    int xN, yN, zN;
    long XZ;           // xN * zN
    int jy0, nyl;      // y-rows touched by the rank's voxel-column shard [c0, c1)
    long c0, c1;
    int ry, rx, rz;    // half widths of the tap window
    long sy, sx;       // strides of the y and x offsets in the extended-lattice tables: (2xN-1)(2zN-1), 2zN-1
    long ext;          // doubles per table
};

// Half width of the window on an axis with voxel size `vox` and `n` voxels: offsets k with k * vox inside the largest
// support radius; radius = max length scale * 1.001 covers the cross blocks ((l1 + l2) / 2 with the 1e-3 bump of
// kernels.py:125-126).  Over-estimating is harmless (the extra taps are zeros of the table).
GB_HD int stencil_half_width(double radius, double vox, long n) {
    double q = radius / vox + 1e-9;
    if (!(q >= 0.0)) q = 0.0;
    long k = q >= (double)(n - 1) ? n - 1 : (long)q;
    return (int)k;
}

GB_HD StencilGeom stencil_geom(long xN, long yN, long zN, long c0, long c1, const double vox[3], const double gp_length[3]) {
    StencilGeom g;
    g.xN = (int)xN; g.yN = (int)yN; g.zN = (int)zN;
    g.XZ = xN * zN;
    g.c0 = c0; g.c1 = c1;
    g.jy0 = (int)(c0 / g.XZ);
    g.nyl = (int)((c1 - 1) / g.XZ) - g.jy0 + 1;
    double r = gp_length[0] > gp_length[1] ? gp_length[0] : gp_length[1];
    if (gp_length[2] > r) r = gp_length[2];
    r *= 1.001;
    g.rx = stencil_half_width(r, vox[0], xN);
    g.ry = stencil_half_width(r, vox[1], yN);
    g.rz = stencil_half_width(r, vox[2], zN);
    g.sy = (2 * xN - 1) * (2 * zN - 1);
    g.sx = 2 * zN - 1;
    g.ext = (2 * xN - 1) * (2 * yN - 1) * (2 * zN - 1);
    return g;
}

// One work item = STENCIL_W consecutive z outputs of voxel column (jy, jx) for the three property blocks r of one data
// block.  tab0: zero offset of the first of the three tables (tables + blk0 * ext + C0); row: output row (already offset to
// the data row), block r at row + r * r_stride_out, voxel column j at [j - c0].
GB_HD void stencil_item(const StencilGeom& g, const double* Arow, const double* tab0, int jy, long item, double* row, long r_stride_out,
                        int accumulate, int nr = 3) {
    const int nstrip = (g.zN + STENCIL_W - 1) / STENCIL_W;
    if (item >= (long)g.xN * nstrip) return;
    const int jx = (int)(item / nstrip), z0 = (int)(item % nstrip) * STENCIL_W;
    double acc[3][STENCIL_W];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int t = 0; t < STENCIL_W; ++t) acc[r][t] = 0.0;
    for (int dy = -g.ry; dy <= g.ry; ++dy) {
        const int iy = jy - dy;
        if (iy < 0 || iy >= g.yN) continue;
        for (int dx = -g.rx; dx <= g.rx; ++dx) {
            const int ix = jx - dx;
            if (ix < 0 || ix >= g.xN) continue;
            const double* in = Arow + ((long)iy * g.xN + ix) * g.zN;
            const double* tap = tab0 + dy * g.sy + dx * g.sx;           // tap[dz] (+ r * ext for block r)
            // out[z0 + t] += tap[dz] * in[z0 + t - dz];  u = -dz runs upwards, the input window slides by one per step
            double w[STENCIL_W];
#pragma unroll
            for (int t = 0; t < STENCIL_W; ++t) {
                const int z = z0 + t - g.rz;
                w[t] = (z >= 0 && z < g.zN) ? in[z] : 0.0;
            }
            for (int u = -g.rz; u <= g.rz; ++u) {
                const double f0 = tap[-u], f1 = tap[g.ext - u], f2 = tap[2 * g.ext - u];
                if (f0 != 0.0 || f1 != 0.0 || f2 != 0.0) {
#pragma unroll
                    for (int t = 0; t < STENCIL_W; ++t) {
                        acc[0][t] += f0 * w[t];
                        acc[1][t] += f1 * w[t];
                        acc[2][t] += f2 * w[t];
                    }
                }
#pragma unroll
                for (int t = 0; t < STENCIL_W - 1; ++t) w[t] = w[t + 1];
                const int z = z0 + STENCIL_W + u;                        // = z0 + (STENCIL_W - 1) + (u + 1)
                w[STENCIL_W - 1] = (z >= 0 && z < g.zN) ? in[z] : 0.0;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < STENCIL_W; ++t) {
        if (z0 + t >= g.zN) continue;
        const long j = (long)jy * g.XZ + (long)jx * g.zN + z0 + t;
        if (j < g.c0 || j >= g.c1) continue;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (r >= nr) continue;                       // without drill data only the first nr = 2 property blocks are stored
            double* o = row + r * r_stride_out + (j - g.c0);
            *o = accumulate ? *o + acc[r][t] : acc[r][t];
        }
    }
}
