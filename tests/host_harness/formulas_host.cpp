// Host build of csrc/formulas.cuh (TEST HARNESS, compiled by tests/test_formulas_host.py with g++ -ffp-contract=off): the
// scalar functions the device kernels are built from, driven by plain loops that mirror the kernels' indexing
// (create_cov_dense_kernel, cov_function_kernel, cov_tables_kernel, corner_func_kernel, a_sens_kernel).
#include "../../geobo_b200/csrc/formulas.cuh"

#include <vector>

extern "C" {

void host_cov_function(int kernel_id, int cross, const double* D2, long count, double l1, double l2, double* out) {
    for (long i = 0; i < count; ++i) {
        const double d2 = D2[i];
        if (!cross)
            out[i] = kernel_id == GB_KERNEL_EXP ? k_exp_same(d2, l1) : kernel_id == GB_KERNEL_MATERN32 ? k_matern_same(d2, l1) : k_sparse_same(d2, l1);
        else
            out[i] = kernel_id == GB_KERNEL_EXP ? k_exp_cross(d2, l1, l2)
                     : kernel_id == GB_KERNEL_MATERN32 ? k_matern_cross(d2, l1, l2) : k_sparse_cross(d2, l1, l2);
    }
}

// kernels.create_cov for a dense D2 (n x n) -> (3n x 3n), times amp
void host_create_cov(int kernel_id, const double* l, const double* w, double amp, const double* D2, long n, double* out) {
    CovParams P;
    P.kernel_id = kernel_id;
    for (int i = 0; i < 3; ++i) { P.l[i] = l[i]; P.w[i] = w[i]; }
    P.amp = amp;
    const long ld = 3 * n;
    for (long i = 0; i < n; ++i)
        for (long j = 0; j < n; ++j)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) out[(r * n + i) * ld + c * n + j] = cov_value(P, r, c, D2[i * n + j]);
}

// the same matrix straight from the grid spec through the stationary tables over the extended difference lattice
// (cov_tables_kernel + the gather K[(c, j), (r, i)] = tab[c][r][L(i) - L(j) + C0] the fused kernels use)
void host_create_cov_grid(int kernel_id, const double* l, const double* w, double amp, const long* ncube, const double* vox, double* out) {
    CovParams P;
    P.kernel_id = kernel_id;
    for (int i = 0; i < 3; ++i) { P.l[i] = l[i]; P.w[i] = w[i]; }
    P.amp = amp;
    const long xN = ncube[0], yN = ncube[1], zN = ncube[2], N = xN * yN * zN;
    const long EX = 2 * xN - 1, EY = 2 * yN - 1, EZ = 2 * zN - 1, ext = EX * EY * EZ;
    std::vector<double> tab(9 * ext);
    for (int cr = 0; cr < 9; ++cr)
        for (long e = 0; e < ext; ++e) {
            const long ez = e % EZ, t = e / EZ, ex = t % EX, ey = t / EX;
            const double D2 = lattice_d2((int)(ex - (xN - 1)), (int)(ey - (yN - 1)), (int)(ez - (zN - 1)), vox[0], vox[1], vox[2]);
            tab[cr * ext + e] = cov_value(P, cr / 3, cr % 3, D2);
        }
    std::vector<long> L(N);
    for (long iy = 0; iy < yN; ++iy)
        for (long ix = 0; ix < xN; ++ix)
            for (long iz = 0; iz < zN; ++iz) L[(iy * xN + ix) * zN + iz] = (iy * EX + ix) * EZ + iz;
    const long C0 = ((yN - 1) * EX + (xN - 1)) * EZ + (zN - 1), ld = 3 * N;
    for (int c = 0; c < 3; ++c)
        for (long j = 0; j < N; ++j)
            for (int r = 0; r < 3; ++r)
                for (long i = 0; i < N; ++i) out[(c * N + j) * ld + r * N + i] = tab[(c * 3 + r) * ext + (L[i] - L[j] + C0)];
}

void host_corner_func(int kind, const double* x, const double* y, const double* z, long count, const double* B, double* out) {
    for (long i = 0; i < count; ++i)
        out[i] = kind == GB_SENS_GRAV ? grav_corner(x[i], y[i], z[i]) : magn_corner(x[i], y[i], z[i], B[0], B[1], B[2]);
}

// sensormodel.A_sens with the loop structure of a_sens_kernel: corner potentials on the shifted, padded edge lattice
// (edges: [3][yN+1][xN+1][zN+1]), 8-corner alternating difference per voxel, unit scaling
void host_a_sens(int kind, const double* B, const double* loc, long nsens, const double* edges, const long* ncube, double mul, double div,
                 double* out) {
    const long xN = ncube[0], yN = ncube[1], zN = ncube[2], px = xN + 1, pz = zN + 1, plane = px * pz, nedge = (yN + 1) * plane;
    const double *xE = edges, *yE = edges + nedge, *zE = edges + 2 * nedge;
    std::vector<double> eZ(nedge);
    for (long n = 0; n < nsens; ++n) {
        const double lx = loc[3 * n], ly = loc[3 * n + 1], lz = loc[3 * n + 2];
        for (long j = 0; j <= yN; ++j)
            for (long q = 0; q < plane; ++q) {
                const long e = j * plane + q;
                double x0 = __dsub_rn(xE[e], lx), y0 = __dsub_rn(yE[e], ly);
                const double z0 = __dsub_rn(zE[e], lz);
                if (j == 0) { x0 = __dsub_rn(x0, GB_ALONG_WAY); y0 = __dsub_rn(y0, GB_ALONG_WAY); }
                if (j == yN) { x0 = __dadd_rn(x0, GB_ALONG_WAY); y0 = __dadd_rn(y0, GB_ALONG_WAY); }
                eZ[e] = kind == GB_SENS_GRAV ? grav_corner(x0, y0, z0) : magn_corner(x0, y0, z0, B[0], B[1], B[2]);
            }
        for (long iy = 0; iy < yN; ++iy) {
            const double *hi = &eZ[(iy + 1) * plane], *lo = &eZ[iy * plane];
            for (long ix = 0; ix < xN; ++ix)
                for (long iz = 0; iz < zN; ++iz) {
                    const long c00 = ix * pz + iz, c01 = c00 + 1, c10 = c00 + pz, c11 = c10 + 1;
                    const double up = __dadd_rn(__dsub_rn(__dsub_rn(hi[c11], hi[c10]), hi[c01]), hi[c00]);
                    const double dn = __dadd_rn(__dsub_rn(__dsub_rn(lo[c11], lo[c10]), lo[c01]), lo[c00]);
                    const double s = -__dsub_rn(up, dn);
                    out[n * (xN * yN * zN) + (iy * xN + ix) * zN + iz] = __ddiv_rn(__dmul_rn(mul, s), div);
                }
        }
    }
}
}
