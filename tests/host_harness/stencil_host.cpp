// Host build of csrc/stencil.cuh (TEST HARNESS, compiled by tests/test_compact.py with g++ -ffp-contract=off): the per-thread
// arithmetic of stencil_kernel driven by loops over the launch geometry of stencil_apply (stencil.cu); the stationary tables
// are built like cov_tables_kernel does (formulas.cuh).
#include "../../geobo_b200/csrc/formulas.cuh"
#include "../../geobo_b200/csrc/stencil.cuh"

#include <vector>

extern "C" {

// out[s][r * r_stride_out + (j - c0)] (+)= sum_i A[s][i] * K_(blk0 + r)[i][j],  rows s < nrows, r = 0..2, j in [c0, c1)
// window_out: the (ry, rx, rz) the library would use for these length scales
void stencil_host_apply(int kernel_id, const double* l, const double* w, double amp, const long* ncube, const double* vox, int blk0,
                        const double* A, long lda, long nrows, long c0, long c1, double* out, long ldo, long r_stride_out, int accumulate,
                        int* window_out) {
    CovParams P;
    P.kernel_id = kernel_id;
    for (int i = 0; i < 3; ++i) { P.l[i] = l[i]; P.w[i] = w[i]; }
    P.amp = amp;
    const long xN = ncube[0], yN = ncube[1], zN = ncube[2];
    const long EX = 2 * xN - 1, EY = 2 * yN - 1, EZ = 2 * zN - 1, ext = EX * EY * EZ;
    const long C0 = ((yN - 1) * EX + (xN - 1)) * EZ + (zN - 1);
    std::vector<double> tab(9 * ext);
    for (int cr = 0; cr < 9; ++cr)
        for (long e = 0; e < ext; ++e) {
            const long ez = e % EZ, t = e / EZ, ex = t % EX, ey = t / EX;
            tab[cr * ext + e] = cov_value(P, cr / 3, cr % 3, lattice_d2((int)(ex - (xN - 1)), (int)(ey - (yN - 1)), (int)(ez - (zN - 1)), vox[0], vox[1], vox[2]));
        }
    const StencilGeom g = stencil_geom(xN, yN, zN, c0, c1, vox, l);
    window_out[0] = g.ry; window_out[1] = g.rx; window_out[2] = g.rz;
    // stencil_kernel<<<(tiles, nyl, rows), STENCIL_THREADS>>>
    const int nstrip = (g.zN + STENCIL_W - 1) / STENCIL_W;
    const long tiles = ((long)g.xN * nstrip + STENCIL_THREADS - 1) / STENCIL_THREADS;
    for (long bz = 0; bz < nrows; ++bz)
        for (int by = 0; by < g.nyl; ++by)
            for (long bx = 0; bx < tiles; ++bx)
                for (int tid = 0; tid < STENCIL_THREADS; ++tid)
                    stencil_item(g, A + bz * lda, tab.data() + (long)blk0 * ext + C0, g.jy0 + by, bx * STENCIL_THREADS + tid, out + bz * ldo,
                                 r_stride_out, accumulate);
}

}  // extern "C"
