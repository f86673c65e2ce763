// Host build of csrc/kron.cuh (TEST HARNESS, compiled by tests/test_kron.py with g++ -ffp-contract=off): the per-thread
// arithmetic of kron_factors_kernel / kron_y_kernel / kron_zx_kernel, driven by loops over blocks, phases and thread ids
// that mirror the launch geometry of kron_apply (kron.cu) -- grid dimensions, row chunks, shared-memory carve-up, the
// barriers between the phases.  The stationary tables are built like cov_tables_kernel does (formulas.cuh).
#include "../../geobo_b200/csrc/formulas.cuh"
#include "../../geobo_b200/csrc/kron.cuh"

#include <cmath>
#include <vector>

// Thread order inside one phase (between two barriers): 0 = ascending, 1 = descending, 2 = odd ids first.  The result must not
// depend on it -- a phase in which one thread reads what another one writes would (tests/test_kron.py runs all three).
static int g_order = 0;
static int tid_at(int i, int n) { return g_order == 0 ? i : g_order == 1 ? n - 1 - i : (i < n / 2 ? 2 * i + 1 : 2 * (i - n / 2)); }
#define FOR_TID(tid, n) for (int tid##_i = 0, tid = tid_at(0, n); tid##_i < (n); ++tid##_i, tid = tid_at(tid##_i < (n) ? tid##_i : 0, n))

extern "C" {

void kron_host_set_thread_order(int order) { g_order = order; }

// out[s][r * r_stride_out + (j - c0)] (+)= sum_i A[s][i] * K_(blk0 + r)[i][j],  rows s < nrows, r = 0..2, j in [c0, c1)
void kron_host_apply(int kernel_id, const double* l, const double* w, double amp, const long* ncube, const double* vox, int blk0,
                     const double* A, long lda, long nrows, long c0, long c1, long chunk_rows, double* out, long ldo, long r_stride_out,
                     int accumulate) {
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
    const KronGeom g = kron_geom(xN, yN, zN, c0, c1);
    // kron_factors_kernel: grid (ceil(FL / 128), 3, 9) x 128 threads
    std::vector<double> kf(9L * 3 * g.FL, std::nan(""));
    for (int b = 0; b < 9; ++b)
        for (int axis = 0; axis < 3; ++axis)
            for (int bx = 0; bx < (g.FL + 127) / 128; ++bx)
                FOR_TID(tid, 128) {
                    const int i = bx * 128 + tid;
                    if (i < g.FL) kf[((long)b * 3 + axis) * g.FL + i] = kron_factor(tab.data() + (long)b * ext + C0, g, axis, i);
                }
    // kron_apply
    const long per_row = 3L * g.nyl * g.XZ;
    const long chunk = chunk_rows < 1 ? 1 : chunk_rows;
    // scratch and "shared memory" start as NaN: device memory is not zero-initialised, a read of an element no thread has written
    // would surface as a NaN in the output
    std::vector<double> T(chunk * per_row, std::nan(""));
    const int qtiles = (int)((g.XZ + KRON_YTHREADS * KRON_YQ - 1) / (KRON_YTHREADS * KRON_YQ));
    const int jgroups = (g.nyl + KRON_JT - 1) / KRON_JT;
    std::vector<double> smem(2L * g.xN * g.zs + 2L * g.FL + 3L * g.FL, std::nan(""));
    for (long s0 = 0; s0 < nrows; s0 += chunk) {
        const long n = nrows - s0 < chunk ? nrows - s0 : chunk;
        const long r_stride = n * g.nyl * g.XZ;
        // kron_y_kernel<<<(qtiles, jgroups, n), KRON_YTHREADS>>>
        for (long bz = 0; bz < n; ++bz)
            for (int by = 0; by < jgroups; ++by)
                for (int bx = 0; bx < qtiles; ++bx) {
                    double* sf = smem.data();
                    FOR_TID(tid, KRON_YTHREADS)
                        for (int i = tid; i < 3 * g.FL; i += KRON_YTHREADS) sf[i] = kf[((long)(blk0 + i / g.FL) * 3 + 0) * g.FL + i % g.FL];
                    // __syncthreads()
                    FOR_TID(tid, KRON_YTHREADS)
                        kron_y_thread(g, A + s0 * lda + bz * lda, sf, bx, by, tid, T.data() + bz * g.nyl * g.XZ, r_stride);
                }
        // kron_zx_kernel<<<(nyl, n, 3), KRON_ZXTHREADS>>>
        for (int r = 0; r < 3; ++r)
            for (long sl = 0; sl < n; ++sl)
                for (int jl = 0; jl < g.nyl; ++jl) {
                    double* pin = smem.data();
                    double* ptmp = pin + (long)g.xN * g.zs;
                    double* fx = ptmp + (long)g.xN * g.zs;
                    double* fz = fx + g.FL;
                    const double* kfb = kf.data() + (long)(blk0 + r) * 3 * g.FL;
                    FOR_TID(tid, KRON_ZXTHREADS)
                        kron_zx_load(g, T.data() + r * r_stride + (sl * g.nyl + jl) * g.XZ, kfb + g.FL, kfb + 2 * g.FL, tid, KRON_ZXTHREADS, pin, fx, fz);
                    // __syncthreads()
                    FOR_TID(tid, KRON_ZXTHREADS) kron_z_phase(g, pin, fz, tid, KRON_ZXTHREADS, ptmp);
                    // __syncthreads()
                    FOR_TID(tid, KRON_ZXTHREADS)
                        kron_x_phase(g, ptmp, fx, g.jy0 + jl, tid, KRON_ZXTHREADS, out + (s0 + sl) * ldo + r * r_stride_out, accumulate);
                }
    }
}

}  // extern "C"
