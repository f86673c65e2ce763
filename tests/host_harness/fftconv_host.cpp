// Host build of csrc/fftconv.cuh (TEST HARNESS, compiled by tests/test_fftconv.py with g++ -ffp-contract=off): the per-thread
// arithmetic of fft_wrap_kernel / fft_pass_kernel<0|1|2> / fft_real_kernel driven by loops over blocks, phases (load, one
// phase per butterfly stage, store -- the barriers of the kernel) and thread ids, with the pass sequence, chunking and scratch
// carve-up of fft_build_spectra / fft_apply (fftconv.cu).  Stationary tables as cov_tables_kernel builds them (formulas.cuh).
#include "../../geobo_b200/csrc/formulas.cuh"
#include "../../geobo_b200/csrc/fftconv.cuh"

#include <cmath>
#include <vector>

// Thread order inside one phase (between two barriers): 0 = ascending, 1 = descending, 2 = odd ids first.  The result must not
// depend on it -- a phase in which one thread reads what another one writes would (tests/test_fftconv.py runs all three).
static int g_order = 0;
static int tid_at(int i, int n) { return g_order == 0 ? i : g_order == 1 ? n - 1 - i : (i < n / 2 ? 2 * i + 1 : 2 * (i - n / 2)); }
#define FOR_TID(tid, n) for (int tid##_i = 0, tid = tid_at(0, n); tid##_i < (n); ++tid##_i, tid = tid_at(tid##_i < (n) ? tid##_i : 0, n))

namespace {

// fft_pass_kernel<MODE><<<ceil(nlines / FFT_LPB), FFT_THREADS>>>
void run_pass(int mode, const FftGeom& g, const FftPass& q, const cplx* in, const double* mul, cplx* out, const cplx* tw, const double* A, long lda,
              long nrows, double* rows_out, long ldo, int accumulate) {
    // "shared memory" and the scratch lattices start as NaN: device memory is not zero-initialised, a read of an element no thread
    // has written would surface as a NaN in the output
    cplx poison;
    poison.re = poison.im = std::nan("");
    std::vector<cplx> sm((size_t)FFT_LPB * q.P + q.P / 2 + 1, poison);
    cplx* tws = sm.data() + FFT_LPB * q.P;
    const long blocks = (q.nlines + FFT_LPB - 1) / FFT_LPB;
    for (long bx = 0; bx < blocks; ++bx) {
        FOR_TID(tid, FFT_THREADS)
            for (int i = tid; i < q.P / 2; i += FFT_THREADS) tws[i] = tw[i];
        const long line0 = bx * FFT_LPB;
        FOR_TID(tid, FFT_THREADS) {
            if (mode == 1) fft_load_rows(g, q, A, lda, nrows, line0, tid, FFT_THREADS, sm.data());
            else fft_load(q, in, mul, line0, tid, FFT_THREADS, sm.data());
        }
        for (int s = 1; s <= q.logP; ++s)      // __syncthreads() between the phases
            FOR_TID(tid, FFT_THREADS) fft_stage(q, s, tws, tid, FFT_THREADS, sm.data());
        FOR_TID(tid, FFT_THREADS) {
            if (mode == 2) fft_store_rows(g, q, rows_out, ldo, nrows, accumulate, line0, tid, FFT_THREADS, sm.data());
            else fft_store(q, out, line0, tid, FFT_THREADS, sm.data());
        }
    }
}

}  // namespace

extern "C" {

void fftconv_host_set_thread_order(int order) { g_order = order; }

// out[s][r * r_stride_out + (j - c0)] (+)= sum_i A[s][i] * K_(blk0 + r)[i][j],  rows s < nrows, r = 0..2, j in [c0, c1)
// B: complex row pairs per chunk; padded_out: (Py, Px, Pz)
void fftconv_host_apply(int kernel_id, const double* l, const double* w, double amp, const long* ncube, const double* vox, int blk0,
                        const double* A, long lda, long nrows, long c0, long c1, long B, double* out, long ldo, long r_stride_out, int accumulate,
                        int* padded_out) {
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
    const FftGeom g = fft_geom(xN, yN, zN, c0, c1);
    padded_out[0] = g.Py; padded_out[1] = g.Px; padded_out[2] = g.Pz;
    // fft_build_twiddles
    std::vector<cplx> tw(3 * (FFT_MAXP / 2));
    const int Pa[3] = {g.Py, g.Px, g.Pz};
    for (int a = 0; a < 3; ++a)
        for (int k = 0; k < Pa[a] / 2; ++k) {
            tw[a * (FFT_MAXP / 2) + k].re = cos(-2.0 * GB_PI * k / Pa[a]);
            tw[a * (FFT_MAXP / 2) + k].im = sin(-2.0 * GB_PI * k / Pa[a]);
        }
    if (B < 1) B = 1;
    cplx poison;
    poison.re = poison.im = std::nan("");
    std::vector<cplx> scratch((size_t)B * (2 * g.P3 + (long)g.nyl * g.xN * g.Pz), poison);
    cplx* X = scratch.data();
    cplx* Y = X + B * g.P3;
    cplx* Z = Y + B * g.P3;
    const cplx *twy = tw.data(), *twx = tw.data() + FFT_MAXP / 2, *twz = tw.data() + 2 * (FFT_MAXP / 2);
    // fft_build_spectra (only the three blocks this call uses)
    std::vector<double> W(9 * g.P3, 0.0);
    {
        const long plane = (long)g.Px * g.Pz;
        const FftPass pz = fft_pass(g.Pz, 1, 1, (long)g.Py * g.Px, g.Pz, g.Pz, g.Pz, 0, g.Pz, 0, 0);
        const FftPass px = fft_pass(g.Px, g.Pz, g.Pz, (long)g.Py * g.Pz, plane, plane, g.Px, 0, g.Px, 0, 1);
        const FftPass py = fft_pass(g.Py, plane, plane, plane, g.Py * plane, g.Py * plane, g.Py, 0, g.Py, 0, 1);
        for (int b = blk0; b < blk0 + 3; ++b) {
            for (long e = 0; e < g.P3; ++e) X[e] = fft_wrapped_tap(g, tab.data() + (long)b * ext + C0, e);
            run_pass(0, g, pz, X, nullptr, Y, twz, nullptr, 0, 0, nullptr, 0, 0);
            run_pass(0, g, px, Y, nullptr, X, twx, nullptr, 0, 0, nullptr, 0, 0);
            run_pass(0, g, py, X, nullptr, Y, twy, nullptr, 0, 0, nullptr, 0, 0);
            for (long e = 0; e < g.P3; ++e) W[(long)b * g.P3 + e] = Y[e].re;
        }
    }
    // fft_apply
    for (long s0 = 0; s0 < nrows; s0 += 2 * B) {
        const long n = nrows - s0 < 2 * B ? nrows - s0 : 2 * B, nb = (n + 1) / 2;
        run_pass(1, g, fft_pass_fwd_z(g, nb), nullptr, nullptr, X, twz, A + s0 * lda, lda, n, nullptr, 0, 0);
        run_pass(0, g, fft_pass_fwd_x(g, nb), X, nullptr, Y, twx, nullptr, 0, 0, nullptr, 0, 0);
        run_pass(0, g, fft_pass_fwd_y(g, nb), Y, nullptr, X, twy, nullptr, 0, 0, nullptr, 0, 0);
        for (int r = 0; r < 3; ++r) {
            run_pass(0, g, fft_pass_inv_y(g, nb), X, W.data() + (long)(blk0 + r) * g.P3, Y, twy, nullptr, 0, 0, nullptr, 0, 0);
            run_pass(0, g, fft_pass_inv_x(g, nb), Y, nullptr, Z, twx, nullptr, 0, 0, nullptr, 0, 0);
            run_pass(2, g, fft_pass_inv_z(g, nb), Z, nullptr, nullptr, twz, nullptr, 0, n, out + s0 * ldo + r * r_stride_out, ldo, accumulate);
        }
    }
}

}  // extern "C"
