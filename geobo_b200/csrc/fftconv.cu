// Block-Toeplitz (FFT) products with the stationary covariance blocks (opt-in,
// gb_hyper.structure = GB_STRUCTURE_FFT, any kernel family): Pt = A3 . K and z = K . w by circulant embedding -- zero-padded
// 3-D transforms of the rows (two real rows per complex transform), multiplication with the real spectrum of the wrapped
// covariance table, inverse transforms restricted to the rank's voxel columns.  Per-thread arithmetic: fftconv.cuh (also
// compiled for the host by the CPU tests).
//
//   fft_twiddle_kernel   : exp(-2 pi i k / P) tables of the three axes (sincospi, fp64)
//   fft_wrap_kernel      : stationary table of one block -> wrapped complex lattice [Py][Px][Pz]
//   fft_pass_kernel<M>   : one batched radix-2 pass (FFT_LPB lines per block in shared memory); M = 0 generic (optional
//                          multiplication with the spectrum on load), 1 loads from rows of A, 2 stores into rows of Pt / z
//   fft_real_kernel      : real part of the transformed lattice = spectrum of the block
// Row pairs are processed in chunks whose three scratch lattices stay L2 resident (GEOBO_B200_FFT_SCRATCH_MB, default 96).
#include "common.cuh"
#include "fftconv.cuh"

__global__ void fft_twiddle_kernel(int P, cplx* __restrict__ tw) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < P / 2) {
        double sn, cs;
        sincospi(-2.0 * (double)k / (double)P, &sn, &cs);
        tw[k].re = cs;
        tw[k].im = sn;
    }
}

__global__ void fft_wrap_kernel(FftGeom g, const double* __restrict__ tab0, cplx* __restrict__ X) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < g.P3) X[e] = fft_wrapped_tap(g, tab0, e);
}

__global__ void fft_real_kernel(long n, const cplx* __restrict__ X, double* __restrict__ W) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) W[e] = X[e].re;
}

template <int MODE>
__global__ void __launch_bounds__(FFT_THREADS) fft_pass_kernel(FftGeom g, FftPass q, const cplx* __restrict__ in, const double* __restrict__ mul,
                                                               cplx* __restrict__ out, const cplx* __restrict__ tw, const double* __restrict__ A,
                                                               long lda, long nrows, double* __restrict__ rows_out, long ldo, int accumulate) {
    extern __shared__ __align__(16) unsigned char fft_smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(fft_smem_raw);
    cplx* tws = sm + FFT_LPB * q.P;
    for (int i = threadIdx.x; i < q.P / 2; i += FFT_THREADS) tws[i] = tw[i];
    const long line0 = (long)blockIdx.x * FFT_LPB;
    if (MODE == 1) fft_load_rows(g, q, A, lda, nrows, line0, (int)threadIdx.x, FFT_THREADS, sm);
    else fft_load(q, in, mul, line0, (int)threadIdx.x, FFT_THREADS, sm);
    __syncthreads();
    for (int s = 1; s <= q.logP; ++s) {
        fft_stage(q, s, tws, (int)threadIdx.x, FFT_THREADS, sm);
        __syncthreads();
    }
    if (MODE == 2) fft_store_rows(g, q, rows_out, ldo, nrows, accumulate, line0, (int)threadIdx.x, FFT_THREADS, sm);
    else fft_store(q, out, line0, (int)threadIdx.x, FFT_THREADS, sm);
}

static size_t fft_smem(int P) { return ((size_t)FFT_LPB * P + P / 2 + 1) * sizeof(cplx); }

template <int MODE>
static cudaError_t launch_pass(const FftGeom& g, const FftPass& q, const cplx* in, const double* mul, cplx* out, const cplx* tw, const double* A,
                               long lda, long nrows, double* rows_out, long ldo, int accumulate, cudaStream_t s, long* nlaunch) {
    const size_t smem = fft_smem(q.P);
    cudaError_t e = cudaFuncSetAttribute(fft_pass_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)((q.nlines + FFT_LPB - 1) / FFT_LPB);
    fft_pass_kernel<MODE><<<blocks, FFT_THREADS, smem, s>>>(g, q, in, mul, out, tw, A, lda, nrows, rows_out, ldo, accumulate);
    if (nlaunch) *nlaunch += 1;
    return cudaGetLastError();
}

int fft_supported(const FftGeom& g, char* why, size_t len) {
    if (g.Px > FFT_MAXP || g.Py > FFT_MAXP || g.Pz > FFT_MAXP) {
        snprintf(why, len, "structure = fft: padded lengths (%d, %d, %d) exceed %d", g.Px, g.Py, g.Pz, FFT_MAXP);
        return 0;
    }
    return 1;
}

// complex row pairs per chunk so that X, Y, Z fit the scratch budget
long fft_chunk_pairs(const FftGeom& g, long rows) {
    long mb = 96;
    if (const char* e = getenv("GEOBO_B200_FFT_SCRATCH_MB")) mb = atol(e) > 0 ? atol(e) : mb;
    const long per_pair = (2 * g.P3 + (long)g.nyl * g.xN * g.Pz) * (long)sizeof(cplx);
    long B = (mb << 20) / per_pair;
    if (B < 1) B = 1;
    if (B > (rows + 1) / 2) B = (rows + 1) / 2;
    return B;
}
long fft_scratch_cplx(const FftGeom& g, long B) { return B * (2 * g.P3 + (long)g.nyl * g.xN * g.Pz); }

cudaError_t fft_build_twiddles(const FftGeom& g, cplx* tw /*[3][FFT_MAXP / 2]: y, x, z*/, cudaStream_t s) {
    const int P[3] = {g.Py, g.Px, g.Pz};
    for (int a = 0; a < 3; ++a)
        if (P[a] > 1) fft_twiddle_kernel<<<(P[a] / 2 + 127) / 128, 128, 0, s>>>(P[a], tw + a * (FFT_MAXP / 2));
    return cudaGetLastError();
}

// W[b] = real spectrum of the wrapped table of block b, b = 0..8; X, Y: two scratch lattices of P3 complex values (ping-pong)
cudaError_t fft_build_spectra(const FftGeom& g, const double* tables, long ext, long C0, const cplx* tw, cplx* X, cplx* Y, double* W,
                              cudaStream_t s, long* nlaunch) {
    const long plane = (long)g.Px * g.Pz;
    const FftPass pz = fft_pass(g.Pz, 1, 1, (long)g.Py * g.Px, g.Pz, g.Pz, g.Pz, 0, g.Pz, 0, 0);
    const FftPass px = fft_pass(g.Px, g.Pz, g.Pz, (long)g.Py * g.Pz, plane, plane, g.Px, 0, g.Px, 0, 1);
    const FftPass py = fft_pass(g.Py, plane, plane, plane, g.Py * plane, g.Py * plane, g.Py, 0, g.Py, 0, 1);
    const unsigned eb = (unsigned)((g.P3 + 255) / 256);
    for (int b = 0; b < 9; ++b) {
        fft_wrap_kernel<<<eb, 256, 0, s>>>(g, tables + (long)b * ext + C0, X);
        cudaError_t e;
        if ((e = launch_pass<0>(g, pz, X, nullptr, Y, tw + 2 * (FFT_MAXP / 2), nullptr, 0, 0, nullptr, 0, 0, s, nlaunch)) != cudaSuccess) return e;
        if ((e = launch_pass<0>(g, px, Y, nullptr, X, tw + 1 * (FFT_MAXP / 2), nullptr, 0, 0, nullptr, 0, 0, s, nlaunch)) != cudaSuccess) return e;
        if ((e = launch_pass<0>(g, py, X, nullptr, Y, tw + 0 * (FFT_MAXP / 2), nullptr, 0, 0, nullptr, 0, 0, s, nlaunch)) != cudaSuccess) return e;
        fft_real_kernel<<<eb, 256, 0, s>>>(g.P3, Y, W + (long)b * g.P3);
        if (nlaunch) *nlaunch += 2;
    }
    return cudaGetLastError();
}

// out[s][r * r_stride_out + (j - c0)] (+)= sum_i A[s][i] * K_(blk0 + r)[i][j]   for rows s < nrows, r = 0..2, j in [c0, c1)
// scratch: B * (2 P3 + nyl xN Pz) complex values (X, Y, Z)
cudaError_t fft_apply(const FftGeom& g, const double* W, const cplx* tw, int blk0, const double* A, long lda, long nrows, cplx* scratch, long B,
                      double* out, long ldo, long r_stride_out, int accumulate, cudaStream_t s, long* nlaunch, int nr) {
    if (B < 1) return cudaErrorInvalidValue;
    cplx* X = scratch;
    cplx* Y = X + B * g.P3;
    cplx* Z = Y + B * g.P3;
    const cplx *twy = tw, *twx = tw + FFT_MAXP / 2, *twz = tw + 2 * (FFT_MAXP / 2);
    cudaError_t e;
    for (long s0 = 0; s0 < nrows; s0 += 2 * B) {
        const long n = nrows - s0 < 2 * B ? nrows - s0 : 2 * B, nb = (n + 1) / 2;
        if ((e = launch_pass<1>(g, fft_pass_fwd_z(g, nb), nullptr, nullptr, X, twz, A + s0 * lda, lda, n, nullptr, 0, 0, s, nlaunch)) != cudaSuccess) return e;
        if ((e = launch_pass<0>(g, fft_pass_fwd_x(g, nb), X, nullptr, Y, twx, nullptr, 0, 0, nullptr, 0, 0, s, nlaunch)) != cudaSuccess) return e;
        if ((e = launch_pass<0>(g, fft_pass_fwd_y(g, nb), Y, nullptr, X, twy, nullptr, 0, 0, nullptr, 0, 0, s, nlaunch)) != cudaSuccess) return e;
        for (int r = 0; r < nr; ++r) {
            if ((e = launch_pass<0>(g, fft_pass_inv_y(g, nb), X, W + (long)(blk0 + r) * g.P3, Y, twy, nullptr, 0, 0, nullptr, 0, 0, s, nlaunch)) != cudaSuccess) return e;
            if ((e = launch_pass<0>(g, fft_pass_inv_x(g, nb), Y, nullptr, Z, twx, nullptr, 0, 0, nullptr, 0, 0, s, nlaunch)) != cudaSuccess) return e;
            if ((e = launch_pass<2>(g, fft_pass_inv_z(g, nb), Z, nullptr, nullptr, twz, nullptr, 0, n, out + s0 * ldo + r * r_stride_out, ldo, accumulate, s,
                                    nlaunch)) != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}
