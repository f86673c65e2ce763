// C ABI of libgeobo_b200 (see include/geobo_b200.h): context, the geobo.kernels / geobo.sensormodel
// compatibility entry points, and the device-resident joint inversion (Inversion.cubing / predict3).
#include <math.h>
#include <stdarg.h>

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kron.cuh"
#include "stencil.cuh"
#include "fftconv.cuh"
#include "comm.h"

std::string g_gb_create_error;

// ====================================================================================== context
extern "C" int gb_version(void) { return 100; }

extern "C" int gb_ctx_create(int device, gb_ctx** out) {
    if (!out) return gb_fail(nullptr, GB_ERR_ARG, "gb_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return gb_fail(nullptr, GB_ERR_CUDA, "no CUDA device available (%s); geobo_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0) GB_CUDA(nullptr, cudaGetDevice(&device));
    if (device >= ndev) return gb_fail(nullptr, GB_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    GB_CUDA(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    GB_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return gb_fail(nullptr, GB_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                       prop.major, prop.minor);
    gb_ctx* ctx = new gb_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = gemm::init();
    if (e != cudaSuccess) {
        int rc = gb_fail(nullptr, GB_ERR_CUDA, "context init: %s", cudaGetErrorString(e));
        delete ctx;
        return rc;
    }
    *out = ctx;
    return GB_OK;
}

// ---------------------------------------------------------------------------------- cached allocators
static void release_cache(gb_ctx* ctx) {
    for (auto& kv : ctx->dev_cache) { cudaFree(kv.second); ctx->dev_sizes.erase(kv.second); }
    ctx->dev_cache.clear();
    for (auto& kv : ctx->host_cache) { cudaFreeHost(kv.second); ctx->host_sizes.erase(kv.second); }
    ctx->host_cache.clear();
}

cudaError_t gb_dev_malloc(gb_ctx* ctx, void** p, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    auto it = ctx->dev_cache.find(bytes);
    if (it != ctx->dev_cache.end()) {
        *p = it->second;
        ctx->dev_cache.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {      // make room: drop everything cached, retry once
        cudaGetLastError();
        release_cache(ctx);
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) ctx->dev_sizes[*p] = bytes;
    return e;
}

void gb_dev_free(gb_ctx* ctx, void* p) {
    if (!p) return;
    auto it = ctx->dev_sizes.find(p);
    if (it == ctx->dev_sizes.end()) { cudaFree(p); return; }
    ctx->dev_cache.insert({it->second, p});
}

cudaError_t gb_host_malloc(gb_ctx* ctx, void** p, size_t bytes) {
    bytes = (bytes + 4095) & ~(size_t)4095;
    auto it = ctx->host_cache.find(bytes);
    if (it != ctx->host_cache.end()) {
        *p = it->second;
        ctx->host_cache.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMallocHost(p, bytes);
    if (e == cudaSuccess) ctx->host_sizes[*p] = bytes;
    return e;
}

void gb_host_free(gb_ctx* ctx, void* p) {
    if (!p) return;
    auto it = ctx->host_sizes.find(p);
    if (it == ctx->host_sizes.end()) { cudaFreeHost(p); return; }
    ctx->host_cache.insert({it->second, p});
}

extern "C" int gb_ctx_release_cache(gb_ctx* ctx) {
    if (!ctx) return GB_ERR_ARG;
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    release_cache(ctx);
    return GB_OK;
}

extern "C" int gb_ctx_destroy(gb_ctx* ctx) {
    if (!ctx) return GB_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    release_cache(ctx);
    comm_destroy(ctx);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return GB_OK;
}

extern "C" const char* gb_last_error(const gb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_gb_create_error.c_str(); }

extern "C" int gb_device_info(gb_ctx* ctx, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor,
                              uint64_t* free_bytes, uint64_t* total_bytes) {
    if (!ctx) return GB_ERR_ARG;
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    GB_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    if (name && name_len > 0) snprintf(name, name_len, "%s", prop.name);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    size_t f = 0, t = 0;
    GB_CUDA(ctx, cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return GB_OK;
}

static int fill_cov_params(gb_ctx* ctx, CovParams& cp, int kernel_id, const double l[3], const double w[3], double amp) {
    if (kernel_id != GB_KERNEL_SPARSE && kernel_id != GB_KERNEL_EXP && kernel_id != GB_KERNEL_MATERN32)
        return gb_fail(ctx, GB_ERR_ARG, "unknown kernel id %d", kernel_id);
    cp.kernel_id = kernel_id;
    for (int i = 0; i < 3; ++i) { cp.l[i] = l[i]; cp.w[i] = w[i]; }
    cp.amp = amp;
    return GB_OK;
}

// ====================================================================================== geobo/kernels.py
extern "C" int gb_grid_points(gb_ctx* ctx, const int64_t lpix[3], const double pixscale[3], double* out) {
    if (!ctx || !lpix || !pixscale || !out) return gb_fail(ctx, GB_ERR_ARG, "gb_grid_points: null argument");
    const int64_t N = lpix[0] * lpix[1] * lpix[2];
    if (lpix[0] < 1 || lpix[1] < 1 || lpix[2] < 1 || N > (1LL << 31)) return gb_fail(ctx, GB_ERR_ARG, "gb_grid_points: bad grid");
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> d;
    GB_CUDA(ctx, d.alloc(3 * N));
    GB_CUDA(ctx, launch_grid_points(lpix, pixscale, d.p, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(out, d.p, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

extern "C" int gb_sqdist(gb_ctx* ctx, const double* points, int64_t n, int dim, double* out) {
    if (!ctx || !points || !out || n < 1 || dim < 1) return gb_fail(ctx, GB_ERR_ARG, "gb_sqdist: bad argument");
    if (n > 65535) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "gb_sqdist: n=%lld > 65535 (dense n x n output)", (long long)n);
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> p, d;
    GB_CUDA(ctx, p.alloc(n * dim));
    GB_CUDA(ctx, d.alloc(n * n));
    GB_CUDA(ctx, cudaMemcpyAsync(p.p, points, n * dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, launch_sqdist(p.p, n, dim, d.p, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(out, d.p, n * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

extern "C" int gb_create_cov(gb_ctx* ctx, const double* D2, int64_t n, const double gpl[3], const double w[3], int kernel_id,
                             double* out) {
    if (!ctx || !D2 || !gpl || !w || !out || n < 1) return gb_fail(ctx, GB_ERR_ARG, "gb_create_cov: bad argument");
    if (n > 65535) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "gb_create_cov: n=%lld > 65535 (dense 3n x 3n output)", (long long)n);
    CovParams cp;
    GB_TRY(fill_cov_params(ctx, cp, kernel_id, gpl, w, 1.0));
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> d2, o;
    GB_CUDA(ctx, d2.alloc(n * n));
    GB_CUDA(ctx, o.alloc(9 * n * n));
    GB_CUDA(ctx, cudaMemcpyAsync(d2.p, D2, n * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, launch_create_cov_dense(cp, d2.p, n, o.p, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(out, o.p, 9 * n * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

extern "C" int gb_create_cov_grid(gb_ctx* ctx, const int64_t ncube[3], const double vox[3], const double gpl[3],
                                  const double w[3], double amp, int kernel_id, double* out, float* ms) {
    if (!ctx || !ncube || !vox || !gpl || !w) return gb_fail(ctx, GB_ERR_ARG, "gb_create_cov_grid: null argument");
    const int64_t N = ncube[0] * ncube[1] * ncube[2];
    if (N < 1 || 3 * N > (1LL << 31) - 1) return gb_fail(ctx, GB_ERR_ARG, "gb_create_cov_grid: bad grid");
    CovParams cp;
    GB_TRY(fill_cov_params(ctx, cp, kernel_id, gpl, w, amp));
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> o, tab;
    DevBuf<int> lat;
    GB_CUDA(ctx, o.alloc(9 * N * N));
    const int64_t ext = (2 * ncube[0] - 1) * (2 * ncube[1] - 1) * (2 * ncube[2] - 1);
    const bool direct = getenv("GEOBO_B200_ASSEMBLY_DIRECT") != nullptr;     // per-element evaluation (kept for comparison)
    if (!direct) {
        if (ext >= (1LL << 31)) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "cube too large for 32-bit lattice ids");
        GB_CUDA(ctx, tab.alloc(9 * ext));
        GB_CUDA(ctx, lat.alloc(N));
    }
    cudaEvent_t e0, e1;
    GB_CUDA(ctx, cudaEventCreate(&e0));
    GB_CUDA(ctx, cudaEventCreate(&e1));
    GB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    GB_CUDA(ctx, launch_create_cov_grid(cp, ncube, vox, o.p, ctx->stream, direct ? nullptr : tab.p, direct ? nullptr : lat.p));
    GB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    if (out) GB_CUDA(ctx, cudaMemcpyAsync(out, o.p, 9 * N * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float t = 0.f;
    GB_CUDA(ctx, cudaEventElapsedTime(&t, e0, e1));
    if (ms) *ms = t;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return GB_OK;
}

extern "C" int gb_cov_function(gb_ctx* ctx, int kernel_id, int cross, const double* D2, int64_t count, double l1, double l2,
                               double* out) {
    if (!ctx || !D2 || !out || count < 0) return gb_fail(ctx, GB_ERR_ARG, "gb_cov_function: bad argument");
    if (kernel_id != GB_KERNEL_SPARSE && kernel_id != GB_KERNEL_EXP && kernel_id != GB_KERNEL_MATERN32)
        return gb_fail(ctx, GB_ERR_ARG, "unknown kernel id %d", kernel_id);
    if (count == 0) return GB_OK;
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> d, o;
    GB_CUDA(ctx, d.alloc(count));
    GB_CUDA(ctx, o.alloc(count));
    GB_CUDA(ctx, cudaMemcpyAsync(d.p, D2, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, launch_cov_function(kernel_id, cross, d.p, count, l1, l2, o.p, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(out, o.p, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

// ====================================================================================== geobo/sensormodel.py
extern "C" int gb_corner_func(gb_ctx* ctx, int kind, const double* x, const double* y, const double* z, int64_t count,
                              const double B[3], double* out) {
    if (!ctx || !x || !y || !z || !out || count < 0) return gb_fail(ctx, GB_ERR_ARG, "gb_corner_func: bad argument");
    if (kind != GB_SENS_GRAV && kind != GB_SENS_MAGN) return gb_fail(ctx, GB_ERR_ARG, "gb_corner_func: unknown func %d", kind);
    if (kind == GB_SENS_MAGN && !B) return gb_fail(ctx, GB_ERR_ARG, "gb_corner_func: B missing");
    if (count == 0) return GB_OK;
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> d, o;
    GB_CUDA(ctx, d.alloc(3 * count));
    GB_CUDA(ctx, o.alloc(count));
    GB_CUDA(ctx, cudaMemcpyAsync(d.p, x, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(d.p + count, y, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(d.p + 2 * count, z, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const double zeroB[3] = {0, 0, 0};
    GB_CUDA(ctx, launch_corner_func(kind, d.p, d.p + count, d.p + 2 * count, count, B ? B : zeroB, o.p, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(out, o.p, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

extern "C" int gb_a_drill(gb_ctx* ctx, const double* loc, int64_t ndrill, const double* voxelpos, int64_t nvox, double* out) {
    if (!ctx || !voxelpos || !out || ndrill < 0 || nvox < 1 || (ndrill > 0 && !loc)) return gb_fail(ctx, GB_ERR_ARG, "gb_a_drill: bad argument");
    if (ndrill == 0) return GB_OK;
    if (ndrill > 65535) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "gb_a_drill: more than 65535 drill rows");
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> l, v, o;
    GB_CUDA(ctx, l.alloc(3 * ndrill));
    GB_CUDA(ctx, v.alloc(3 * nvox));
    GB_CUDA(ctx, o.alloc(ndrill * nvox));
    GB_CUDA(ctx, cudaMemcpyAsync(l.p, loc, 3 * ndrill * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(v.p, voxelpos, 3 * nvox * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, launch_a_drill(l.p, ndrill, v.p, nvox, o.p, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(out, o.p, ndrill * nvox * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

extern "C" int gb_a_sens(gb_ctx* ctx, int kind, const double B[3], const double* locations, int64_t nsens, const double* edges,
                         const int64_t ncube[3], double mul, double div, double* out) {
    if (!ctx || !B || !locations || !edges || !ncube || !out || nsens < 1) return gb_fail(ctx, GB_ERR_ARG, "gb_a_sens: bad argument");
    if (kind != GB_SENS_GRAV && kind != GB_SENS_MAGN) return gb_fail(ctx, GB_ERR_ARG, "gb_a_sens: unknown func %d", kind);
    const int64_t N = ncube[0] * ncube[1] * ncube[2];
    const int64_t nedge = (ncube[0] + 1) * (ncube[1] + 1) * (ncube[2] + 1);
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> e, l, o;
    GB_CUDA(ctx, e.alloc(3 * nedge));
    GB_CUDA(ctx, l.alloc(3 * nsens));
    GB_CUDA(ctx, o.alloc(nsens * N));
    GB_CUDA(ctx, cudaMemcpyAsync(e.p, edges, 3 * nedge * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(l.p, locations, 3 * nsens * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, launch_a_sens(kind, B, l.p, nsens, e.p, ncube, mul, div, o.p, N, ctx->sm_count, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(out, o.p, nsens * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

// ====================================================================================== geobo/inversion.py
struct gb_problem {
    gb_ctx* ctx = nullptr;
    int64_t n[3] = {0, 0, 0};
    double vox[3] = {0, 0, 0};
    int64_t N = 0, Ns = 0, nd = 0, M = 0, Mp = 0;
    int64_t c0 = 0, c1 = 0, ncol = 0, ncp = 0, ldp = 0;   // local voxel-column shard of Pt; ldp = nrp * ncp (set per call)
    // Property blocks carried through the pipeline.  Without drill data (nd = 0) the reference's third property (drill) comes out
    // as NaN cubes (inversion.py:213-214: std of an empty array), so its block of Pt / V / mean / variance is never computed here:
    // nrp = 2 saves a third of the projection, the triangular solve and the variance product.  cap_nrp sizes the buffers: 3 for
    // cubes small enough for gb_posterior_cov (which needs all three blocks), else nrp.
    int nrp = 3, cap_nrp = 3;
    int64_t Kp = 0, lda = 0, ext = 0, C0 = 0;
    std::vector<int64_t> drill;
    double* A[2] = {nullptr, nullptr};   // [Ns][lda] sensitivities (grav, magn); null in lean mode
    // Lean mode (large cubes: 2 Ns N doubles are 65 GB at 96x96x48, 275 GB at 128x128x64): the fp64 sensitivities are NOT
    // resident.  Only their int8 digit blocks are kept (the operands of the tensor-core products); every fp64 pass over them --
    // digit extraction, A3^T alpha and A3 z of the refinement, gb_forward -- regenerates them in column chunks (voxel rows
    // [y0, y1) of the cube: the voxel index is y-major) with a_sens_kernel, which costs about as much as reading them from HBM.
    bool lean = false;
    double* edges_dev = nullptr;         // [3][(yN+1)(xN+1)(zN+1)] (kept in lean mode)
    double* loc_dev = nullptr;           // [Ns][3]
    double Bfield[3] = {0, 0, 0};
    double grav_mul = 1, grav_div = 1, magn_mul = 1, magn_div = 1;
    double* Achunk[2] = {nullptr, nullptr};   // [Ns][chunk_ld] one column chunk of each survey
    int64_t chunk_y = 0, chunk_ld = 0;   // voxel rows per chunk, leading dimension of the chunk buffers
    double* a_amax = nullptr;            // [2][Ns] running row maxima (digit exponents of rows that are never resident as a whole)
    int* L = nullptr;                    // [Kp] extended-lattice ids
    int64_t* drill_dev = nullptr;
    double* tables = nullptr;            // [9][ext]
    double* Pt = nullptr;                // [Mp][ldp]  Pt = A3 . K, overwritten by V = L^-1 Pt
    double* tmp = nullptr;               // [128][ldp]
    double* Bm = nullptr;                // [Mp][Mp]   AkA -> L
    double* ysol = nullptr;              // [Mp][16]   column 0: y -> u
    double* ytmp = nullptr;              // [128][16]
    double* ydev = nullptr;              // [Mp] staged data vector
    double* linv = nullptr;
    double* scal = nullptr;              // [0] logdet  [1] u.u
    int* info = nullptr;
    double* mu = nullptr;                // [3][ncol]
    double* var = nullptr;               // [3][ncol]
    uint8_t* a8[2] = {nullptr, nullptr};   // int8 digit planes of the sensitivities (tcgen05 path, built lazily)
    int* a_exp[2] = {nullptr, nullptr};
    uint8_t* t8 = nullptr;
    int* t_exp = nullptr;
    unsigned int* sync_ctr = nullptr;    // [4] words: [0] arrival counter of the projection kernel's tile rounds, [2..3] K steps visited (u64)
    double ksteps_total = 0.0;           // K steps of the last projection launch without culling (tiles x steps)
    int* tile_perm = nullptr;            // [2][6][ceil(ncol / 128)] sorted tile order of the projection kernel + its keys
    int* cull_ext = nullptr;             // [9][2] extents of the non-zero table digits (zero-digit culling of the projection's K steps)
    int a8_slices = 0;
    // streamed contraction (lean problems whose full-width digit blocks would not fit: 172 GB at 128x128x64): the projection runs
    // once per column chunk on digit blocks built for that chunk only (a8c) and adds into Pt; the resident digit blocks then cover
    // only this rank's voxel columns (scope 2: N side of the AkA products)
    bool stream_a8 = false;
    uint8_t* a8c[2] = {nullptr, nullptr};
    int a8c_slices = 0;
    int a8_scope = 0;                    // 0: not built; 1: all N contraction columns (dense projection + AkA); 2: only this rank's voxel
                                         // columns (lean + structured projection: the digits are then only the N side of the AkA products)
    // int8 variance path: explicit Linv, its digit blocks, transposed digit blocks of Pt, per-row-tile column sums of squares
    double* Linv = nullptr;              // [Mp][Mp]
    double* tmpL = nullptr;              // [128][Mp]
    double* alpha = nullptr;             // [Mp]  (A K A^T + Sigma)^-1 y
    uint8_t* l8 = nullptr;
    int* l_exp = nullptr;
    uint8_t* b8 = nullptr;
    int* b_exp = nullptr;
    double* partial = nullptr;           // [Mp/128][ldp]
    double* vscratch = nullptr;          // running sums of V when Mp > 16384 (several accumulator flushes)
    double* chol_stage = nullptr;        // packed panel of the distributed Cholesky
    double* chol_pan = nullptr;
    int* chol_paninfo = nullptr;
    int var_slices = 0;
    size_t b8_bytes = 0;
    // fp64 matrix-free refinement scratch
    double* rf_w = nullptr;              // [3][Kp]   A3^T alpha
    double* rf_z = nullptr;              // [3][ncp]  K w on this rank's voxel columns
    double* rf_part = nullptr;           // [8][2][Kp]
    double* rf_t = nullptr;              // [Mp] x 3: t, r, tmp
    // Kronecker-structured products (gb_hyper.structure = GB_STRUCTURE_KRON, exp kernel): factor lines and row-chunk scratch
    double* kron_f = nullptr;            // [9][3][FL]
    double* kron_T = nullptr;            // [3][chunk][nyl][xN * zN]
    long kron_T_doubles = 0;
    // block-Toeplitz FFT products (GB_STRUCTURE_FFT): spectra of the 9 wrapped tables, twiddles, scratch lattices X, Y, Z
    double* fft_W = nullptr;             // [9][Py * Px * Pz]
    cplx* fft_tw = nullptr;              // [3][FFT_MAXP / 2]
    cplx* fft_scratch = nullptr;
    long fft_B = 0;                      // complex row pairs per chunk
    long nlaunch = 0;
    double* y_host_pinned = nullptr;
    double* out_pinned = nullptr;        // [6*ncol + 4]
    cudaEvent_t ev[GB_NUM_TIMERS + 1];
    bool ev_ok = false;
    double ms[GB_NUM_TIMERS];
    uint64_t bytes = 0;
    bool have_data = false;
    bool last_full = false;
};

template <typename T>
static cudaError_t dev_alloc(gb_problem* p, T** ptr, size_t count, bool zero) {
    cudaError_t e = gb_dev_malloc(p->ctx, (void**)ptr, count * sizeof(T));
    if (e != cudaSuccess) return e;
    p->bytes += count * sizeof(T);
    if (zero) e = cudaMemsetAsync(*ptr, 0, count * sizeof(T), p->ctx->stream);
    return e;
}

extern "C" int gb_problem_destroy(gb_problem* p) {
    if (!p) return GB_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    void* ptrs[] = {p->A[0], p->A[1], p->L, p->drill_dev, p->tables, p->Pt, p->tmp, p->Bm, p->ysol, p->ytmp, p->ydev,
                    p->a8[0], p->a8[1], p->a_exp[0], p->a_exp[1], p->t8, p->t_exp, p->cull_ext, p->sync_ctr, p->tile_perm, p->a8c[0], p->a8c[1],
                    p->Linv, p->tmpL, p->alpha, p->l8, p->l_exp, p->b8, p->b_exp, p->partial,
                    p->edges_dev, p->loc_dev, p->Achunk[0], p->Achunk[1], p->a_amax,
                    p->rf_w, p->rf_z, p->rf_part, p->rf_t, p->vscratch, p->chol_stage, p->chol_pan, p->chol_paninfo, p->kron_f, p->kron_T, p->fft_W, p->fft_tw, p->fft_scratch,
                    p->linv, p->scal, p->info, p->mu, p->var};
    for (void* q : ptrs)
        if (q) gb_dev_free(p->ctx, q);
    gb_host_free(p->ctx, p->y_host_pinned);
    gb_host_free(p->ctx, p->out_pinned);
    if (p->ev_ok)
        for (auto& e : p->ev) cudaEventDestroy(e);
    delete p;
    return GB_OK;
}

extern "C" uint64_t gb_problem_device_bytes(gb_problem* p) { return p ? p->bytes : 0; }

// GEOBO_B200_MEMLOG=1: free / total device memory at the allocation milestones, on stderr
static void memlog(gb_problem* p, const char* what) {
    static int on = -1;
    if (on < 0) { const char* ev = getenv("GEOBO_B200_MEMLOG"); on = ev && atoi(ev) != 0; }
    if (!on) return;
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    fprintf(stderr, "[geobo_b200 rank %d] %-28s free %.2f GB of %.2f GB, problem buffers %.2f GB\n", p->ctx->rank, what, fr / 1e9, tot / 1e9, p->bytes / 1e9);
}

// both sensitivity matrices for the voxel rows [y0, y1): out_c[s * ld + (j - y0 xN zN)]   (inversion.py:223-224; gravity is called
// with magneticField * 0)
static cudaError_t sens_generate(gb_problem* p, int y0, int y1, double* out_g, double* out_m, int64_t ld) {
    gb_ctx* ctx = p->ctx;
    const double zeroB[3] = {0.0, 0.0, 0.0};
    cudaError_t e = launch_a_sens_range(GB_SENS_GRAV, zeroB, p->loc_dev, p->Ns, p->edges_dev, p->n, p->grav_mul, p->grav_div, out_g, ld, y0, y1,
                                        ctx->sm_count, ctx->stream);
    if (e != cudaSuccess) return e;
    return launch_a_sens_range(GB_SENS_MAGN, p->Bfield, p->loc_dev, p->Ns, p->edges_dev, p->n, p->magn_mul, p->magn_div, out_m, ld, y0, y1,
                               ctx->sm_count, ctx->stream);
}

// lean mode: f(j0, ncols) for every column chunk, with the chunk's columns [j0, j0 + ncols) regenerated in p->Achunk[0 / 1]
// (only the chunks that intersect the columns [ja, jb) when that range is given)
template <typename F>
static int lean_for_each_chunk(gb_problem* p, F f, int64_t ja = 0, int64_t jb = -1) {
    gb_ctx* ctx = p->ctx;
    const int64_t XZ = p->n[0] * p->n[2], yN = p->n[1];
    for (int64_t y0 = 0; y0 < yN; y0 += p->chunk_y) {
        const int64_t y1 = std::min<int64_t>(yN, y0 + p->chunk_y);
        if (jb >= 0 && (y1 * XZ <= ja || y0 * XZ >= jb)) continue;
        GB_CUDA(ctx, sens_generate(p, (int)y0, (int)y1, p->Achunk[0], p->Achunk[1], p->chunk_ld));
        p->nlaunch += 2;
        GB_CUDA(ctx, f(y0 * XZ, (y1 - y0) * XZ));
    }
    return GB_OK;
}

extern "C" int gb_problem_create(gb_ctx* ctx, const gb_problem_desc* d, gb_problem** out) {
    if (!ctx || !d || !out) return gb_fail(ctx, GB_ERR_ARG, "gb_problem_create: null argument");
    *out = nullptr;
    const int64_t xN = d->ncube[0], yN = d->ncube[1], zN = d->ncube[2];
    if (xN < 1 || yN < 1 || zN < 1) return gb_fail(ctx, GB_ERR_ARG, "gb_problem_create: cube dimensions must be >= 1");
    const int64_t N = xN * yN * zN;
    if ((2 * xN - 1) * (2 * yN - 1) * (2 * zN - 1) >= (1LL << 31)) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "cube too large for 32-bit lattice ids");
    if (!d->edges || !d->locations) return gb_fail(ctx, GB_ERR_ARG, "gb_problem_create: edges/locations missing");
    if (d->nsens != xN * yN)   // sensormodel.py:54,58 hard-wires the sensor count
        return gb_fail(ctx, GB_ERR_ARG, "nsens=%lld but the forward model is defined for xNcube*yNcube=%lld sensors",
                       (long long)d->nsens, (long long)(xN * yN));
    if (d->ndrill < 0 || (d->ndrill > 0 && !d->drill_idx)) return gb_fail(ctx, GB_ERR_ARG, "gb_problem_create: drill_idx missing");
    for (int64_t i = 0; i < d->ndrill; ++i)
        if (d->drill_idx[i] < 0 || d->drill_idx[i] >= N) return gb_fail(ctx, GB_ERR_ARG, "drill_idx[%lld] out of range", (long long)i);
    int64_t c0 = d->col_begin, c1 = d->col_end;
    if (c0 == 0 && c1 == 0) c1 = N;
    if (c0 < 0 || c1 > N || c0 >= c1 || (c0 % 16) != 0) return gb_fail(ctx, GB_ERR_ARG, "bad voxel-column shard [%lld, %lld): begin must be a multiple of 16", (long long)c0, (long long)c1);
    if ((size_t)2 * (xN + 1) * (zN + 1) * 8 > 200 * 1024) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "(xN+1)(zN+1) too large for the A_sens shared-memory planes");

    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    gb_problem* p = new gb_problem();
    p->ctx = ctx;
    for (int i = 0; i < 3; ++i) { p->n[i] = d->ncube[i]; p->vox[i] = d->voxsize[i]; }
    p->N = N; p->Ns = d->nsens; p->nd = d->ndrill; p->M = 2 * p->Ns + p->nd; p->Mp = round_up(p->M, 128);
    p->c0 = c0; p->c1 = c1; p->ncol = c1 - c0; p->ncp = round_up(p->ncol, 32);
    p->cap_nrp = (p->nd == 0 && 3 * N > 46000) ? 2 : 3;
    p->nrp = p->nd == 0 ? 2 : 3;
    p->ldp = p->cap_nrp * p->ncp;                          // allocation width; run_predict sets the width of the call
    p->Kp = round_up(N, 32); p->lda = p->Kp + 32;   // multiples of the largest GEMM slab (BK = 32)
    p->ext = (2 * xN - 1) * (2 * yN - 1) * (2 * zN - 1);
    p->C0 = ((yN - 1) * (2 * xN - 1) + (xN - 1)) * (2 * zN - 1) + (zN - 1);
    p->drill.assign(d->drill_idx, d->drill_idx + d->ndrill);

#define PCUDA(call)                                                                                          \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            int rc__ = gb_fail(ctx, e__ == cudaErrorMemoryAllocation ? GB_ERR_NOMEM : GB_ERR_CUDA, "%s:%d %s: %s", \
                               __FILE__, __LINE__, #call, cudaGetErrorString(e__));                          \
            cudaGetLastError();                                                                              \
            gb_problem_destroy(p);                                                                           \
            return rc__;                                                                                     \
        }                                                                                                    \
    } while (0)

    for (auto& e : p->ev) PCUDA(cudaEventCreate(&e));
    p->ev_ok = true;
    memset(p->ms, 0, sizeof p->ms);
    {
        // lean mode: GEOBO_B200_LEAN_A = 1 / 0 forces it on / off; default: when the two matrices would take more than a fifth of
        // the device memory (they are 8.6 GB at 64x64x32 -- resident --, 65 GB at 96x96x48 -- lean)
        size_t fr = 0, tot = 0;
        PCUDA(cudaMemGetInfo(&fr, &tot));
        const size_t a_bytes = (size_t)2 * p->Ns * p->lda * sizeof(double);
        p->lean = a_bytes > tot / 5;
        if (const char* ev = getenv("GEOBO_B200_LEAN_A")) p->lean = atoi(ev) != 0;
    }
    if (!p->lean) {
        PCUDA(dev_alloc(p, &p->A[0], (size_t)p->Ns * p->lda, true));
        PCUDA(dev_alloc(p, &p->A[1], (size_t)p->Ns * p->lda, true));
    } else {
        // chunk = cy voxel rows; cy * xN * zN must be a multiple of 32 (k steps of the digit blocks), budget per survey
        // GEOBO_B200_LEAN_CHUNK_MB (default 2048)
        const int64_t XZ = xN * zN;
        int64_t unit = 1;
        while ((unit * XZ) % 32 != 0) unit *= 2;
        long mb = 2048;
        if (const char* ev = getenv("GEOBO_B200_LEAN_CHUNK_MB")) mb = atol(ev) > 0 ? atol(ev) : mb;
        int64_t cy = ((int64_t)mb << 20) / (int64_t)sizeof(double) / (p->Ns * XZ);
        if (const char* ev = getenv("GEOBO_B200_LEAN_CHUNK_ROWS")) cy = atol(ev);     // voxel rows per chunk (tests: several ragged chunks)
        cy = cy / unit * unit;
        if (cy < unit) cy = unit;
        if (cy > yN) cy = yN;
        p->chunk_y = cy;
        p->chunk_ld = round_up(cy * XZ, 32);
        PCUDA(dev_alloc(p, &p->Achunk[0], (size_t)p->Ns * p->chunk_ld, true));
        PCUDA(dev_alloc(p, &p->Achunk[1], (size_t)p->Ns * p->chunk_ld, true));
        PCUDA(dev_alloc(p, &p->a_amax, (size_t)2 * p->Ns, true));
    }
    PCUDA(dev_alloc(p, &p->L, (size_t)p->Kp, false));
    PCUDA(dev_alloc(p, &p->drill_dev, (size_t)p->nd + 1, false));
    PCUDA(dev_alloc(p, &p->tables, (size_t)9 * p->ext, false));
    PCUDA(dev_alloc(p, &p->Pt, (size_t)p->Mp * p->ldp, true));
    if (!p->lean) PCUDA(dev_alloc(p, &p->tmp, (size_t)128 * p->ldp, true));      // row-block scratch of the fp64 solve (lean problems never run it)
    PCUDA(dev_alloc(p, &p->Bm, (size_t)p->Mp * p->Mp, true));
    PCUDA(dev_alloc(p, &p->ysol, (size_t)p->Mp * 16, true));
    PCUDA(dev_alloc(p, &p->ytmp, (size_t)128 * 16, true));
    PCUDA(dev_alloc(p, &p->ydev, (size_t)p->Mp, true));
    PCUDA(dev_alloc(p, &p->linv, (size_t)p->Mp * 128, false));
    PCUDA(dev_alloc(p, &p->scal, 4, true));
    PCUDA(dev_alloc(p, &p->info, 4, true));
    PCUDA(dev_alloc(p, &p->mu, (size_t)3 * p->ncol, true));
    PCUDA(dev_alloc(p, &p->var, (size_t)3 * p->ncol, true));
    PCUDA(gb_host_malloc(ctx, (void**)&p->y_host_pinned, (size_t)p->Mp * sizeof(double)));
    PCUDA(gb_host_malloc(ctx, (void**)&p->out_pinned, (size_t)(6 * p->ncol + 8) * sizeof(double)));
    if (p->nd) PCUDA(cudaMemcpyAsync(p->drill_dev, p->drill.data(), p->nd * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    PCUDA(launch_lattice_ids(p->n, p->L, p->Kp, ctx->stream));

    // sensitivities (inversion.py:223-224) computed on the device; edges/locations are the only H2D traffic
    {
        const int64_t nedge = (xN + 1) * (yN + 1) * (zN + 1);
        PCUDA(dev_alloc(p, &p->edges_dev, (size_t)3 * nedge, false));
        PCUDA(dev_alloc(p, &p->loc_dev, (size_t)3 * p->Ns, false));
        for (int i = 0; i < 3; ++i) p->Bfield[i] = d->magnetic_field[i];
        p->grav_mul = d->grav_mul; p->grav_div = d->grav_div; p->magn_mul = d->magn_mul; p->magn_div = d->magn_div;
        PCUDA(cudaMemcpyAsync(p->edges_dev, d->edges, 3 * nedge * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        PCUDA(cudaMemcpyAsync(p->loc_dev, d->locations, 3 * p->Ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        PCUDA(cudaEventRecord(p->ev[0], ctx->stream));
        if (!p->lean) {
            PCUDA(sens_generate(p, 0, (int)yN, p->A[0], p->A[1], p->lda));
        } else {
            for (int64_t y0 = 0; y0 < yN; y0 += p->chunk_y)      // one full generation pass (timed: what every later fp64 pass costs)
                PCUDA(sens_generate(p, (int)y0, (int)std::min<int64_t>(yN, y0 + p->chunk_y), p->Achunk[0], p->Achunk[1], p->chunk_ld));
        }
        PCUDA(cudaEventRecord(p->ev[1], ctx->stream));
        PCUDA(cudaStreamSynchronize(ctx->stream));
        float t = 0.f;
        PCUDA(cudaEventElapsedTime(&t, p->ev[0], p->ev[1]));
        p->ms[GB_T_SENS] = t;
    }
#undef PCUDA
    memlog(p, "problem created");
    *out = p;
    return GB_OK;
}

extern "C" int gb_problem_set_data(gb_problem* p, const double* fs3) {
    if (!p || !fs3) return GB_ERR_ARG;
    gb_ctx* ctx = p->ctx;
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(p->y_host_pinned, fs3, p->M * sizeof(double));
    p->have_data = true;
    return GB_OK;
}

// ---------------------------------------------------------------------------------- small kernels of predict
__global__ void drill_rows_pt_kernel(const double* __restrict__ tables, long ext, long C0, const int* __restrict__ L,
                                     const int64_t* __restrict__ drill, long c0, long ncol, long ncp, long ldp, long row0,
                                     double* __restrict__ Pt) {
    // Pt[(2Ns + d), r*ncp + i] = K[(2, j_d), (r, c0 + i)]  (A_drill is one-hot: a row gather of the covariance)
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int dd = blockIdx.y, r = blockIdx.z;
    if (i >= ncol) return;
    const int lj = L[drill[dd]];
    Pt[(row0 + dd) * ldp + r * ncp + i] = tables[(long)(6 + r) * ext + C0 + (L[c0 + i] - lj)];
}

__global__ void set_y_kernel(const double* __restrict__ y, long M, long Mp, double* __restrict__ ysol) {
    const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mp) return;
    ysol[m * 16] = m < M ? y[m] : 0.0;
}

__global__ void drill_rows_aka_kernel(const double* __restrict__ Pt, long ldp, long ncp, const int64_t* __restrict__ drill,
                                      long c0, long c1, long row0, long M, double* __restrict__ Bm, long ldb) {
    // AkA[(2Ns + d), m] = Pt[m, (2, j_d)] restricted to this rank's columns (other ranks contribute through the all-reduce)
    const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int dd = blockIdx.y;
    if (m >= M) return;
    const long j = drill[dd];
    if (j < c0 || j >= c1) return;
    Bm[(row0 + dd) * ldb + m] = Pt[m * ldp + 2 * ncp + (j - c0)];
}

__global__ void add_noise_diag_kernel(double* __restrict__ Bm, long ldb, long Ns, long M, long Mp, double s0, double s1, double s2) {
    const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mp) return;
    if (m >= M) { Bm[m * ldb + m] = 1.0; return; }          // identity tail of the padded matrix
    const double s = m < Ns ? s0 : (m < 2 * Ns ? s1 : s2);
    Bm[m * ldb + m] += s * s;                                // inversion.py:94-96: + diag(yerr**2)
}

// mu[col] = sum_m V[m,col] u[m];  var[col] = amp - sum_m V[m,col]^2   (inversion.py:115,117 diag only; Q10 diag(K) = amp)
__global__ void __launch_bounds__(256) mean_var_kernel(const double* __restrict__ V, long ldp, long ncp, long ncol, long M,
                                                       const double* __restrict__ ysol, double amp, double* __restrict__ mu,
                                                       double* __restrict__ var) {
    __shared__ double us[1024];
    const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;   // over ncp
    const int r = blockIdx.y;
    double m0 = 0.0, m1 = 0.0, s0 = 0.0, s1 = 0.0;
    const bool live = col < ncol;
    const double* v = V + r * ncp + (live ? col : 0);
    for (long mb = 0; mb < M; mb += 1024) {
        __syncthreads();
        for (int q = threadIdx.x; q < 1024; q += blockDim.x) us[q] = (mb + q < M) ? ysol[(mb + q) * 16] : 0.0;
        __syncthreads();
        const int lim = (int)min(1024L, M - mb);
        int q = 0;
        for (; q + 1 < lim; q += 2) {
            const double a = v[(mb + q) * ldp], b = v[(mb + q + 1) * ldp];
            m0 = fma(a, us[q], m0); s0 = fma(a, a, s0);
            m1 = fma(b, us[q + 1], m1); s1 = fma(b, b, s1);
        }
        if (q < lim) { const double a = v[(mb + q) * ldp]; m0 = fma(a, us[q], m0); s0 = fma(a, a, s0); }
    }
    if (live) {
        mu[r * ncol + col] = m0 + m1;
        var[r * ncol + col] = amp - (s0 + s1);
    }
}

__global__ void dot_self_kernel(const double* __restrict__ ysol, long M, double* __restrict__ out) {
    __shared__ double red[256];
    double acc = 0.0;
    for (long m = threadIdx.x; m < M; m += blockDim.x) { const double u = ysol[m * 16]; acc = fma(u, u, acc); }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

static gemm::Task mk(const double* A, long lda, const double* B, long ldb, double* C, long ldc, int M, int N, int K, int lower) {
    gemm::Task t;
    memset(&t, 0, sizeof t);
    t.A = A; t.lda = lda; t.B = B; t.ldb = ldb; t.C = C; t.ldc = ldc; t.M = M; t.N = N; t.K = K;
    t.alpha = 1.0; t.beta = 0.0; t.lower = lower;
    return t;
}

// Runs the device pipeline up to `stage`: 0 = through Cholesky + u (logl only), 1 = everything.
static int run_predict(gb_problem* p, const gb_hyper* h, bool full, bool all_blocks = false) {
    gb_ctx* ctx = p->ctx;
    cudaStream_t s = ctx->stream;
    if (!p->have_data) return gb_fail(ctx, GB_ERR_ARG, "gb_predict before gb_problem_set_data");
    const int nrp = (p->nd == 0 && !all_blocks) ? 2 : 3;
    if (nrp > p->cap_nrp) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "all three property blocks requested on a problem sized for two");
    p->nrp = nrp;
    p->ldp = (int64_t)nrp * p->ncp;
    p->ksteps_total = 0.0;
    p->last_full = full;
    p->nlaunch = 0;
    CovParams cp;
    GB_TRY(fill_cov_params(ctx, cp, h->kernel_id, h->gp_length, h->coeffm, h->gp_amp));
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    const long Ns = p->Ns, M = p->M, Mp = p->Mp, ldp = p->ldp, ncp = p->ncp, ncol = p->ncol;

    GB_CUDA(ctx, cudaEventRecord(p->ev[0], s));
    // ---- reset pads / accumulators
    GB_CUDA(ctx, cudaMemsetAsync(p->Bm, 0, (size_t)Mp * Mp * sizeof(double), s));
    GB_CUDA(ctx, cudaMemsetAsync(p->scal, 0, 4 * sizeof(double), s));
    GB_CUDA(ctx, cudaMemsetAsync(p->info, 0, 4 * sizeof(int), s));
    if (Mp > M) GB_CUDA(ctx, cudaMemsetAsync(p->Pt + M * ldp, 0, (size_t)(Mp - M) * ldp * sizeof(double), s));
    if (nrp < 3) {   // the drill property without drill data: NaN like the reference's cubes (all-ones bit pattern = quiet NaN)
        GB_CUDA(ctx, cudaMemsetAsync(p->mu + 2 * ncol, 0xFF, (size_t)ncol * sizeof(double), s));
        GB_CUDA(ctx, cudaMemsetAsync(p->var + 2 * ncol, 0xFF, (size_t)ncol * sizeof(double), s));
    }
    if (ncp > ncol)
        for (int r = 0; r < nrp; ++r)
            GB_CUDA(ctx, cudaMemset2DAsync(p->Pt + r * ncp + ncol, ldp * sizeof(double), 0, (ncp - ncol) * sizeof(double), Mp, s));
    GB_CUDA(ctx, cudaMemcpyAsync(p->ydev, p->y_host_pinned, M * sizeof(double), cudaMemcpyHostToDevice, s));
    set_y_kernel<<<(unsigned)((Mp + 255) / 256), 256, 0, s>>>(p->ydev, M, Mp, p->ysol);
    GB_CUDA(ctx, cudaGetLastError());
    // ---- stationary covariance tables (kernels.create_cov evaluated once per lattice offset)
    // a non-finite covariance value is reported like a failed factorisation (info = 1): the reference's dense products spread it
    // over all of AkA and its Cholesky raises (inversion.py:98-104)
    GB_CUDA(ctx, launch_cov_tables(cp, p->n, p->vox, p->tables, s, p->info));
    GB_CUDA(ctx, cudaEventRecord(p->ev[1], s));

    // ---- opt-in structure-exploiting path: factor lines of the separable exp blocks from the tables
    const bool kron = h->structure == GB_STRUCTURE_KRON, compact = h->structure == GB_STRUCTURE_COMPACT, fft = h->structure == GB_STRUCTURE_FFT;
    const bool structured = kron || compact || fft;
    if (h->structure != GB_STRUCTURE_DENSE && !structured)
        return gb_fail(ctx, GB_ERR_ARG, "gb_hyper.structure must be GB_STRUCTURE_DENSE, _KRON, _COMPACT or _FFT; got %d", h->structure);
    const FftGeom fg = fft_geom(p->n[0], p->n[1], p->n[2], p->c0, p->c1);
    const KronGeom kg = kron_geom(p->n[0], p->n[1], p->n[2], p->c0, p->c1);
    const StencilGeom sg = stencil_geom(p->n[0], p->n[1], p->n[2], p->c0, p->c1, p->vox, h->gp_length);
    if (p->lean && h->slices == 0)
        return gb_fail(ctx, GB_ERR_UNSUPPORTED, "this problem keeps no resident fp64 sensitivities (lean mode, %lld x %lld): only the int8 "
                       "tensor-core paths (slices = 4, 5, 6) run on it; set GEOBO_B200_LEAN_A=0 if the matrices fit",
                       (long long)p->Ns, (long long)p->N);
    if (compact && h->kernel_id != GB_KERNEL_SPARSE)
        return gb_fail(ctx, GB_ERR_UNSUPPORTED, "structure = compact needs kernelfunc 'sparse': only the compact-support kernels (kernels.py:101-138) "
                       "vanish outside a window of the voxel grid; use structure = dense");
    // out[s][r * ncp + (j - c0)] (+)= sum_i A[s][i] K_(blk0 + r)[i][j] through the structured form of the blocks
    auto apply_structured = [&](int blk0, const double* A, long lda, long nrows, double* out, long ldo, int accumulate) -> cudaError_t {
        if (kron) return kron_apply(kg, p->kron_f, blk0, A, lda, nrows, p->kron_T, p->kron_T_doubles, out, ldo, ncp, accumulate, s, &p->nlaunch, nrp);
        if (fft) return fft_apply(fg, p->fft_W, p->fft_tw, blk0, A, lda, nrows, p->fft_scratch, p->fft_B, out, ldo, ncp, accumulate, s, &p->nlaunch, nrp);
        return stencil_apply(sg, p->tables + (long)blk0 * p->ext + p->C0, A, lda, nrows, out, ldo, ncp, accumulate, s, &p->nlaunch, nrp);
    };
    if (kron) {
        if (h->kernel_id != GB_KERNEL_EXP)
            return gb_fail(ctx, GB_ERR_UNSUPPORTED, "structure = kron needs kernelfunc 'exp': only the squared-exponential blocks (kernels.py:81-99) are "
                           "Kronecker products on the voxel grid; use structure = dense");
        char why[256];
        if (!kron_supported(kg, why, sizeof why)) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "%s", why);
        if (!p->kron_f) {
            GB_CUDA(ctx, dev_alloc(p, &p->kron_f, (size_t)kron_factor_doubles(kg), false));
            p->kron_T_doubles = kron_scratch_doubles(kg, Ns);
            GB_CUDA(ctx, dev_alloc(p, &p->kron_T, (size_t)p->kron_T_doubles, false));
        }
        GB_CUDA(ctx, kron_build_factors(p->tables, p->ext, p->C0, kg, p->kron_f, s));
        p->nlaunch += 1;
    }
    if (fft) {
        char why[256];
        if (!fft_supported(fg, why, sizeof why)) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "%s", why);
        if (!p->fft_W) {
            p->fft_B = fft_chunk_pairs(fg, Ns);
            GB_CUDA(ctx, dev_alloc(p, &p->fft_W, (size_t)9 * fg.P3, false));
            GB_CUDA(ctx, dev_alloc(p, &p->fft_tw, (size_t)3 * (FFT_MAXP / 2), false));
            GB_CUDA(ctx, dev_alloc(p, &p->fft_scratch, (size_t)fft_scratch_cplx(fg, p->fft_B), false));
            GB_CUDA(ctx, fft_build_twiddles(fg, p->fft_tw, s));
            p->nlaunch += 3;
        }
        // spectra of the 9 wrapped tables (they change with the hyper-parameters): X, Y = the first two scratch lattices
        GB_CUDA(ctx, fft_build_spectra(fg, p->tables, p->ext, p->C0, p->fft_tw, p->fft_scratch, p->fft_scratch + p->fft_B * fg.P3, p->fft_W, s, &p->nlaunch));
    }

    // ---- Pt = A3 . K : fused assembly + projection, 6 (data block c, property block r) products
    if (h->slices != 0) {
        // int8 digit-slice products on tcgen05 / TMEM (ozaki.cu); everything downstream stays fp64
        const int S = h->slices;
        if (ozaki_tile_n(S) == 0) return gb_fail(ctx, GB_ERR_ARG, "gb_hyper.slices must be 0 (fp64) or 4, 5, 6; got %d", S);
        if (p->n[2] % 16 != 0 || ncol % 16 != 0)
            return gb_fail(ctx, GB_ERR_UNSUPPORTED, "the int8 tensor-core path needs zNcube %% 16 == 0 and a voxel-column shard that is a multiple of 16 "
                           "(got zNcube=%lld, columns=%ld); use slices = 0", (long long)p->n[2], ncol);
        // scope of the digit blocks: the dense projection contracts over all N columns; with a structured projection on a lean
        // problem they are only the N side of the AkA products, i.e. this rank's voxel columns (128x128x64: 21 GB instead of 172)
        if (p->lean && !structured && (p->a8_slices != S || p->a8_scope != 1)) {      // (also when a structured run left shard-only blocks)
            size_t fr = 0, tot = 0;
            GB_CUDA(ctx, cudaMemGetInfo(&fr, &tot));
            p->stream_a8 = 2 * (size_t)ozaki_rows_bytes(Ns, p->Kp, S, ozaki_tile_np(S)) > tot / 4;
            if (const char* ev = getenv("GEOBO_B200_STREAM_A8")) p->stream_a8 = atoi(ev) != 0;
        }
        const bool streamed = p->lean && !structured && p->stream_a8;
        const int want_scope = (p->lean && (structured || streamed)) ? 2 : 1;
        if (p->a8_slices != S || (p->a8_scope != want_scope && !(p->a8_scope == 1 && want_scope == 2))) {
            // digit planes of the sensitivities: built once per problem, slice count and scope
            const long a8_kp = want_scope == 2 ? ncp : p->Kp;
            for (int c = 0; c < 2; ++c) {
                if (p->a8[c]) { gb_dev_free(ctx, p->a8[c]); p->a8[c] = nullptr; }
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->a8[c], (size_t)ozaki_rows_bytes(Ns, a8_kp, S, ozaki_tile_np(S))));
                if (!p->a_exp[c]) GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->a_exp[c], (size_t)Ns * sizeof(int)));
                if (!p->lean) GB_CUDA(ctx, ozaki_slice_sens(p->A[c], Ns, p->N, p->lda, S, p->a_exp[c], p->a8[c], p->Kp, s));
            }
            if (p->lean) {
                // two generation passes: running row maxima -> exponents, then the digits of every chunk into its k steps
                GB_CUDA(ctx, cudaMemsetAsync(p->a_amax, 0, (size_t)2 * Ns * sizeof(double), s));
                GB_TRY(lean_for_each_chunk(p, [&](int64_t, int64_t ncols) -> cudaError_t {
                    for (int c = 0; c < 2; ++c) {
                        cudaError_t e = ozaki_row_absmax_accum(p->Achunk[c], Ns, ncols, p->chunk_ld, p->a_amax + c * Ns, s);
                        if (e != cudaSuccess) return e;
                    }
                    return cudaSuccess;
                }));
                for (int c = 0; c < 2; ++c) {
                    GB_CUDA(ctx, ozaki_exps_from_absmax(p->a_amax + c * Ns, Ns, p->a_exp[c], s));
                    GB_CUDA(ctx, cudaMemsetAsync(p->a8[c], 0, (size_t)ozaki_rows_bytes(Ns, a8_kp, S, ozaki_tile_np(S)), s));
                }
                GB_TRY(lean_for_each_chunk(p, [&](int64_t j0, int64_t ncols) -> cudaError_t {
                    // columns [ja, jb) of this chunk go into the digit blocks: all of them, or the part inside this rank's shard
                    const int64_t ja = want_scope == 2 ? std::max<int64_t>(j0, p->c0) : j0;
                    const int64_t jb = want_scope == 2 ? std::min<int64_t>(j0 + ncols, p->c1) : j0 + ncols;
                    if (jb <= ja) return cudaSuccess;
                    const int64_t kbase = want_scope == 2 ? p->c0 : 0;
                    for (int c = 0; c < 2; ++c) {
                        cudaError_t e = ozaki_slice_rows_range(p->Achunk[c] + (ja - j0), Ns, jb - ja, p->chunk_ld, S, p->a_exp[c], p->a8[c], a8_kp,
                                                               ozaki_tile_np(S), (ja - kbase) / 32, s);
                        if (e != cudaSuccess) return e;
                    }
                    return cudaSuccess;
                }));
            }
            p->a8_scope = want_scope;
            if (p->t8) { gb_dev_free(ctx, p->t8); p->t8 = nullptr; }
            GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->t8, (size_t)ozaki_table_bytes(p->ext, S)));
            if (!p->t_exp) GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->t_exp, 16 * sizeof(int)));
            if (!p->cull_ext) GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->cull_ext, 18 * sizeof(int)));
            if (!p->sync_ctr) GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->sync_ctr, 4 * sizeof(unsigned int)));
            if (!p->tile_perm) GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->tile_perm, (size_t)2 * 6 * ((ncol + 127) / 128) * sizeof(int)));
            // digit scratch shared by the AkA products (row digits of three Pt blocks) and the variance product
            // (transposed digits of all of Pt): the two uses are sequential
            if (p->b8) { gb_dev_free(ctx, p->b8); p->b8 = nullptr; }
            if (p->b_exp) { gb_dev_free(ctx, p->b_exp); p->b_exp = nullptr; }
            // (a) row digits of ONE Pt block (M side of an AkA product; the three products run one after the other through it),
            // (b) transposed digits of a column chunk of Pt (N side of the variance product): the scratch is the larger of (a) and
            // min(all of Pt, GEOBO_B200_B8_MB, default 12 GB) -- the variance product walks Pt in column chunks of that size
            const size_t aka_bytes = (size_t)ozaki_rows_bytes(Ns, ncp, S, 128);
            const size_t var_all = (size_t)ozaki_cols_bytes(ldp, Mp, S);
            size_t var_budget = (size_t)(p->lean ? 8192 : 12288) << 20;      // lean problems are the memory-tight ones
            if (const char* ev = getenv("GEOBO_B200_B8_MB")) var_budget = (size_t)(atol(ev) > 0 ? atol(ev) : 12288) << 20;
            const size_t var_bytes = var_all < var_budget ? var_all : var_budget;
            p->b8_bytes = aka_bytes > var_bytes ? aka_bytes : var_bytes;
            GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->b8, p->b8_bytes));
            GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->b_exp, (size_t)(ldp > 3 * Ns ? ldp : 3 * Ns) * sizeof(int)));
            p->bytes += 2 * (size_t)ozaki_rows_bytes(Ns, a8_kp, S, ozaki_tile_np(S)) + p->b8_bytes;
            p->a8_slices = S;
            memlog(p, "digit blocks + scratch");
        }
        if (streamed && p->a8c_slices != S) {
            for (int c = 0; c < 2; ++c) {
                if (p->a8c[c]) { gb_dev_free(ctx, p->a8c[c]); p->a8c[c] = nullptr; }
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->a8c[c], (size_t)ozaki_rows_bytes(Ns, p->chunk_ld, S, ozaki_tile_np(S))));
                p->bytes += (size_t)ozaki_rows_bytes(Ns, p->chunk_ld, S, ozaki_tile_np(S));
            }
            p->a8c_slices = S;
        }
    }
    if (structured) {
        // fft: zero-padded 3-D FFT convolutions with the stationary tables (any kernel)
        // kron: three Toeplitz mode products per block (rows of A_c -> y mode into the L2-resident scratch -> z and x modes -> rows of Pt)
        // compact: tap sum over the support window of the compact kernels
        if (!p->lean) {
            for (int c = 0; c < 2; ++c) GB_CUDA(ctx, apply_structured(c * 3, p->A[c], p->lda, Ns, p->Pt + (long)c * Ns * ldp, ldp, 0));
        } else {
            // lean problem: the rows of A_c are regenerated in sensor-row chunks (full width) through the chunk buffer
            const double zeroB[3] = {0.0, 0.0, 0.0};
            long rows_chunk = (long)((p->Ns * p->chunk_ld) / p->lda);
            if (rows_chunk < 1) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "lean chunk buffer smaller than one sensitivity row; raise GEOBO_B200_LEAN_CHUNK_MB");
            if (const char* ev = getenv("GEOBO_B200_LEAN_ROW_CHUNK")) { const long v = atol(ev); if (v >= 1 && v < rows_chunk) rows_chunk = v; }
            for (int c = 0; c < 2; ++c)
                for (long s0 = 0; s0 < Ns; s0 += rows_chunk) {
                    const long nrow = Ns - s0 < rows_chunk ? Ns - s0 : rows_chunk;
                    GB_CUDA(ctx, cudaMemsetAsync(p->Achunk[0], 0, (size_t)nrow * p->lda * sizeof(double), s));      // pad columns of the rows
                    GB_CUDA(ctx, launch_a_sens_range(c == 0 ? GB_SENS_GRAV : GB_SENS_MAGN, c == 0 ? zeroB : p->Bfield, p->loc_dev + 3 * s0, nrow, p->edges_dev,
                                                     p->n, c == 0 ? p->grav_mul : p->magn_mul, c == 0 ? p->grav_div : p->magn_div, p->Achunk[0], p->lda, 0,
                                                     (int)p->n[1], ctx->sm_count, s));
                    GB_CUDA(ctx, apply_structured(c * 3, p->Achunk[0], p->lda, nrow, p->Pt + ((long)c * Ns + s0) * ldp, ldp, 0));
                    p->nlaunch += 1;
                }
        }
    } else if (h->slices != 0) {
        const int S = h->slices;
        GB_CUDA(ctx, ozaki_slice_tables(p->tables, p->ext, S, p->t_exp, p->t8, s));
        OzakiArgs oa;
        oa.a8[0] = p->a8[0]; oa.a8[1] = p->a8[1]; oa.a_exp[0] = p->a_exp[0]; oa.a_exp[1] = p->a_exp[1];
        oa.t8 = p->t8; oa.t_exp = p->t_exp; oa.L = p->L; oa.Pt = p->Pt;
        oa.ext = p->ext; oa.C0 = p->C0; oa.kp = p->Kp; oa.ldp = ldp; oa.ncp = ncp;
        oa.Ns = (int)Ns; oa.ncol = (int)ncol; oa.c0 = (int)p->c0; oa.nr = nrp;
        // zero-digit culling (bitwise-neutral, default on; GEOBO_B200_CULL=0 visits every K step)
        bool cull = true;
        if (const char* ev = getenv("GEOBO_B200_CULL")) cull = atoi(ev) != 0;
        oa.cull = nullptr;
        if (cull) {
            GB_CUDA(ctx, ozaki_table_extents(p->t8, p->ext, S, p->n, p->cull_ext, s));
            oa.cull = p->cull_ext;
            p->nlaunch += 1;
        }
        oa.n[0] = (int)p->n[0]; oa.n[1] = (int)p->n[1]; oa.n[2] = (int)p->n[2];
        // tile-round pacing (default on; GEOBO_B200_TILE_SYNC=0: free-running CTAs)
        // GEOBO_B200_TILE_SYNC: 0 = free-running CTAs, 1 = all CTAs start a round together, n > 1 = a CTA may run n - 1 rounds ahead.
        // Default (measured, profiles/r2_bench_n_*.json): strict when a sensor-row tile spans several rounds (>= 4 x SM count
        // voxel-column tiles: 64x64x32 on one GPU -- 42.8 k vs 41.5 k voxels/s, and the lowest HBM traffic), else one round of
        // slack (32^3, multi-GPU shards: few rounds per sensor-row tile, the round boundaries weigh more -- 911 k vs 875 k at 32^3)
        int pace_mode = ((ncol + 127) / 128 >= 4L * ctx->sm_count) ? 1 : 2;
        if (const char* ev = getenv("GEOBO_B200_TILE_SYNC")) pace_mode = atoi(ev);
        const bool pace = pace_mode > 0;
        oa.sync_slack = pace_mode > 1 ? pace_mode - 1 : 0;
        // GEOBO_B200_TILE_SORT=0: natural tile order (default: tiles of a round sorted to equal K-step counts)
        bool tsort = true;
        if (const char* ev = getenv("GEOBO_B200_TILE_SORT")) tsort = atoi(ev) != 0;
        oa.perm_scratch = tsort ? p->tile_perm : nullptr;
        if (tsort && cull) p->nlaunch += 2;
        oa.sync_ctr = nullptr;
        GB_CUDA(ctx, cudaMemsetAsync(p->sync_ctr, 0, 4 * sizeof(unsigned int), s));
        if (pace) oa.sync_ctr = p->sync_ctr;
        oa.steps_ctr = reinterpret_cast<unsigned long long*>(p->sync_ctr + 2);
        {
            const int nt_ = ozaki_tile_np(S);
            p->ksteps_total = 2.0 * nrp * (double)((Ns + nt_ - 1) / nt_) * (double)((ncol + 127) / 128) * (double)(p->Kp / 32);
        }
        oa.ks_base = 0; oa.cy0 = 0; oa.cy1 = 0; oa.accumulate = 0;
        if (!(p->lean && p->stream_a8)) {
            GB_CUDA(ctx, ozaki_project(oa, S, ctx->sm_count, s));
        } else {
            // streamed contraction: Pt = sum over column chunks of (digits of the chunk's columns) . (covariance rows of the chunk)
            int eymax = (int)p->n[1];
            if (cull) {        // chunks further than the largest y extent from every voxel row of the shard hold only skipped steps
                int ext_h[18];
                GB_CUDA(ctx, cudaMemcpyAsync(ext_h, p->cull_ext, sizeof ext_h, cudaMemcpyDeviceToHost, s));
                GB_CUDA(ctx, cudaStreamSynchronize(s));
                eymax = 0;
                for (int c = 0; c < 2; ++c)
                    for (int r = 0; r < nrp; ++r) eymax = std::max(eymax, ext_h[2 * (c * 3 + r)]);
            }
            const int64_t XZ = p->n[0] * p->n[2];
            const int64_t sy0 = p->c0 / XZ, sy1 = (p->c1 - 1) / XZ;
            GB_CUDA(ctx, cudaMemsetAsync(p->Pt, 0, (size_t)2 * Ns * ldp * sizeof(double), s));
            GB_TRY(lean_for_each_chunk(p, [&](int64_t j0, int64_t ncols) -> cudaError_t {
                const int64_t y0 = j0 / XZ, y1 = (j0 + ncols) / XZ;
                if (y1 <= sy0 - eymax || y0 > sy1 + eymax) return cudaSuccess;
                for (int c = 0; c < 2; ++c) {
                    cudaError_t e = ozaki_slice_rows_range(p->Achunk[c], Ns, ncols, p->chunk_ld, S, p->a_exp[c], p->a8c[c], round_up(ncols, 32),
                                                           ozaki_tile_np(S), 0, s);
                    if (e != cudaSuccess) return e;
                }
                OzakiArgs ob = oa;
                ob.a8[0] = p->a8c[0]; ob.a8[1] = p->a8c[1];
                ob.kp = round_up(ncols, 32);
                ob.ks_base = (int)(j0 / 32); ob.cy0 = (int)y0; ob.cy1 = (int)y1; ob.accumulate = 1;
                cudaError_t e = cudaMemsetAsync(p->sync_ctr, 0, sizeof(unsigned int), s);
                if (e != cudaSuccess) return e;
                p->nlaunch += 3;
                return ozaki_project(ob, S, ctx->sm_count, s);
            }));
        }
        p->nlaunch += 2;                  // table slicing (absmax + digits)
    } else {
        gemm::TaskBatch b;
        b.n = 0;
        for (int c = 0; c < 2; ++c)
            for (int r = 0; r < nrp; ++r) {
                gemm::Task t = mk(p->A[c], p->lda, p->tables + (long)(c * 3 + r) * p->ext + p->C0, 0,
                                  p->Pt + (long)c * Ns * ldp + r * ncp, ldp, (int)Ns, (int)ncol, (int)p->Kp, 0);
                t.Lrow = p->L + p->c0;
                t.Lcol = p->L;
                b.t[b.n++] = t;
            }
        GB_CUDA(ctx, gemm::launch(b, gemm::B_GEN, s));
    }
    GB_CUDA(ctx, cudaEventRecord(p->ev[2], s));
    if (p->nd) {
        dim3 grid((unsigned)((ncol + 255) / 256), (unsigned)p->nd, 3);
        drill_rows_pt_kernel<<<grid, 256, 0, s>>>(p->tables, p->ext, p->C0, p->L, p->drill_dev, p->c0, ncol, ncp, ldp, 2 * Ns, p->Pt);
        GB_CUDA(ctx, cudaGetLastError());
    }
    GB_CUDA(ctx, cudaEventRecord(p->ev[3], s));

    // ---- AkA = A3 . Pt^T (lower triangle) over this rank's voxel columns
    if (h->slices != 0) {
        // int8 digit products.  Block (c', c) of AkA: C[s'][s] = sum_j Pt[(c', s'), (c, j)] * A_c[s][j]: the Pt block is sliced
        // row-wise into M-side digit blocks (128-row tiles), the sensitivities' N-side digit blocks of the projection are
        // reused (K steps c0/32 ..).  Lower triangle: blocks (0,0) and (1,1) lower, (1,0) full.
        const int S = h->slices;
        const size_t blk = (size_t)ozaki_rows_bytes(Ns, ncp, S, 128);
        // N side = the sensitivities' digit blocks: K steps c0 / 32 .. of the full-width blocks, or the shard-only blocks from step 0
        const int a_ksteps = p->a8_scope == 2 ? (int)(ncp / 32) : (int)(p->Kp / 32), a_k0 = p->a8_scope == 2 ? 0 : (int)(p->c0 / 32), ks = (int)(ncp / 32);
        const int cc[3][2] = {{0, 0}, {1, 0}, {1, 1}};      // (c', c): block row = Pt rows of c', block column = A_c
        for (int t = 0; t < 3; ++t) {
            const int cp_ = cc[t][0], c = cc[t][1];
            (void)blk;           // one block buffer, reused by the three stream-ordered products
            GB_CUDA(ctx, ozaki_slice_rows(p->Pt + (long)cp_ * Ns * ldp + (long)c * ncp, Ns, ncp, ldp, S, p->b_exp + t * Ns, p->b8, ncp, 128, s));
            GB_CUDA(ctx, ozaki_gemm_store(p->b8, p->b_exp + t * Ns, ks, 0, p->a8[c], p->a_exp[c], a_ksteps, a_k0, ks, (int)Ns, (int)Ns,
                                          p->Bm + (long)cp_ * Ns * Mp + (long)c * Ns, Mp, c == cp_ ? 1 : 0, S, ctx->sm_count, s));
        }
        p->nlaunch += 9;                  // 3 x (row absmax, row slicing, GEMM)
        if (p->nd) {
            dim3 grid((unsigned)((M + 255) / 256), (unsigned)p->nd);
            drill_rows_aka_kernel<<<grid, 256, 0, s>>>(p->Pt, ldp, ncp, p->drill_dev, p->c0, p->c1, 2 * Ns, M, p->Bm, Mp);
            GB_CUDA(ctx, cudaGetLastError());
        }
    } else {

        gemm::TaskBatch b;
        b.n = 3;
        b.t[0] = mk(p->A[0] + p->c0, p->lda, p->Pt, ldp, p->Bm, Mp, (int)Ns, (int)Ns, (int)ncp, 1);
        b.t[1] = mk(p->A[1] + p->c0, p->lda, p->Pt + ncp, ldp, p->Bm + Ns * Mp, Mp, (int)Ns, (int)Ns, (int)ncp, 0);
        b.t[2] = mk(p->A[1] + p->c0, p->lda, p->Pt + Ns * ldp + ncp, ldp, p->Bm + Ns * Mp + Ns, Mp, (int)Ns, (int)Ns, (int)ncp, 1);
        GB_CUDA(ctx, gemm::launch(b, gemm::B_T, s));
        if (p->nd) {
            dim3 grid((unsigned)((M + 255) / 256), (unsigned)p->nd);
            drill_rows_aka_kernel<<<grid, 256, 0, s>>>(p->Pt, ldp, ncp, p->drill_dev, p->c0, p->c1, 2 * Ns, M, p->Bm, Mp);
            GB_CUDA(ctx, cudaGetLastError());
        }
    }
    GB_CUDA(ctx, cudaEventRecord(p->ev[4], s));
    GB_TRY(comm_allreduce_sum_f64(ctx, p->Bm, (size_t)Mp * Mp));
    add_noise_diag_kernel<<<(unsigned)((Mp + 255) / 256), 256, 0, s>>>(p->Bm, Mp, Ns, M, Mp, h->gp_sigma[0], h->gp_sigma[1], h->gp_sigma[2]);
    GB_CUDA(ctx, cudaGetLastError());
    GB_CUDA(ctx, cudaEventRecord(p->ev[5], s));

    // ---- Cholesky (inversion.py:100) and u = L^-1 y (:105)
    CholWork w;
    w.linv = p->linv; w.logdet = p->scal; w.info = p->info;
    // multi-GPU: block-cyclic factorisation with NCCL panel broadcasts once the matrix is large enough for the trailing
    // updates to matter (M >= 8192), replicated otherwise (the per-panel broadcast latency would dominate);
    // GEOBO_B200_DIST_CHOL=1/0 forces either.
    bool dist_chol = ctx->nranks > 1 && Mp >= 8192;
    if (const char* e = getenv("GEOBO_B200_DIST_CHOL")) dist_chol = ctx->nranks > 1 && atoi(e) != 0;
    if (dist_chol) {
        if (!p->chol_stage) {
            GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->chol_stage, ((size_t)Mp * 128 + 128 * 128) * sizeof(double)));
            GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->chol_pan, (size_t)2 * (Mp / 128) * sizeof(double)));
            GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->chol_paninfo, (size_t)(Mp / 128) * sizeof(int)));
        }
        GB_TRY(chol_factor_dist(ctx, p->Bm, Mp, (int)Mp, (int)M, w, p->chol_stage, p->chol_pan, p->chol_paninfo, s));
    } else {
        GB_CUDA(ctx, chol_factor(p->Bm, Mp, (int)Mp, (int)M, w, s));
    }
    GB_CUDA(ctx, cudaEventRecord(p->ev[6], s));
    const bool int8_var = full && h->slices != 0;
    if (!int8_var) {
        GB_CUDA(ctx, chol_forward_solve(p->Bm, Mp, (int)Mp, w, p->ysol, 16, 1, p->ytmp, s));
        dot_self_kernel<<<1, 256, 0, s>>>(p->ysol, M, p->scal + 1);
        GB_CUDA(ctx, cudaGetLastError());
        p->nlaunch += 2 * (Mp / 128) + 1;
    }
    if (int8_var) {
        // ---- variance and mean without ever forming V = L^-1 Pt (:114-117) in memory:
        //   Linv = L^-1 explicitly (recursive doubling from the diagonal-block inverses), u = Linv y, alpha = Linv^T u,
        //   refinement of alpha against the fp64 matrix-free operator and mu = K A3^T alpha (:115),
        //   one pass over Pt for its transposed digit blocks, then colsumsq(Linv . Pt) on the int8 tensor cores with the
        //   reduction in the epilogue (:117, diag only).
        const int S = h->slices;
        if (p->var_slices != S) {
            if (!p->Linv) {
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->Linv, (size_t)Mp * Mp * sizeof(double)));
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->alpha, (size_t)Mp * sizeof(double)));
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->l_exp, (size_t)Mp * sizeof(int)));
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->partial, (size_t)(Mp / 128) * ldp * sizeof(double)));
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->rf_w, (size_t)3 * p->Kp * sizeof(double)));
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->rf_z, (size_t)8 * 3 * ncp * sizeof(double)));
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->rf_part, (size_t)16 * p->Kp * sizeof(double)));
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->rf_t, (size_t)3 * Mp * sizeof(double)));
                GB_CUDA(ctx, cudaMemsetAsync(p->rf_t, 0, (size_t)3 * Mp * sizeof(double), s));
                p->bytes += (size_t)Mp * Mp * 8 + (size_t)(Mp / 128) * ldp * 8;
            }
            if (p->vscratch) { gb_dev_free(ctx, p->vscratch); p->vscratch = nullptr; }
            if (ozaki_colsumsq_scratch_bytes((int)Mp, S, ctx->sm_count))
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->vscratch, (size_t)ozaki_colsumsq_scratch_bytes((int)Mp, S, ctx->sm_count)));
            p->var_slices = S;
            memlog(p, "variance-stage buffers");
        }
        // Two buffers of this stage live inside others whose contents are dead by then (at 128x128x64 they would be 8.7 + 5.4 GB):
        //  * the Mp x Mp scratch of the triangular inverse sits in the digit scratch b8 (idle between the AkA products and the
        //    variance product) when that is large enough, else in its own allocation;
        //  * the digit blocks of Linv (Mp^2 S bytes) sit in Bm: the factor L is not read again once Linv exists.
        double* tmpL = reinterpret_cast<double*>(p->b8);
        if (p->b8_bytes < (size_t)Mp * Mp * sizeof(double)) {
            if (!p->tmpL) {
                GB_CUDA(ctx, gb_dev_malloc(ctx, (void**)&p->tmpL, (size_t)Mp * Mp * sizeof(double)));
                p->bytes += (size_t)Mp * Mp * 8;
            }
            tmpL = p->tmpL;
        }
        uint8_t* l8 = reinterpret_cast<uint8_t*>(p->Bm);
        GB_CUDA(ctx, chol_inverse(p->Bm, Mp, (int)Mp, w, p->Linv, tmpL, s, &p->nlaunch));
        // u = Linv y (:105), u.u for logl, alpha = L^-T u -- then refined against the fp64 matrix-free operator
        // (the factor came from digit-rounded operands)
        double* rt0 = p->rf_t + 2 * Mp;
        GB_CUDA(ctx, refine_apply_inverse(p->Linv, Mp, p->ydev, rt0, p->alpha, 0, s));
        GB_CUDA(ctx, refine_dot(rt0, rt0, M, p->scal + 1, s));
        RefineArgs ra;
        ra.A[0] = p->A[0]; ra.A[1] = p->A[1]; ra.tables = p->tables; ra.drill = p->drill_dev; ra.partial = p->rf_part;
        ra.Ns = Ns; ra.N = p->N; ra.lda = p->lda; ra.Kp = p->Kp; ra.ext = p->ext; ra.C0 = p->C0; ra.nd = p->nd;
        ra.c0 = p->c0; ra.ncol = ncol; ra.ncp = ncp;
        ra.n[0] = (int)p->n[0]; ra.n[1] = (int)p->n[1]; ra.n[2] = (int)p->n[2];
        ra.nsplit = 8;
        ra.nprop = nrp;
        double* rt = p->rf_t;           // t = A3 K A3^T alpha
        double* rr = p->rf_t + Mp;      // residual
        double* rtmp = p->rf_t + 2 * Mp;
        const int nref = h->refine < 0 ? 0 : h->refine;
        for (int itr = 0; itr <= nref; ++itr) {
            if (!p->lean) {
                GB_CUDA(ctx, refine_at_alpha(ra, p->alpha, p->rf_w, s));   // w = A3^T alpha
            } else {
                // every rank regenerates only the chunks of its own voxel columns; the other entries of w = A3^T alpha come from the
                // other ranks through one all-reduce (each entry has exactly one non-zero contribution: exact in any order)
                if (ctx->nranks > 1) GB_CUDA(ctx, cudaMemsetAsync(ra.partial, 0, (size_t)16 * p->Kp * sizeof(double), s));
                GB_TRY(lean_for_each_chunk(p, [&](int64_t j0, int64_t ncols) -> cudaError_t {
                    const int64_t a0 = ctx->nranks > 1 ? std::max<int64_t>(j0, p->c0) : j0;
                    const int64_t a1 = ctx->nranks > 1 ? std::min<int64_t>(j0 + ncols, p->c1) : j0 + ncols;
                    if (a1 <= a0) return cudaSuccess;
                    return refine_at_alpha_chunk(ra, p->Achunk[0] + (a0 - j0), p->Achunk[1] + (a0 - j0), p->chunk_ld, a0, a1 - a0, p->alpha, s);
                }, ctx->nranks > 1 ? p->c0 : 0, ctx->nranks > 1 ? p->c1 : -1));
                GB_CUDA(ctx, refine_at_alpha_finish(ra, p->alpha, p->rf_w, s));
                if (ctx->nranks > 1) {
                    if (p->nd) GB_CUDA(ctx, cudaMemsetAsync(p->rf_w + 2 * p->Kp, 0, (size_t)p->Kp * sizeof(double), s));   // the drill block is scattered below
                    GB_TRY(comm_allreduce_sum_f64(ctx, p->rf_w, (size_t)2 * p->Kp));
                    if (p->nd) GB_CUDA(ctx, refine_at_alpha_finish_drill(ra, p->alpha, p->rf_w, s));
                }
            }
            if (structured) {                                              // z = K w   (this rank's voxel columns)
                for (int c = 0; c < nrp; ++c)                              // fixed order c = 0, 1, 2: deterministic sums (c = 2: zero weights without drill data)
                    GB_CUDA(ctx, apply_structured(c * 3, p->rf_w + (long)c * p->Kp, p->Kp, 1, p->rf_z, 0, c > 0));
                p->nlaunch += 2 + (p->nd ? 1 : 0);
            } else {
                GB_CUDA(ctx, refine_kw(ra, p->rf_w, p->rf_z, s));
                p->nlaunch += 3 + (refine_kw_slices(ncol) > 1 ? 1 : 0) + (p->nd ? 1 : 0);
            }
            if (itr == nref) break;                                        // z = K A3^T alpha = posterior mean
            if (!p->lean) {
                GB_CUDA(ctx, refine_a_z(ra, p->rf_z, rt, s));              // t = A3 z  (partial over this rank's columns)
            } else {
                bool first = true;
                GB_TRY(lean_for_each_chunk(p, [&](int64_t j0, int64_t ncols) -> cudaError_t {
                    const int64_t ja = std::max<int64_t>(j0, p->c0), jb = std::min<int64_t>(j0 + ncols, p->c1);
                    if (jb <= ja) return cudaSuccess;
                    cudaError_t e = refine_a_z_chunk(ra, p->Achunk[0], p->Achunk[1], p->chunk_ld, j0, ja, jb, p->rf_z, rt, first ? 0 : 1, s);
                    first = false;
                    return e;
                }, p->c0, p->c1));
                GB_CUDA(ctx, refine_a_z_drill(ra, p->rf_z, rt, s));
            }
            GB_TRY(comm_allreduce_sum_f64(ctx, rt, (size_t)Mp));
            GB_CUDA(ctx, refine_residual(p->ydev, rt, p->alpha, Ns, M, Mp, h->gp_sigma, rr, s));
            GB_CUDA(ctx, refine_apply_inverse(p->Linv, Mp, rr, rtmp, p->alpha, 1, s));   // alpha += L^-T L^-1 r
            p->nlaunch += 4 + (p->nd ? 1 : 0);
        }
        GB_CUDA(ctx, refine_scatter_mu(p->rf_z, ncp, ncol, p->mu, s, nrp));
        if (nref > 0) GB_CUDA(ctx, refine_dot(p->ydev, p->alpha, M, p->scal + 1, s));    // u.u = y^T (AkA)^-1 y with the refined alpha
        GB_CUDA(ctx, ozaki_slice_rows(p->Linv, Mp, Mp, Mp, S, p->l_exp, l8, Mp, 128, s));
        {
            // column chunks of Pt that fit the digit scratch (multiples of the column-tile width; one chunk for small problems)
            const long ntc = ozaki_tile_n(S);
            long qn = (long)(p->b8_bytes / ((size_t)Mp * S)) / ntc * ntc;
            if (qn < ntc) qn = ntc;
            for (long q0 = 0; q0 < ldp; q0 += qn) {
                const long qc = ldp - q0 < qn ? ldp - q0 : qn;
                GB_CUDA(ctx, ozaki_slice_cols_mean(p->Pt + q0, Mp, qc, ldp, S, p->b_exp + q0, p->b8, p->alpha, nullptr, ncp, ncol, s));
                GB_CUDA(ctx, ozaki_colsumsq_tri(l8, p->l_exp, p->b8, p->b_exp + q0, (int)Mp, qc, S, p->partial + q0, p->vscratch, ctx->sm_count, s, ldp));
                p->nlaunch += 3;
            }
            p->nlaunch -= 3;     // the first chunk is part of the fixed count below
        }
        GB_CUDA(ctx, cudaEventRecord(p->ev[7], s));
        GB_CUDA(ctx, ozaki_var_finalize(p->partial, (int)(Mp / 128), ldp, ncp, ncol, h->gp_amp, p->var, s, nrp));
        p->nlaunch += 11;                 // u / alpha (2), two dots, mean scatter, 2 x 2 slicing kernels, colsumsq GEMM, variance
    } else if (full) {
        // ---- V = L^-1 Pt (:114) in place, then mean (:115) and variance diagonal (:117)
        GB_CUDA(ctx, chol_forward_solve(p->Bm, Mp, (int)Mp, w, p->Pt, ldp, (int)ldp, p->tmp, s));
        GB_CUDA(ctx, cudaEventRecord(p->ev[7], s));
        dim3 grid((unsigned)((ncp + 255) / 256), (unsigned)nrp);
        mean_var_kernel<<<grid, 256, 0, s>>>(p->Pt, ldp, ncp, ncol, M, p->ysol, h->gp_amp, p->mu, p->var);
        GB_CUDA(ctx, cudaGetLastError());
        p->nlaunch += 2 * (Mp / 128) + 1;
    } else {
        GB_CUDA(ctx, cudaEventRecord(p->ev[7], s));
    }
    GB_CUDA(ctx, cudaEventRecord(p->ev[8], s));
    return GB_OK;
}

static int collect_timings(gb_problem* p) {
    gb_ctx* ctx = p->ctx;
    const int map[8] = {GB_T_TABLES, GB_T_PROJECT, GB_T_DRILLROWS, GB_T_AKA, GB_T_ALLREDUCE, GB_T_CHOL, GB_T_TRSM, GB_T_MEANVAR};
    for (int i = 0; i < 8; ++i) {
        float t = 0.f;
        GB_CUDA(ctx, cudaEventElapsedTime(&t, p->ev[i], p->ev[i + 1]));
        p->ms[map[i]] = t;
    }
    float t = 0.f;
    GB_CUDA(ctx, cudaEventElapsedTime(&t, p->ev[0], p->ev[8]));
    p->ms[GB_T_TOTAL] = t;
    p->ms[GB_T_KSTEPS] = 1.0;
    if (p->sync_ctr && p->ksteps_total > 0.0) {      // the stream is idle here (the callers synchronise before collecting)
        unsigned long long done = 0;
        GB_CUDA(ctx, cudaMemcpy(&done, p->sync_ctr + 2, sizeof done, cudaMemcpyDeviceToHost));
        if (done > 0) p->ms[GB_T_KSTEPS] = (double)done / p->ksteps_total;
    }
    {
        // kernels launched by run_predict: set_y, tables, projection (+ digit slicing), [drill rows x2], aka, noise diag,
        // Cholesky (potrf + panel + trailing per block), u solve (2 GEMMs per block), dot; the mean / variance
        // stage adds its own count (p->nlaunch) in run_predict
        const long nblk = p->Mp / 128;
        p->ms[GB_T_LAUNCHES] = (double)(4 + (p->nd ? 2 : 0) + (3 * nblk - 2) + p->nlaunch);
    }
    return GB_OK;
}

extern "C" int gb_predict(gb_problem* p, const gb_hyper* h, int flags, double* mu, double* var, double* logl, int* info) {
    if (!p || !h) return GB_ERR_ARG;
    gb_ctx* ctx = p->ctx;
    const bool full = (flags & (GB_FLAG_MEAN | GB_FLAG_VAR)) != 0;
    GB_TRY(run_predict(p, h, full));
    cudaStream_t s = ctx->stream;
    const long ncol = p->ncol;
    double* o = p->out_pinned;
    GB_CUDA(ctx, cudaEventRecord(p->ev[9], s));
    if (full && mu) GB_CUDA(ctx, cudaMemcpyAsync(o, p->mu, 3 * ncol * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (full && var) GB_CUDA(ctx, cudaMemcpyAsync(o + 3 * ncol, p->var, 3 * ncol * sizeof(double), cudaMemcpyDeviceToHost, s));
    GB_CUDA(ctx, cudaMemcpyAsync(o + 6 * ncol, p->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    GB_CUDA(ctx, cudaMemcpyAsync(o + 6 * ncol + 2, p->info, sizeof(int), cudaMemcpyDeviceToHost, s));
    GB_CUDA(ctx, cudaEventRecord(p->ev[10], s));
    GB_CUDA(ctx, cudaStreamSynchronize(s));
    GB_TRY(collect_timings(p));
    {
        float t = 0.f;
        GB_CUDA(ctx, cudaEventElapsedTime(&t, p->ev[9], p->ev[10]));
        p->ms[GB_T_D2H] = t;
    }
    int inf = 0;
    memcpy(&inf, o + 6 * ncol + 2, sizeof(int));
    if (info) *info = inf;
    if (full && mu) memcpy(mu, o, 3 * ncol * sizeof(double));
    if (full && var) memcpy(var, o + 3 * ncol, 3 * ncol * sizeof(double));
    if (logl) {
        // inversion.py:107-110 (Q5: N_vox log 2 pi, not M)
        const double logdet = o[6 * ncol], uu = o[6 * ncol + 1];
        *logl = -0.5 * (uu + logdet + (double)p->N * log(2.0 * M_PI));
    }
    return inf > 0 ? inf : GB_OK;
}

extern "C" int gb_neg_logl(gb_problem* p, const gb_hyper* h, double* neg_logl, int* info) {
    if (!p || !h || !neg_logl) return GB_ERR_ARG;
    gb_ctx* ctx = p->ctx;
    GB_TRY(run_predict(p, h, false));
    double sc[2];
    int inf = 0;
    GB_CUDA(ctx, cudaMemcpyAsync(p->out_pinned, p->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(p->out_pinned + 2, p->info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GB_TRY(collect_timings(p));
    memcpy(sc, p->out_pinned, sizeof sc);
    memcpy(&inf, p->out_pinned + 2, sizeof(int));
    if (info) *info = inf;
    const double logl = -0.5 * (sc[1] + sc[0]);   // inversion.py:148 (no N log 2 pi)
    *neg_logl = (inf > 0 || !isfinite(logl)) ? INFINITY : -logl;   // :150-152
    return GB_OK;
}

// y (+)= A[:, j0 : j0 + ncols) x[j0 : j0 + ncols)   (one warp per row; lean mode of gb_forward)
__global__ void gemv_chunk_kernel(const double* __restrict__ A, long rows, long cols, long ld, const double* __restrict__ x,
                                  double* __restrict__ y, int accumulate) {
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (long c = lane; c < cols; c += 32) acc = fma(A[row * ld + c], x[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[row] = accumulate ? y[row] + acc : acc;
}

extern "C" int gb_problem_get_sens(gb_problem* p, int kind, double* out) {
    if (!p || !out || (kind != GB_SENS_GRAV && kind != GB_SENS_MAGN)) return GB_ERR_ARG;
    gb_ctx* ctx = p->ctx;
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (p->lean) {
        GB_TRY(lean_for_each_chunk(p, [&](int64_t j0, int64_t ncols) -> cudaError_t {
            cudaError_t e = cudaMemcpy2DAsync(out + j0, p->N * sizeof(double), p->Achunk[kind], p->chunk_ld * sizeof(double), ncols * sizeof(double),
                                              p->Ns, cudaMemcpyDeviceToHost, ctx->stream);
            return e != cudaSuccess ? e : cudaStreamSynchronize(ctx->stream);
        }));
        return GB_OK;
    }
    GB_CUDA(ctx, cudaMemcpy2DAsync(out, p->N * sizeof(double), p->A[kind], p->lda * sizeof(double), p->N * sizeof(double), p->Ns,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

extern "C" int gb_forward(gb_problem* p, int kind, const double* x, double* out) {
    if (!p || !x || !out || (kind != GB_SENS_GRAV && kind != GB_SENS_MAGN)) return GB_ERR_ARG;
    gb_ctx* ctx = p->ctx;
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> xd, yd;
    GB_CUDA(ctx, xd.alloc(p->N));
    GB_CUDA(ctx, yd.alloc(p->Ns));
    GB_CUDA(ctx, cudaMemcpyAsync(xd.p, x, p->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (p->lean) {
        GB_TRY(lean_for_each_chunk(p, [&](int64_t j0, int64_t ncols) -> cudaError_t {
            gemv_chunk_kernel<<<(unsigned)((p->Ns + 7) / 8), 256, 0, ctx->stream>>>(p->Achunk[kind], p->Ns, ncols, p->chunk_ld, xd.p + j0, yd.p, j0 > 0);
            return cudaGetLastError();
        }));
    } else {
        GB_CUDA(ctx, launch_gemv(p->A[kind], p->Ns, p->N, p->lda, xd.p, yd.p, ctx->stream));
    }
    GB_CUDA(ctx, cudaMemcpyAsync(out, yd.p, p->Ns * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

// Dense posterior covariance  K - V^T V  (inversion.py:117) for callers that read Inversion.cov_rec; small cubes only.
__global__ void posterior_cov_kernel(const double* __restrict__ tables, long ext, long C0, const int* __restrict__ L,
                                     const double* __restrict__ V, long ldp, long ncp, long N, long M, double* __restrict__ out) {
    __shared__ double va[16][17], vb[16][17];
    const long a = (long)blockIdx.y * 16 + threadIdx.y, b = (long)blockIdx.x * 16 + threadIdx.x;   // rows / cols of the 3N x 3N output
    const long a_load = (long)blockIdx.y * 16 + threadIdx.x, b_load = b;
    const long n3 = 3 * N;
    double acc = 0.0;
    for (long m0 = 0; m0 < M; m0 += 16) {
        const long m = m0 + threadIdx.y;
        va[threadIdx.y][threadIdx.x] = (m < M && a_load < n3) ? V[m * ldp + (a_load / N) * ncp + a_load % N] : 0.0;
        vb[threadIdx.y][threadIdx.x] = (m < M && b_load < n3) ? V[m * ldp + (b_load / N) * ncp + b_load % N] : 0.0;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) acc = fma(va[q][threadIdx.y], vb[q][threadIdx.x], acc);
        __syncthreads();
    }
    if (a < n3 && b < n3) {
        const int ra = (int)(a / N), rb = (int)(b / N);
        const double k = tables[(long)(ra * 3 + rb) * ext + C0 + (L[b % N] - L[a % N])];
        out[a * n3 + b] = k - acc;
    }
}

extern "C" int gb_posterior_cov(gb_problem* p, const gb_hyper* h, double* out) {
    if (!p || !h || !out) return GB_ERR_ARG;
    gb_ctx* ctx = p->ctx;
    if (p->ncol != p->N) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "gb_posterior_cov needs an unsharded problem");
    if (3 * p->N > 46000) return gb_fail(ctx, GB_ERR_UNSUPPORTED, "dense 3N x 3N posterior covariance refused for N=%lld", (long long)p->N);
    gb_hyper h64 = *h;
    h64.slices = 0;                       // the dense posterior covariance is built from V = L^-1 Pt, which only the fp64 path stores
    GB_TRY(run_predict(p, &h64, true, /*all_blocks=*/true));
    const long n3 = 3 * p->N;
    DevBuf<double> o;
    GB_CUDA(ctx, o.alloc((size_t)n3 * n3));
    dim3 grid((unsigned)((n3 + 15) / 16), (unsigned)((n3 + 15) / 16)), block(16, 16);
    posterior_cov_kernel<<<grid, block, 0, ctx->stream>>>(p->tables, p->ext, p->C0, p->L, p->Pt, p->ldp, p->ncp, p->N, p->M, o.p);
    GB_CUDA(ctx, cudaGetLastError());
    GB_CUDA(ctx, cudaMemcpyAsync(out, o.p, (size_t)n3 * n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int inf = 0;
    GB_CUDA(ctx, cudaMemcpy(&inf, p->info, sizeof(int), cudaMemcpyDeviceToHost));
    return inf > 0 ? inf : GB_OK;
}

extern "C" int gb_get_timings(gb_problem* p, double* ms, int n) {
    if (!p || !ms) return GB_ERR_ARG;
    for (int i = 0; i < n && i < GB_NUM_TIMERS; ++i) ms[i] = p->ms[i];
    return GB_OK;
}
