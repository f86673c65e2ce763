// Shared declarations of libgeobo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/geobo_b200.h"
#include "formulas.cuh"

struct gb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // NCCL (resolved lazily with dlopen, see comm.cu)
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    // size-keyed cache of device / pinned-host buffers released by destroyed problems: repeated Inversion.cubing()
    // calls on the same cube reuse them instead of paying cudaMalloc / cudaFree of several GB per call
    std::multimap<size_t, void*> dev_cache, host_cache;
    std::unordered_map<void*, size_t> dev_sizes, host_sizes;
};

// cached allocators (api.cu); the cache is emptied on allocation failure and by gb_ctx_release_cache / gb_ctx_destroy
cudaError_t gb_dev_malloc(gb_ctx* ctx, void** p, size_t bytes);
void gb_dev_free(gb_ctx* ctx, void* p);
cudaError_t gb_host_malloc(gb_ctx* ctx, void** p, size_t bytes);
void gb_host_free(gb_ctx* ctx, void* p);

extern std::string g_gb_create_error;

inline int gb_fail(gb_ctx* ctx, int code, const char* fmt, ...) __attribute__((format(printf, 3, 4)));
inline int gb_fail(gb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_gb_create_error = buf;
    return code;
}

#define GB_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return gb_fail(ctx, e__ == cudaErrorMemoryAllocation ? GB_ERR_NOMEM : GB_ERR_CUDA,      \
                           "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));     \
    } while (0)

#define GB_TRY(call)                  \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != GB_OK) return rc__; \
    } while (0)

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// RAII device buffer used for scratch inside one API call
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        n = count;
        return cudaMalloc((void**)&p, count * sizeof(T) + 16);
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

// ---------------------------------------------------------------------------- covariance functions (cov.cu)
// CovParams: formulas.cuh

// ---------------------------------------------------------------------------- fp64 tensor-pipe GEMM engine (gemm_f64.cu)
namespace gemm {
enum BMode { B_T = 0, B_N = 1, B_GEN = 2 };
struct Task {
    const double* A;      // [M][lda], K contiguous
    const double* B;      // B_T: [N][ldb] (K contiguous); B_N: [K][ldb] (N contiguous); B_GEN: table base (+C0)
    const double* Cin;    // used when beta != 0 (may alias C)
    double* C;            // [M][ldc]
    long lda, ldb, ldcin, ldc;
    int M, N, K;          // K must be a multiple of 16 (operands zero padded)
    int lower;            // skip tiles strictly above the diagonal
    double alpha, beta;   // C = alpha * A.B + beta * Cin
    const int* Lrow;      // B_GEN: extended-lattice id of output column n  (Lrow[n])
    const int* Lcol;      // B_GEN: extended-lattice id of contraction index k (Lcol[k])
};
constexpr int MAX_TASKS = 32;
struct TaskBatch {
    Task t[MAX_TASKS];
    int n;
};
// Launches one grid covering all tasks in the batch (blockIdx.z = task).
cudaError_t launch(const TaskBatch& batch, BMode mode, cudaStream_t stream);
cudaError_t init();
}  // namespace gemm

// ---------------------------------------------------------------------------- kernel launchers
cudaError_t launch_cov_tables(const CovParams& cp, const int64_t ncube[3], const double vox[3], double* tables /*9 x ext*/,
                              cudaStream_t s, int* nonfinite = nullptr /* device flag set to 1 if any table value is NaN / inf */);
cudaError_t launch_lattice_ids(const int64_t ncube[3], int* L, int64_t n_padded, cudaStream_t s);
cudaError_t launch_create_cov_dense(const CovParams& cp, const double* D2, int64_t n, double* out, cudaStream_t s);
// tables ([9][ext]) and L ([N]) given: stationary tables + gather (HBM-write bound); null: per-element evaluation
cudaError_t launch_create_cov_grid(const CovParams& cp, const int64_t ncube[3], const double vox[3], double* out,
                                   cudaStream_t s, double* tables = nullptr, int* L = nullptr);
cudaError_t launch_grid_points(const int64_t lpix[3], const double sc[3], double* out, cudaStream_t s);
cudaError_t launch_sqdist(const double* pts, int64_t n, int dim, double* out, cudaStream_t s);
cudaError_t launch_a_sens(int kind, const double B[3], const double* loc, int64_t nsens, const double* edges,
                          const int64_t ncube[3], double mul, double div, double* out, int64_t ld, int sm_count,
                          cudaStream_t s);
cudaError_t launch_a_sens_range(int kind, const double B[3], const double* loc, int64_t nsens, const double* edges,
                                const int64_t ncube[3], double mul, double div, double* out, int64_t ld, int iy_begin, int iy_end,
                                int sm_count, cudaStream_t s);
cudaError_t launch_cov_function(int kernel_id, int cross, const double* D2, int64_t count, double l1, double l2, double* out,
                                cudaStream_t s);
cudaError_t launch_corner_func(int kind, const double* x, const double* y, const double* z, int64_t count, const double B[3],
                               double* out, cudaStream_t s);
cudaError_t launch_a_drill(const double* loc, int64_t nd, const double* vp, int64_t N, double* out, cudaStream_t s);
cudaError_t launch_gemv(const double* A, int64_t rows, int64_t cols, int64_t ld, const double* x, double* y,
                        cudaStream_t s);

// int8 digit-slice projection on tcgen05 (ozaki.cu)
struct OzakiArgs {
    const uint8_t* a8[2];    // pre-tiled digit blocks of A_grav / A_magn: [row tile][k step][slice][4096 B]
    const int* a_exp[2];     // per-sensor-row exponents
    const uint8_t* t8;       // digit planes of the 9 covariance tables
    const int* t_exp;        // per-table exponents
    const int* L;            // extended-lattice ids [kp]
    double* Pt;
    long ext, C0, kp, ldp, ncp;
    int Ns, ncol, c0;
    int nr;                  // property blocks computed per data block: 3, or 2 when there is no drill data (block 2 is NaN in the reference)
    const int* cull;         // [9][2] device: extents (|dy|, |dx|) of the non-zero digits of every table (ozaki_table_extents), or null
    int n[3];                // xN, yN, zN
    unsigned long long* steps_ctr;   // device counter, ZERO at launch: K steps visited (sum over tiles); or null
    int ks_base, cy0, cy1, accumulate;   // streamed contraction (see ozaki::Params); all 0 for a launch over the whole contraction
    int sync_slack;          // rounds of slack of the pacing (0 = strict)
    int* perm_scratch;       // [2][6][ceil(ncol / 128)] ints (tile order sorted by K-step count + keys), or null = natural tile order
    unsigned int* sync_ctr;  // device counter, ZERO at launch: tile-round pacing of the copy lanes (keeps the shared K strips in L2); or null
};
// cull[2 tb + 0 / 1] = largest |dy| / |dx| lattice offset with a non-zero digit in any plane of table tb (after ozaki_slice_tables)
cudaError_t ozaki_table_extents(const uint8_t* t8, long ext, int slices, const int64_t n[3], int* cull /*[18]*/, cudaStream_t s);
int ozaki_tile_n(int slices);
int ozaki_tile_np(int slices);
int ozaki_chunk();
long ozaki_table_bytes(long ext, int slices);
long ozaki_rows_bytes(long rows, long kp, int slices, int tr = 128);
cudaError_t ozaki_slice_rows(const double* A, long rows, long cols, long ld, int slices, int* exps, uint8_t* out, long kp, int tr,
                             cudaStream_t s);
cudaError_t ozaki_slice_sens(const double* A, long rows, long cols, long ld, int slices, int* exps, uint8_t* out, long kp,
                             cudaStream_t s);
// the same digit blocks built from column chunks of the operand (sensitivities that are regenerated chunk by chunk instead of
// being resident): running row maxima over the chunks -> exponents -> digits of one chunk into k steps [ks0, ks0 + cols / 32)
cudaError_t ozaki_row_absmax_accum(const double* A, long rows, long cols, long ld, double* amax /*[rows], in-out*/, cudaStream_t s);
cudaError_t ozaki_exps_from_absmax(const double* amax, long rows, int* exps, cudaStream_t s);
cudaError_t ozaki_slice_rows_range(const double* A, long rows, long cols, long ld, int slices, const int* exps, uint8_t* out, long kp_total,
                                   int tr, long ks0, cudaStream_t s);
cudaError_t ozaki_slice_tables(const double* tables, long ext, int slices, int* exps, uint8_t* out, cudaStream_t s);
cudaError_t ozaki_project(const OzakiArgs& a, int slices, int sm_count, cudaStream_t s);
// int8 digit-slice GEMM with both operands in memory (ozaki_gemm.cu)
long ozaki_cols_bytes(long cols, long kp, int slices);
cudaError_t ozaki_slice_cols_mean(const double* X, long rows, long cols, long ld, int slices, int* exps, uint8_t* out,
                                  const double* alpha, double* mu, long ncp, long ncol, cudaStream_t s);
long ozaki_colsumsq_scratch_bytes(int Mp, int slices, int sm_count);
// ldpart: leading dimension of `partial` ([Mp / 128][ldpart]; 0 = ncols) -- a column chunk of a wider matrix writes into its columns
cudaError_t ozaki_colsumsq_tri(const uint8_t* a8, const int* a_exp, const uint8_t* b8, const int* b_exp, int Mp, long ncols,
                               int slices, double* partial, double* scratch, int sm_count, cudaStream_t s, long ldpart = 0);
cudaError_t ozaki_gemm_store(const uint8_t* a8, const int* a_exp, int a_ksteps, int a_k0, const uint8_t* b8, const int* b_exp,
                             int b_ksteps, int b_k0, int ksteps, int M, int N, double* C, long ldc, int lower, int slices,
                             int sm_count, cudaStream_t s);
cudaError_t ozaki_var_finalize(const double* partial, int n_mtile, long ldpart, long ncp, long ncol, double amp, double* var,
                               cudaStream_t s, int nr = 3);

// fp64 matrix-free operator pieces for the iterative refinement of alpha and the posterior mean (refine.cu)
struct RefineArgs {
    const double* A[2];      // [Ns][lda]
    const double* tables;    // [9][ext]
    const int64_t* drill;    // [nd] (device)
    double* partial;         // [nsplit][2][Kp] scratch of A3^T alpha
    long Ns, N, lda, Kp, ext, C0, nd, c0, ncol, ncp;
    int n[3];
    int nsplit;
    int nprop;               // 3, or 2 without drill data: neither the drill data block (c = 2, its weights are zero) nor the drill
                             // property block (r = 2, NaN in the reference) enters K.w
};
cudaError_t refine_at_alpha(const RefineArgs& a, const double* alpha, double* w /*[3][Kp]*/, cudaStream_t s);
// the same in column chunks (sensitivities regenerated chunk by chunk): partial sums of the chunk's columns [j0, j0 + ncols), then one finish
cudaError_t refine_at_alpha_chunk(const RefineArgs& a, const double* A0, const double* A1, long ld, long j0, long ncols, const double* alpha,
                                  cudaStream_t s);
cudaError_t refine_at_alpha_finish(const RefineArgs& a, const double* alpha, double* w, cudaStream_t s);
cudaError_t refine_at_alpha_finish_drill(const RefineArgs& a, const double* alpha, double* w, cudaStream_t s);
// t (+)= A_c[:, ja : jb) z[c][ja - c0 : jb - c0)  for a column chunk held at A0 / A1 (column ja of the cube = column ja - j0 of the chunk)
cudaError_t refine_a_z_chunk(const RefineArgs& a, const double* A0, const double* A1, long ld, long j0, long ja, long jb, const double* z,
                             double* t, int accumulate, cudaStream_t s);
cudaError_t refine_a_z_drill(const RefineArgs& a, const double* z, double* t, cudaStream_t s);
int refine_kw_slices(long ncol);
cudaError_t refine_kw(const RefineArgs& a, const double* w, double* z /*[refine_kw_slices][3][ncp]*/, cudaStream_t s);
cudaError_t refine_a_z(const RefineArgs& a, const double* z, double* t /*[Mp]*/, cudaStream_t s);
cudaError_t refine_residual(const double* y, const double* t, const double* alpha, long Ns, long M, long Mp, const double sigma[3],
                            double* r, cudaStream_t s);
cudaError_t refine_apply_inverse(const double* Linv, long Mp, const double* x, double* tmp, double* out, int accumulate, cudaStream_t s);
cudaError_t refine_linv_t(const double* Linv, long Mp, const double* x, int xs, double* out, cudaStream_t s);
cudaError_t refine_dot(const double* a, const double* b, long n, double* out, cudaStream_t s);
cudaError_t refine_scatter_mu(const double* z, long ncp, long ncol, double* mu, cudaStream_t s, int nr = 3);

// Kronecker-structured products with the exp covariance blocks (kron.cu; opt-in, gb_hyper.structure = GB_STRUCTURE_KRON)
struct KronGeom;
long kron_factor_doubles(const KronGeom& g);
long kron_scratch_doubles(const KronGeom& g, long rows);
int kron_supported(const KronGeom& g, char* why, size_t len);
cudaError_t kron_build_factors(const double* tables, long ext, long C0, const KronGeom& g, double* kf /*[9][3][FL]*/, cudaStream_t s);
// out[s][r * r_stride_out + (j - c0)] (+)= sum_i A[s][i] * K_(blk0 + r)[i][j]  for rows s < nrows, r = 0..2, voxel columns j of the shard
cudaError_t kron_apply(const KronGeom& g, const double* kf, int blk0, const double* A, long lda, long nrows, double* T, long T_doubles,
                       double* out, long ldo, long r_stride_out, int accumulate, cudaStream_t s, long* nlaunch, int nr = 3);

// Compact-support (stencil) products with the 'sparse' covariance blocks (stencil.cu; opt-in, GB_STRUCTURE_COMPACT)
struct StencilGeom;
cudaError_t stencil_apply(const StencilGeom& g, const double* tab0 /* tables + blk0 * ext + C0 */, const double* A, long lda, long nrows,
                          double* out, long ldo, long r_stride_out, int accumulate, cudaStream_t s, long* nlaunch, int nr = 3);

// Block-Toeplitz (FFT) products with the stationary covariance blocks (fftconv.cu; opt-in, GB_STRUCTURE_FFT, any kernel)
struct FftGeom;
struct cplx;
int fft_supported(const FftGeom& g, char* why, size_t len);
long fft_chunk_pairs(const FftGeom& g, long rows);
long fft_scratch_cplx(const FftGeom& g, long B);
cudaError_t fft_build_twiddles(const FftGeom& g, cplx* tw /*[3][FFT_MAXP / 2]: y, x, z*/, cudaStream_t s);
cudaError_t fft_build_spectra(const FftGeom& g, const double* tables, long ext, long C0, const cplx* tw, cplx* X, cplx* Y, double* W /*[9][P3]*/,
                              cudaStream_t s, long* nlaunch);
cudaError_t fft_apply(const FftGeom& g, const double* W, const cplx* tw, int blk0, const double* A, long lda, long nrows, cplx* scratch, long B,
                      double* out, long ldo, long r_stride_out, int accumulate, cudaStream_t s, long* nlaunch, int nr = 3);

// Cholesky / triangular solve (chol.cu)
struct CholWork {
    double* linv = nullptr;   // [Mp/128][128][128] inverses of the diagonal blocks
    double* logdet = nullptr; // device scalar: sum log(L_ii^2)
    int* info = nullptr;      // device scalar: first non-positive pivot (1-based), 0 = ok
};
cudaError_t chol_factor(double* Bm, long ldb, int Mp, int Mtrue, const CholWork& w, cudaStream_t s);
// 1-D block-cyclic factorisation over the ranks of the context (one ncclBroadcast per 128-column panel); L ends up replicated
int chol_factor_dist(gb_ctx* ctx, double* Bm, long ldb, int Mp, int Mtrue, const CholWork& w, double* stage, double* pan, int* paninfo,
                     cudaStream_t s);
// V = L^-1 Pt in place (Pt: [Mp][ldp]); tmp: [128][ldp]
// Linv = L^-1 (lower triangular, Mp x Mp, ld = Mp) by recursive doubling from the 128 x 128 diagonal-block inverses that
// chol_factor left in w.linv: [L11 0; L21 L22]^-1 = [X11 0; -X22 L21 X11, X22]; two batched GEMM launches per level.
// Ltmp: Mp x Mp scratch.
cudaError_t chol_inverse(const double* L, long ldl, int Mp, const CholWork& w, double* Linv, double* Ltmp, cudaStream_t s, long* nlaunch);
// tri != 0: the right-hand side is lower triangular (e.g. the identity): block row kb only has columns < 128 (kb + 1)
cudaError_t chol_forward_solve(const double* L, long ldl, int Mp, const CholWork& w, double* Pt, long ldp, int ncols,
                               double* tmp, cudaStream_t s, int tri = 0);
