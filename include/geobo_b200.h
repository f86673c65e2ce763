/*
 * geobo_b200 -- C ABI of the B200-native GeoBO joint-inversion hot path.
 *
 * The reference (sebhaan/geobo) is pure Python with no FFI of its own; its boundary for
 * this path is the Python module surface geobo.kernels / geobo.sensormodel /
 * geobo.inversion.  Each entry point below names the reference interface it replaces
 * (file:line in the reference tree).  INTEGRATION.md shows the ctypes stubs a maintainer
 * of the reference would add to call them.
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = error (text via gb_last_error),
 *     >0 only from gb_predict/gb_neg_logl = LAPACK-style index (1-based) of the first
 *     non-positive pivot of the Cholesky of A K A^T + Sigma;
 *   - plain pointers and sizes only; all arrays are C-contiguous float64 / int64 unless
 *     noted; the caller owns every host buffer, the library owns all device memory;
 *   - one calling thread per gb_ctx; no callbacks, no exceptions across the boundary;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *     with GB_ERR_CUDA.
 */
#ifndef GEOBO_B200_H
#define GEOBO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_OK 0
#define GB_ERR_ARG (-1)
#define GB_ERR_CUDA (-2)
#define GB_ERR_NCCL (-3)
#define GB_ERR_NOMEM (-4)
#define GB_ERR_UNSUPPORTED (-5)

/* kernel families, geobo/kernels.py:183-194 ('sparse' is the reference default) */
#define GB_KERNEL_SPARSE 0
#define GB_KERNEL_EXP 1
#define GB_KERNEL_MATERN32 2

/* gb_hyper.structure */
#define GB_STRUCTURE_DENSE 0
#define GB_STRUCTURE_KRON 1
#define GB_STRUCTURE_COMPACT 2
#define GB_STRUCTURE_FFT 3

/* forward-model kinds, geobo/sensormodel.py:71-74 */
#define GB_SENS_GRAV 0
#define GB_SENS_MAGN 1

/* gb_predict flags */
#define GB_FLAG_MEAN 1     /* posterior mean         (inversion.py:115) */
#define GB_FLAG_VAR 2      /* diag of posterior cov. (inversion.py:117,238) */
#define GB_FLAG_LOGL 4     /* log marginal likelihood (inversion.py:107-110) */
#define GB_FLAG_ALL 7

#define GB_NUM_TIMERS 16
/* indices into gb_get_timings() (milliseconds, CUDA events on the context stream) */
#define GB_T_SENS 0        /* A_sens x2 (problem_create)          */
#define GB_T_TABLES 1      /* stationary covariance tables        */
#define GB_T_PROJECT 2     /* Pt = A . K  (fused assembly + GEMM) */
#define GB_T_DRILLROWS 3   /* drill rows of Pt / AkA gathers      */
#define GB_T_AKA 4         /* AkA = A . Pt^T + Sigma (int8 paths: incl. slicing Pt rows) */
#define GB_T_ALLREDUCE 5   /* NCCL all-reduce of AkA (multi-GPU)  */
#define GB_T_CHOL 6        /* Cholesky of AkA                     */
#define GB_T_TRSM 7        /* u = L^-1 y; fp64: V = L^-1 Pt; int8: Linv, refinement of alpha, mean, colsumsq(Linv.Pt) */
#define GB_T_MEANVAR 8     /* fp64: mean + variance diag from V; int8: variance finalisation */
#define GB_T_TOTAL 9       /* whole gb_predict on device          */
#define GB_T_D2H 10        /* result copies to host               */
#define GB_T_KSTEPS 12     /* int8 projection: fraction of the K steps (32 contraction voxels x one tile) actually visited; the others
                              multiply by all-zero covariance digits and are skipped (a ratio, not ms; 1 = every step)  */
#define GB_T_LAUNCHES 11   /* number of library kernels launched by the last gb_predict / gb_neg_logl (a count, not ms) */

typedef struct gb_ctx gb_ctx;
typedef struct gb_problem gb_problem;

/* ------------------------------------------------------------------ context */
int gb_version(void);
/* device < 0: use the current device. */
int gb_ctx_create(int device, gb_ctx** out);
int gb_ctx_destroy(gb_ctx* ctx);
/* Device and pinned-host buffers of destroyed problems are kept in a size-keyed cache on the context so that repeated
 * Inversion.cubing() calls on the same cube do not pay cudaMalloc/cudaFree of several GB each; this returns them to the
 * driver (also done automatically when an allocation fails and by gb_ctx_destroy). */
int gb_ctx_release_cache(gb_ctx* ctx);
/* last error text of this context (ctx == NULL: of the last failed gb_ctx_create). */
const char* gb_last_error(const gb_ctx* ctx);
int gb_device_info(gb_ctx* ctx, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor,
                   uint64_t* free_bytes, uint64_t* total_bytes);

/* Multi-GPU (one process per GPU).  gb_comm_unique_id fills a 128-byte NCCL id on rank 0;
 * the host distributes it (e.g. torch.distributed / a file) and every rank calls gb_comm_init. */
int gb_comm_unique_id(gb_ctx* ctx, void* id128);
int gb_comm_init(gb_ctx* ctx, const void* id128, int rank, int nranks);
/* out[r * count + i] = rank r's local[i] (host buffers, NCCL all-gather over NVLink): gathers the mean / variance shards so
 * that Inversion.cubing returns whole cubes on every rank.  Single rank: a copy. */
int gb_comm_allgather(gb_ctx* ctx, const double* local, int64_t count, double* out);

/* ------------------------------------------------------------------ geobo/kernels.py */
/* kernels.calcGridPoints3D(Lpix, pixscale)  (kernels.py:27-42): out[(iy*xN+ix)*zN+iz][0..2] */
int gb_grid_points(gb_ctx* ctx, const int64_t lpix[3], const double pixscale[3], double* out);
/* kernels.calcDistanceMatrix(points)        (kernels.py:45-61): out[n*n] squared distances */
int gb_sqdist(gb_ctx* ctx, const double* points, int64_t n, int dim, double* out);
/* kernels.create_cov(D2, gplength, crossweights, fkernel) (kernels.py:158-195).
 * gplength must already be de-duplicated (kernels.py:174-180 is host logic because it mutates
 * the caller's ndarray); out is (3n)x(3n). */
int gb_create_cov(gb_ctx* ctx, const double* D2, int64_t n, const double gplength[3], const double crossweights[3],
                  int kernel_id, double* out);
/* kernels.gpkernel / gpkernel_sparse / gpkernel_matern32 (cross = 0, kernels.py:81-88,101-114,140-146) and
 * gpkernel2 / gpkernel_sparse2 / gpkernel_matern32_2 (cross = 1, gammas = (l1, l2), kernels.py:90-99,116-138,148-156)
 * applied elementwise to `count` squared distances. */
int gb_cov_function(gb_ctx* ctx, int kernel_id, int cross, const double* D2, int64_t count, double l1, double l2, double* out);
/* Same matrix generated on the device from the grid spec alone (no D2 input): the HBM-write-bound
 * assembly kernel.  out may be NULL (device only, for timing); ms receives the kernel time. */
int gb_create_cov_grid(gb_ctx* ctx, const int64_t ncube[3], const double voxsize[3], const double gplength[3],
                       const double crossweights[3], double amp, int kernel_id, double* out, float* ms);

/* ------------------------------------------------------------------ geobo/sensormodel.py */
/* sensormodel.A_sens(magneticField, locations, Edges, func) (sensormodel.py:29-93).
 * edges: (3, yN+1, xN+1, zN+1); locations: (nsens, 3); out: (nsens, xN*yN*zN).
 * out = (mul * s) / div with s the 8-corner difference: grav: mul = c_MILLIGALS_UNITS, div = fcor_grav;
 * magn: mul = 1, div = fcor_mag (sensormodel.py:88-91). */
int gb_a_sens(gb_ctx* ctx, int kind, const double B[3], const double* locations, int64_t nsens, const double* edges,
              const int64_t ncube[3], double mul, double div, double* out);

/* sensormodel.grav_func(x, y, z) / magn_func(x, y, z, bx, by, bz) (sensormodel.py:96-133), elementwise. */
int gb_corner_func(gb_ctx* ctx, int kind, const double* x, const double* y, const double* z, int64_t count, const double B[3],
                   double* out);
/* sensormodel.A_drill(loc, voxelpos) (sensormodel.py:136-153): loc (3, ndrill), voxelpos (3, nvox) -> out (ndrill, nvox). */
int gb_a_drill(gb_ctx* ctx, const double* loc, int64_t ndrill, const double* voxelpos, int64_t nvox, double* out);

/* ------------------------------------------------------------------ geobo/inversion.py */
typedef struct gb_problem_desc {
    int64_t ncube[3];          /* xNcube, yNcube, zNcube                                   */
    double voxsize[3];         /* xvoxsize, yvoxsize, zvoxsize (config_loader.py:56-58)    */
    const double* edges;       /* (3, yN+1, xN+1, zN+1)  Inversion.Edges (inversion.py:61-66) */
    const double* locations;   /* (nsens, 3) sensor_locations; nsens must equal xN*yN (sensormodel.py:54,58) */
    int64_t nsens;
    double magnetic_field[3];  /* magneticField (config_loader.py:46)                      */
    double grav_mul, grav_div; /* c_MILLIGALS_UNITS, fcor_grav                             */
    double magn_mul, magn_div; /* 1.0, fcor_mag                                            */
    const int64_t* drill_idx;  /* flat voxel index of every drill row (A_drill one-hot column, sensormodel.py:136-153) */
    int64_t ndrill;
    /* voxel-column shard [col_begin, col_end) of Pt = A.K owned by this rank (0, N on one GPU) */
    int64_t col_begin, col_end;
} gb_problem_desc;

typedef struct gb_hyper {
    double gp_length[3];       /* effective (de-duplicated) length scales                  */
    double gp_sigma[3];        /* noise std per data group (inversion.py:94)               */
    double coeffm[3];          /* cross weights w1,w2,w3 (kernels.py:181)                  */
    double gp_amp;             /* amplitude (inversion.py:92)                              */
    int kernel_id;             /* GB_KERNEL_*                                              */
    int slices;                /* 0: fp64 tensor pipe (DMMA) for the projection; 4, 5, 6: error-free int8 digit
                                  products on tcgen05/TMEM with 7+8(slices-1) bits per operand (needs zNcube % 16 == 0) */
    int refine;                /* slices != 0 only: steps of iterative refinement of (A K A^T + Sigma)^-1 y against the fp64
                                  matrix-free operator before the mean K A3^T alpha is formed (0 = none; 1 is the default
                                  of the Python surface)                                                              */
    int structure;             /* GB_STRUCTURE_DENSE (0, default): the dense contraction Pt = A3 . K the north star names;
                                  GB_STRUCTURE_KRON (1): opt-in fast path for kernel_id = GB_KERNEL_EXP only -- on the regular
                                  voxel grid every exp block of create_cov (kernels.py:81-99) is a Kronecker product
                                  Ky (x) Kx (x) Kz, so Pt = A3 . K and the K . w of the refinement run as three Toeplitz mode
                                  products (2 N (xN + yN + zN) flops per row instead of 2 N^2).  Any other
                                  kernel is refused with GB_ERR_UNSUPPORTED.  Everything downstream is unchanged.
                                  GB_STRUCTURE_COMPACT (2): opt-in fast path for kernel_id = GB_KERNEL_SPARSE only -- the compact
                                  kernels (kernels.py:101-138) vanish beyond their length scale, so the same two products run as
                                  a 3-D tap sum over the offsets inside the support (taps from the stationary tables).
                                  GB_STRUCTURE_FFT (3): opt-in fast path for ANY kernel -- every block of create_cov is block-Toeplitz
                                  on the voxel grid, so the same two products run as zero-padded 3-D FFT convolutions with the
                                  stationary tables (circulant embedding, fp64, exact up to rounding).                        */
} gb_hyper;

/* Builds the device-resident problem: computes both sensitivity matrices on the GPU
 * (Inversion.cubing, inversion.py:216-230, without materialising the zero-padded Asens3).
 * Lean mode: when the two fp64 matrices would take more than a fifth of the device memory (96x96x48 and up; GEOBO_B200_LEAN_A=1/0
 * forces it) they are NOT kept: only their int8 digit blocks are, and every fp64 pass over them regenerates them chunk by chunk.
 * Such a problem only runs the int8 paths (gb_hyper.slices = 4, 5, 6); slices = 0 is refused with GB_ERR_UNSUPPORTED. */
int gb_problem_create(gb_ctx* ctx, const gb_problem_desc* desc, gb_problem** out);
int gb_problem_destroy(gb_problem* p);
/* Fs3 = normalised [grav, mag, drill] data vector, length M = 2*nsens + ndrill (inversion.py:221). */
int gb_problem_set_data(gb_problem* p, const double* fs3);
/* Inversion.predict3 (inversion.py:77-122) without the 3N x 3N covariance:
 * mu and var = diag(cov) are returned for this rank's voxel-column shard as [3][col_end - col_begin]
 * (property-major; the whole 3N vector on one GPU); logl and info are identical on every rank.
 * mu / var / logl may be NULL (results stay on the device).  Returns info (>0) if AkA is not PD -- also when a covariance value is
 * not finite (e.g. matern32 with equal length scales), which makes the reference's Cholesky raise (inversion.py:98-104).
 * With ndrill == 0 the third property block of mu / var is NOT computed and returned as NaN: the reference multiplies it by the std
 * of an empty array (inversion.py:213-214), so its drill cubes are NaN as well. */
int gb_predict(gb_problem* p, const gb_hyper* h, int flags, double* mu, double* var, double* logl, int* info);
/* Inversion.calc_logl (inversion.py:125-152): returns -logl without the N log(2 pi) term; +inf if not PD. */
int gb_neg_logl(gb_problem* p, const gb_hyper* h, double* neg_logl, int* info);
/* Asens_grav / Asens_mag as dense host arrays (nsens x N) -- for callers that read Inversion.Asens3. */
int gb_problem_get_sens(gb_problem* p, int kind, double* out);
/* simcube.create_synsurvey (simcube.py:147-150): out[nsens] = A_kind . x[N] */
int gb_forward(gb_problem* p, int kind, const double* x, double* out);

/* ------------------------------------------------------------------ geobo/run_geobo.py (acquisition)
 * Exhaustive evaluation of the Bayesian-optimisation utility the reference feeds to scipy.optimize.shgo.
 * gb_acquisition_vertical: futility_vertical (run_geobo.py:175-200) for every column (a, b) of the (n0, n1, n2) cubes:
 *   out[a*n1 + b] = sum_z rec + kappa * sqrt(sum_z var) - beta * sum_z costs  for 0 < a < n0-1, 0 < b < n1-1, -inf elsewhere
 *   (the positive utility; the reference returns its negative).  costs may be NULL (zeros).
 * gb_acquisition_drill: futility_drill (run_geobo.py:203-235) for n candidates params[i] = (x0, y0, azimuth, dip):
 *   ray of `length` from (x0, y0, zmax), int(2*length/min(voxsize)) samples, voxel index (int(x/vx), int(y/vy), int(-z/vz))
 *   on axes (0, 1, 2), negative indices wrap as in NumPy, out[i] = 0 when the ray leaves the cube (the reference's except). */
int gb_acquisition_vertical(gb_ctx* ctx, const double* rec, const double* var, const double* costs, const int64_t shape[3],
                            double kappa, double beta, double* out);
int gb_acquisition_drill(gb_ctx* ctx, const double* rec, const double* var, const double* costs, const int64_t shape[3],
                         const double voxsize[3], double zmax, double length, const double* params, int64_t n, double kappa,
                         double beta, double* out);
/* ------------------------------------------------------------------ geobo/run_geobo.py (drill data)
 * align_drill (run_geobo.py:132-159; utils.align_drill2, utils.py:55-83): drill-core samples -> voxel cube.
 *   voxelpos [3][n_vox]: voxel-centre coordinates as Inversion.create_cubegeometry returns them (rows x, y, z); the output
 *   is flat in the same voxel order, so the caller reshapes it like the arrays xxx / yyy / zzz it would have passed;
 *   coord [ns][3], data [ns]: sample positions (local coordinates) and values; voxsize = (xvoxsize, yvoxsize, zvoxsize).
 *   out[v] = mean of the non-NaN samples with  centre - voxsize <= coordinate < centre + voxsize  on all three axes (a
 *   window of two voxel sizes, as coded), 0 where there is no such sample or the mean is not finite.  ns may be 0. */
int gb_align_drill(gb_ctx* ctx, const double* voxelpos, int64_t n_vox, const double* coord, const double* data, int64_t ns,
                   const double voxsize[3], double* out);
/* Dense posterior covariance block (small cubes only; inversion.py:117): out (3N x 3N). */
int gb_posterior_cov(gb_problem* p, const gb_hyper* h, double* out);
int gb_get_timings(gb_problem* p, double* ms, int n);
/* bytes of device memory held by the problem */
uint64_t gb_problem_device_bytes(gb_problem* p);

#ifdef __cplusplus
}
#endif
#endif /* GEOBO_B200_H */
