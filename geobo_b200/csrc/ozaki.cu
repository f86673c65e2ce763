// Error-free (Ozaki-type) int8 slice products on the tcgen05 tensor cores for the fused covariance
// assembly + projection  Pt = A . K  (the 95 % kernel of the path).
//
// tcgen05 has no fp64 kind and the path is too ill-conditioned for tf32/bf16, so the
// fp64 operands are split into S fixed-point digits of 7 / 8 bits:
//     x = 2^e * sum_q d_q 2^-(7+8q),   balanced digits d_q in [-128, 127] (all signed),
// with one exponent per sensor row of A and one per covariance table.  Digit products are accumulated EXACTLY
// in int32 by `tcgen05.mma.kind::i8` (SASS UTCIMMA), one TMEM accumulator per significance level
// l = qa + qb <= S-1 (S(S+1)/2 digit products per K = 32 step, issued as wide-N instructions that multiply one A digit
// with up to 256 / NT consecutive B digits at once), flushed every CHUNK contraction indices (overflow bound)
// into the fp64 result with an exact int64 recombination.  The only error is the truncation of the operands
// to 7 + 8(S-1) bits (S = 5: 2^-39 relative to the row / table scale) and of levels >= S.
//
// Tile = 128 voxel columns (M side) x NT sensor rows (N side); CTA = 18 warps, persistent (one CTA per SM):
//   * warps 6-17, covariance-digit producers: the M-side operand is GENERATED and written straight into TENSOR MEMORY
//     (tcgen05.st), from where the MMA reads its A operand -- no shared-memory traffic for it.  For 16 consecutive output
//     voxels i and 16 consecutive contraction voxels j (both in one z-column, z fastest) the 16 x 16 digit block is
//     Toeplitz in the stationary-covariance byte table (31 table bytes): a half-warp loads the aligned words once (one
//     32-bit load per lane), every lane picks its unaligned 16-byte window with five shuffles + funnel shifts.  A warp
//     owns one TMEM lane quarter and every third K step; lattice ids are prefetched two own steps ahead, table words one.
//   * warp 5, one lane: the N-side operand (sensitivity digits, stored in global memory PRE-TILED in the UMMA canonical
//     K-major no-swizzle layout [row tile of NT][k step][digit][NT x 32 B]) is one contiguous block per K step, fetched
//     with a single bulk async copy (cp.async.bulk -> SASS UBLKCP) into a deep shared-memory ring;
//   * warp 4, MMA issue: the whole warp runs the uniform loop with running ring state, one elected lane issues; one A
//     digit (TMEM) times up to 256 / NT consecutive B digit planes per instruction; ONE full / done mbarrier pair and
//     ONE tcgen05.commit per K step release the TMEM A buffer and the shared-memory stage together;
//   * warps 0-3, epilogue: TMEM -> exact int64 recombination -> fp64 read-modify-write of the Pt tile (32 lanes write 32
//     consecutive doubles of one Pt row).
// TMEM: S x NT accumulator columns + 2 x S x 8 columns of double-buffered A digits (480 of 512 at S = 5, NT = 80).
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"
#include "ozaki.cuh"

using namespace umma;

namespace ozaki {

__global__ void row_absmax_kernel(const double* __restrict__ A, long rows, long cols, long ld, int* __restrict__ exps) {
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    double m = 0.0;
    for (long c = lane; c < cols; c += 32) m = fmax(m, fabs(A[row * ld + c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) exps[row] = scale_exp(m);
}

// running row maxima over column chunks of an operand that is never resident as a whole
__global__ void row_absmax_accum_kernel(const double* __restrict__ A, long rows, long cols, long ld, double* __restrict__ amax) {
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    double m = 0.0;
    for (long c = lane; c < cols; c += 32) m = fmax(m, fabs(A[row * ld + c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) amax[row] = fmax(amax[row], m);
}

__global__ void exps_from_absmax_kernel(const double* __restrict__ amax, long rows, int* __restrict__ exps) {
    const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row < rows) exps[row] = scale_exp(amax[row]);
}

// A (rows x ld fp64, cols valid) -> pre-tiled digit blocks [row tile of TR rows][k step][digit][TR x 32 B canonical layout];
// one thread per (row, 16 consecutive contraction indices): S 16-byte stores.  Rows >= rows and columns >= cols are zero.
// TR = 128 for an M-side (A) operand, TR = NT for an N-side (B) operand.
template <int S>
__global__ void slice_rows_tiled_kernel(const double* __restrict__ A, long rows, long cols, long ld, const int* __restrict__ exps,
                                        uint8_t* __restrict__ out, long ksteps, int TR, long ks0, long ks_count) {
    // ksteps: k steps of the whole layout; this launch fills steps [ks0, ks0 + ks_count) from the columns of A (a column chunk)
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;    // 16-column group of the chunk
    const long row = blockIdx.y;
    if (g >= 2 * ks_count) return;
    uint32_t pk[S][4];
#pragma unroll
    for (int q = 0; q < S; ++q) pk[q][0] = pk[q][1] = pk[q][2] = pk[q][3] = 0u;
    if (row < rows) {
        const int e = exps[row];
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const long col = g * 16 + b;
            if (col < cols) {
                uint8_t d[S];
                digits<S>(ldexp(A[row * ld + col], -e), d);
#pragma unroll
                for (int q = 0; q < S; ++q) pk[q][b >> 2] |= (uint32_t)d[q] << (8 * (b & 3));
            }
        }
    }
    const long tile = row / TR, ks = ks0 + (g >> 1);
    const long blk = (long)TR * 32;
    const uint32_t off = core_offset((uint32_t)(row % TR), (uint32_t)(g & 1));
#pragma unroll
    for (int q = 0; q < S; ++q)
        *reinterpret_cast<uint4*>(out + ((tile * ksteps + ks) * S + q) * blk + off) = make_uint4(pk[q][0], pk[q][1], pk[q][2], pk[q][3]);
}

__global__ void table_absmax_kernel(const double* __restrict__ tab, long ext, int* __restrict__ exps) {
    __shared__ double red[256];
    const double* t = tab + (long)blockIdx.x * ext;
    double m = 0.0;
    for (long e = threadIdx.x; e < ext; e += blockDim.x) m = fmax(m, fabs(t[e]));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) exps[blockIdx.x] = scale_exp(red[0]);
}

// table (ntab x ext fp64) -> byte planes [tab][q][ext + TABLE_PAD]
template <int S>
__global__ void slice_table_kernel(const double* __restrict__ tab, long ext, const int* __restrict__ exps, uint8_t* __restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tb = blockIdx.y;
    if (e >= plane_stride(ext)) return;
    uint8_t d[S];
    if (e < ext) {
        digits<S>(ldexp(tab[(long)tb * ext + e], -exps[tb]), d);
    } else {
#pragma unroll
        for (int q = 0; q < S; ++q) d[q] = 0;
    }
#pragma unroll
    for (int q = 0; q < S; ++q) out[((long)tb * S + q) * plane_stride(ext) + e] = d[q];
}

// ------------------------------------------------------------------------------------------------ main kernel
struct Params {
    const uint8_t* a8[2];     // pre-tiled digit blocks of A_grav / A_magn, N-side layout: [row tile of NT][k step][digit][NT*32]
    const int* a_exp[2];      // [Ns]
    const uint8_t* t8;        // [9][S][plane_stride(ext)] digit planes of the covariance tables (only c = 0, 1 are used)
    const int* t_exp;         // [9]
    const int* L;             // [kp] extended-lattice ids
    double* Pt;               // [Mp][ldp]
    long ext, C0, kp, ldp, ncp;
    int Ns, ncol, c0, chunk;  // chunk: contraction indices per accumulator flush (multiple of 32)
    int nr;                   // property blocks per data block (3, or 2 without drill data): task t = (c = t / nr, r = t % nr)
    // zero-digit culling: cull[2 * task + 0 / 1] = largest |dy| / |dx| lattice offset at which ANY digit plane of the task's
    // covariance table is non-zero (device memory, written by table_extent_kernel); null = visit every K step
    const int* cull;
    int xN, yN, zN;
    // tile-round pacing: the CTAs take tiles round-robin, and the (up to) gridDim.x tiles of one round share ONE sensitivity-digit
    // K strip (same sensor-row tile), which therefore only has to come from HBM once -- as long as the CTAs stay within an
    // L2-full of each other.  Left alone they drift apart (measured at 64x64x32: 3.98 TB read from HBM for a 5.4 GB operand), so
    // the copy lanes count arrivals in this device counter and start a round's loads together.  null = no pacing.
    unsigned int* sync_ctr;
    int sync_slack;                  // rounds a CTA may run ahead of the slowest one (0: all start a round together)
    // streamed contraction: this launch covers only the contraction voxels of the voxel rows [cy0, cy1) (cy1 <= cy0: all rows);
    // a8 then holds the digit blocks of just those columns, its first K step being step ks_base of the cube, and the result
    // is added to Pt (accumulate != 0) instead of replacing it
    int ks_base, cy0, cy1, accumulate;
    // tile order: with culling the tiles of one round have different numbers of K steps (cube edges), and paced CTAs would wait
    // for the longest one.  perm[tq * n_itile + pos] = voxel-column tile handled at position pos of task tq, sorted by decreasing
    // step count, so that the (up to) gridDim.x tiles of a round cost the same.  null = natural order.
    const int* perm;
    unsigned long long* steps_ctr;   // += K steps visited (one add per tile), for the roofline accounting; or null
    int n_stile, n_itile;     // sensor-row tiles of NT, voxel-column tiles of 128
};

// K steps a tile has to visit.  Beyond the offsets (ey, ex) every digit of the task's covariance table is zero (a kernel with
// compact support, or one that decays below the last digit: exp(-d^2 / 2 gamma^2) < 2^-39 beyond 7.3 gamma), so a K step whose
// 32 contraction voxels all lie further than that from all 128 output voxels of the tile multiplies by zero digits only: skipping
// it leaves every integer accumulator -- and the result -- bitwise unchanged.  The contraction index is y-major, so the kept steps
// are `nrows` voxel rows jy >= jya with the same window of `w` steps at offset `koff` inside each row (rowsteps = xN zN / 32 per
// row):  ks(t) = (jya + t / w) * rowsteps + koff + t % w,  t = 0 .. nt - 1.  Without culling: one "row" of all ksteps steps.
struct KRange {
    int jya, nrows, koff, w, rowsteps;
    uint32_t nt;
    __device__ __forceinline__ int ks(uint32_t t) const { const uint32_t jj = t / (uint32_t)w; return (jya + (int)jj) * rowsteps + koff + (int)(t - jj * (uint32_t)w); }
};

__device__ __forceinline__ KRange tile_krange(const Params& P, int task, int itile, int ksteps) {
    KRange r;
    const int XZ = P.xN * P.zN;
    if (P.cull == nullptr || (XZ & 31) != 0) {
        r.jya = 0; r.nrows = 1; r.koff = P.ks_base; r.w = ksteps; r.rowsteps = 0; r.nt = (uint32_t)ksteps;
        return r;
    }
    const int ey = P.cull[2 * task], ex = P.cull[2 * task + 1];
    const long g0 = (long)P.c0 + (long)itile * 128;
    const long g1 = min(g0 + 127, (long)P.c0 + P.ncol - 1);
    const int iya = (int)(g0 / XZ), iyb = (int)(g1 / XZ);
    int ixa = 0, ixb = P.xN - 1;
    if (iya == iyb) { ixa = (int)(g0 % XZ) / P.zN; ixb = (int)(g1 % XZ) / P.zN; }
    int jya = max(0, iya - ey), jyb = min(P.yN - 1, iyb + ey);
    if (P.cy1 > P.cy0) { jya = max(jya, P.cy0); jyb = min(jyb, P.cy1 - 1); }       // streamed contraction: rows of this launch only
    const int jxa = max(0, ixa - ex), jxb = min(P.xN - 1, ixb + ex);
    if (jyb < jya) { r.jya = 0; r.nrows = 0; r.koff = 0; r.w = 1; r.rowsteps = 0; r.nt = 0; return r; }
    r.rowsteps = XZ / 32;
    r.jya = jya;
    r.nrows = jyb - jya + 1;
    r.koff = (jxa * P.zN) / 32;
    r.w = ((jxb + 1) * P.zN + 31) / 32 - r.koff;
    r.nt = (uint32_t)(r.nrows * r.w);
    return r;
}

// extents of the non-zero digits of the 9 covariance tables: out[2 tb + 0 / 1] = max |dy| / |dx| with a non-zero digit in any plane
template <int S>
__global__ void table_extent_kernel(const uint8_t* __restrict__ t8, long ext, int xN, int yN, int zN, int* __restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tb = blockIdx.y;
    if (e >= ext) return;
    const long plane = plane_stride(ext);
    bool nz = false;
#pragma unroll
    for (int q = 0; q < S; ++q) nz |= t8[((long)tb * S + q) * plane + e] != 0;
    if (!nz) return;
    const int EX = 2 * xN - 1, EZ = 2 * zN - 1;
    const long t = e / EZ;
    const int dx = (int)(t % EX) - (xN - 1), dy = (int)(t / EX) - (yN - 1);
    atomicMax(out + 2 * tb, abs(dy));
    atomicMax(out + 2 * tb + 1, abs(dx));
}

// voxel-column tile at position `pos` of task tq
// (rem = tile index inside the task = stile * n_itile + pos).  The sorted order runs from the most to the fewest K steps for even
// sensor-row tiles and back for odd ones: a round of gridDim.x consecutive tiles that straddles two sensor-row tiles then holds
// tiles of similar cost from both (n_itile is not a multiple of the grid size), instead of the cheapest of one and the dearest of
// the next, which made every CTA of that round wait for the dear ones.
__device__ __forceinline__ int tile_at(const Params& P, int tq, long rem) {
    const int stile = (int)(rem / P.n_itile);
    int pos = (int)(rem - (long)stile * P.n_itile);
    if (!P.perm) return pos;
    if (stile & 1) pos = P.n_itile - 1 - pos;
    return P.perm[(long)tq * P.n_itile + pos];
}

// keys[tq][i] = K steps of voxel-column tile i in task tq
__global__ void tile_steps_kernel(const Params P, int ksteps, int* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, tq = blockIdx.y;
    if (i < P.n_itile) keys[(long)tq * P.n_itile + i] = (int)tile_krange(P, (tq / P.nr) * 3 + tq % P.nr, i, ksteps).nt;
}

// perm[tq][rank of i] = i, rank by decreasing key (ties: increasing i) -- n_itile is a few thousand: a rank by counting is enough
__global__ void tile_rank_kernel(const int* __restrict__ keys, int n, int* __restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, tq = blockIdx.y;
    if (i >= n) return;
    const int* k = keys + (long)tq * n;
    const int ki = k[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
        const int kj = k[j];
        rank += (kj > ki) || (kj == ki && j < i);
    }
    perm[(long)tq * n + rank] = i;
}

constexpr int PROD_SLOTS = 3;                          // producer warps per TMEM lane quarter
constexpr int TS_THREADS = (6 + 4 * PROD_SLOTS) * 32;  // 4 epilogue + MMA + copy + 12 producer warps
// shared-memory ring of the sensitivity digits: the stages are small (S * NT * 32 B) and a bulk copy takes a few thousand
// clocks end to end, so the ring is as deep as shared memory allows (<= 32 stages, <= 200 KB)
template <int S> __host__ __device__ constexpr int ring_stages() { return (200 * 1024) / (S * Cfg<S>::NTP * 32) < 32 ? (200 * 1024) / (S * Cfg<S>::NTP * 32) : 32; }

// Tile = 128 voxel columns (M side) x NT sensor rows (N side); D[m][n] = sum_j K[(c,j),(r,i0+m)] * A_c[s0+n][j].
//   * M-side operand = covariance digits, GENERATED by the producer warps straight into TENSOR MEMORY (tcgen05.st; the MMA
//     reads its A operand from TMEM, so it costs no shared-memory bandwidth and no small-N penalty);
//   * N-side operand = sensitivity digits, one bulk async copy per K step into an 8-deep shared-memory ring;
//   * accumulators: S levels x NT columns of TMEM; 2 x S x 8 columns hold the double-buffered A digits.
template <int S>
__global__ void __launch_bounds__(TS_THREADS, 1) ozaki_project_kernel(const __grid_constant__ Params P) {
    constexpr int NT = Cfg<S>::NTP, SB = ring_stages<S>();
    constexpr int B_SLICE = NT * 32, B_BYTES = S * B_SLICE;
    constexpr int TMEM_COLS = 512, ACC = S * NT, ABUF = 8 * S;
    constexpr int G = (256 / NT) < S ? (256 / NT) : S;          // B digit planes per MMA instruction (N <= 256)
    static_assert(ACC + NABUF * ABUF <= TMEM_COLS, "accumulators + A digit buffers do not fit in TMEM");
    static_assert(SB >= 2 * NABUF, "ring must be deeper than the A buffers (barrier parities stay unambiguous)");
    extern __shared__ __align__(1024) uint8_t smem[];
    // one full / done barrier pair per ring slot (slot = it % SB): `full` collects the four producer warps of the step and the
    // bytes of the sensitivity-digit bulk copy, `done` is signalled by ONE tcgen05.commit per step and releases both the
    // TMEM A buffer (to the producers of step it + NABUF) and the shared-memory stage (to the copy of step it + SB)
    __shared__ uint64_t full_bar[SB], done_bar[SB], tfull_bar, tempty_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 4) {
        tmem_alloc(&tmem_base_s, TMEM_COLS);
        if (lane == 0) {
            for (int s = 0; s < SB; ++s) { mbar_init(&full_bar[s], 5); mbar_init(&done_bar[s], 1); }
            mbar_init(&tfull_bar, 1);
            mbar_init(&tempty_bar, 128);
            fence_barrier_init();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const long tiles_per_task = (long)P.n_stile * P.n_itile;
    const long ntiles = 2 * P.nr * tiles_per_task;
    const int ksteps = (int)(P.kp / 32);
    const int chunk_steps = P.chunk / 32;

    if (warp >= 6) {
        // =============================================================== covariance-digit producers -> TMEM
        // warp <-> (TMEM lane quarter q = warp % 4, slot sl): handles the K steps with it % PROD_SLOTS == sl for rows 32 q ..
        const int q4 = warp & 3, sl = (warp - 6) >> 2;
        const int n = lane & 15;                       // column inside the 16-column segment of this half-warp
        const long plane = plane_stride(P.ext);
        const int C0 = (int)P.C0;
        uint32_t it_tile0 = 0;
        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int tq = (int)(tile / tiles_per_task);
            const int task = (tq / P.nr) * 3 + tq % P.nr;             // c * 3 + r
            const int itile = tile_at(P, tq, tile % tiles_per_task);
            const KRange kr = tile_krange(P, task, itile, ksteps);
            const int nt = (int)kr.nt;
            const int i0 = itile * 128 + 32 * q4 + (lane & 16);       // first voxel column of this half-warp's segment
            const uint8_t* t8 = P.t8 + (size_t)task * S * plane;
            const int lseg = P.L[P.c0 + min(i0, P.ncol - 16)];
            int ks = (int)((uint32_t)(sl + PROD_SLOTS - (int)(it_tile0 % PROD_SLOTS)) % PROD_SLOTS);
            // software pipeline over this warp's K steps: the lattice ids are loaded two own steps ahead, the table words one
            // own step ahead, the digit windows are assembled in registers BEFORE waiting for the TMEM buffer, so that only the
            // five tcgen05.st sit between "MMAs of step it-2 done" and "A digits of step it ready"
            int lj_nn[2];                                  // lattice ids of step ks + PROD_SLOTS
            uint32_t word[2][S];                           // table words of step ks
            int wi[2];
            uint32_t sh[2];
            // K step of tile step t, for the monotone stream t = ks, ks + PROD_SLOTS, ks + 2 PROD_SLOTS, ... of this warp: running
            // (row, step inside the row window) instead of a division per step (this warp sits on the done -> st -> full chain)
            int lj_row = 0, lj_q = 0, lj_t = 0;
            const int lj_last = nt > 0 ? kr.ks((uint32_t)(nt - 1)) : 0;
            auto load_lj = [&](int (&lj)[2]) {
                const int kk = lj_t < nt ? (kr.jya + lj_row) * kr.rowsteps + kr.koff + lj_q : lj_last;
                lj[0] = P.L[kk * 32];
                lj[1] = P.L[kk * 32 + 16];
                lj_t += PROD_SLOTS;
                lj_q += PROD_SLOTS;
                while (lj_q >= kr.w) { lj_q -= kr.w; ++lj_row; }
            };
            auto load_words = [&](const int (&lj)[2]) {
#pragma unroll
                for (int kh = 0; kh < 2; ++kh) {
                    const int offmin = C0 + lj[kh] - lseg - 15;       // table offset of (column n = 15, first j of the half); symmetric table
                    const int d = (offmin & 3) + 15 - n;              // this lane's window starts d bytes after the aligned base
                    wi[kh] = d >> 2;
                    sh[kh] = (uint32_t)(d & 3) * 8;
                    const uint8_t* src = t8 + (offmin & ~3) + 4 * n;
#pragma unroll
                    for (int q = 0; q < S; ++q) word[kh][q] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)q * plane));
                }
            };
            if (nt > 0) {
                lj_t = ks;
                lj_row = ks / kr.w;                 // one division per tile
                lj_q = ks - lj_row * kr.w;
                int lj0[2];
                load_lj(lj0);                       // step ks
                load_lj(lj_nn);                     // step ks + PROD_SLOTS
                load_words(lj0);
            }
            for (; ks < nt; ks += PROD_SLOTS) {                // ks: step index t of this tile (kr.ks(t) = K step)
                const uint32_t it = it_tile0 + (uint32_t)ks;
                // ---- assemble the unaligned 16-byte windows of this step: v[q][0..3] = K half 0, v[q][4..7] = K half 1
                uint32_t v[S][8];
#pragma unroll
                for (int q = 0; q < S; ++q) {
#pragma unroll
                    for (int kh = 0; kh < 2; ++kh) {
                        const int src_lane = (lane & 16) + wi[kh];
                        const uint32_t w0 = __shfl_sync(0xffffffffu, word[kh][q], src_lane);
                        const uint32_t w1 = __shfl_sync(0xffffffffu, word[kh][q], src_lane + 1);
                        const uint32_t w2 = __shfl_sync(0xffffffffu, word[kh][q], src_lane + 2);
                        const uint32_t w3 = __shfl_sync(0xffffffffu, word[kh][q], src_lane + 3);
                        const uint32_t w4 = __shfl_sync(0xffffffffu, word[kh][q], src_lane + 4);
                        v[q][4 * kh + 0] = __funnelshift_r(w0, w1, sh[kh]);
                        v[q][4 * kh + 1] = __funnelshift_r(w1, w2, sh[kh]);
                        v[q][4 * kh + 2] = __funnelshift_r(w2, w3, sh[kh]);
                        v[q][4 * kh + 3] = __funnelshift_r(w3, w4, sh[kh]);
                    }
                }
                // ---- prefetch: table words of the next own step, lattice ids of the one after
                {
                    int ljn[2] = {lj_nn[0], lj_nn[1]};
                    load_lj(lj_nn);                 // step ks + 2 PROD_SLOTS
                    load_words(ljn);
                }
                // ---- the TMEM buffer (it % NABUF) is free once the MMAs of step it - NABUF have completed
                if (it >= NABUF) mbar_wait(&done_bar[(it - NABUF) % SB], ((it - NABUF) / SB) & 1);
                tc_fence_after();
                const uint32_t abuf = tmem_base + ((uint32_t)(32 * q4) << 16) + ACC + (it % NABUF) * ABUF;
#pragma unroll
                for (int q = 0; q < S; ++q) tmem_st8(abuf + 8 * q, v[q]);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[it % SB]);
            }
            it_tile0 += kr.nt;
        }
    } else if (warp == 5) {
        // =============================================================== sensitivity-digit copies (one elected lane)
        if (lane == 0) {
            uint32_t it = 0;
            for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                if (P.sync_ctr) {
                    // all CTAs are co-resident (grid <= SM count, one CTA per SM), so waiting on the others cannot deadlock.
                    // Arrivals up to and including round k: min(ntiles, (k + 1) gridDim.x).  A CTA entering round k waits until
                    // every CTA has entered round k - slack.
                    const long k = (tile - blockIdx.x) / gridDim.x - P.sync_slack;
                    atomicAdd(P.sync_ctr, 1u);
                    if (k >= 0) {
                        const unsigned int target = (unsigned int)min(ntiles, (k + 1) * (long)gridDim.x);
                        while (*(volatile unsigned int*)P.sync_ctr < target) __nanosleep(256);
                    }
                }
                const int tq = (int)(tile / tiles_per_task);
                const int stile = (int)((tile % tiles_per_task) / P.n_itile);
                const uint8_t* src = P.a8[tq / P.nr] + (size_t)stile * ksteps * B_BYTES;
                const KRange kr = tile_krange(P, (tq / P.nr) * 3 + tq % P.nr, tile_at(P, tq, tile % tiles_per_task), ksteps);
                if (P.steps_ctr) atomicAdd(P.steps_ctr, (unsigned long long)kr.nt);
                for (int jj = 0; jj < kr.nrows; ++jj) {
                    const uint8_t* rowsrc = src + (size_t)((kr.jya + jj) * kr.rowsteps + kr.koff - P.ks_base) * B_BYTES;
                    for (int q = 0; q < kr.w; ++q, ++it) {
                        const int st = (int)(it % SB);
                        mbar_wait(&done_bar[st], ((it / SB) & 1) ^ 1);
                        mbar_arrive_expect_tx(&full_bar[st], B_BYTES);
                        bulk_g2s(smem + st * B_BYTES, rowsrc + (size_t)q * B_BYTES, B_BYTES, &full_bar[st]);
                    }
                }
            }
        }
    } else if (warp == 4) {
        // =============================================================== MMA issuer: the whole warp runs the (uniform) loop,
        // one elected lane issues the tcgen05 instructions
        {
            uint32_t chunk_id = 0;
            // running ring state (no div / mod / multiplies in the issue loop): ring slot, its barrier addresses and phase,
            // the low word of the B descriptor (address >> 4) and the TMEM address of the A buffer
            int rp = 0, ab = 0;
            uint32_t ph = 0;
            const uint64_t bdesc0 = smem_desc(smem_u32(smem), kLBO, kSBO);
            uint64_t bdesc = bdesc0;
            uint32_t at = tmem_base + ACC;
            uint64_t* fb = full_bar;
            uint64_t* db = done_bar;
            for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int tq = (int)(tile / tiles_per_task);
                const int nt = (int)tile_krange(P, (tq / P.nr) * 3 + tq % P.nr, tile_at(P, tq, tile % tiles_per_task), ksteps).nt;
                for (int k0 = 0; k0 < nt; k0 += chunk_steps, ++chunk_id) {
                    const int k1 = min(nt, k0 + chunk_steps);
                    mbar_wait(&tempty_bar, (chunk_id & 1) ^ 1);     // epilogue has drained the accumulators
                    tc_fence_after();
                    for (int ks = k0; ks < k1; ++ks) {
                        mbar_wait(fb, ph);
                        tc_fence_after();
                        const uint32_t acc = ks == k0 ? 0u : 1u;
                        if (elect_one()) {
                        // covariance digit qa (TMEM) times sensitivity digits qb0 .. qb0+g-1 (consecutive planes of the canonical
                        // layout) in ONE instruction of N = g * NT columns; product (qa, qb) lands in level qa + qb
#pragma unroll
                        for (int qa = 0; qa < S; ++qa) {
#pragma unroll
                            for (int qb0 = 0; qb0 < S - qa; qb0 += G) {
                                const int g = (S - qa - qb0) < G ? (S - qa - qb0) : G;
                                mma_i8_ts(tmem_base + (qa + qb0) * NT, at + 8 * qa, bdesc + (uint64_t)((qb0 * B_SLICE) >> 4), idesc_i8(1, 1, g * NT),
                                          qa == 0 ? acc : 1u);
                            }
                        }
                        mma_commit(db);                   // TMEM A buffer and shared-memory stage reusable once these MMAs have read them
                        }
                        ++fb; ++db; bdesc += (uint64_t)(B_BYTES >> 4); at += ABUF;
                        if (++rp == SB) { rp = 0; ph ^= 1; fb = full_bar; db = done_bar; bdesc = bdesc0; }
                        if (++ab == NABUF) { ab = 0; at = tmem_base + ACC; }
                    }
                    if (elect_one()) mma_commit(&tfull_bar);               // accumulators of this chunk complete
                }
            }
        }
    } else {
        // =============================================================== epilogue (warps 0-3 <-> TMEM lanes 32 w .. 32 w + 31)
        uint32_t chunk_id = 0;
        const int m = warp * 32 + lane;                   // row of D = voxel column inside the tile
        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int tq = (int)(tile / tiles_per_task);
            const long rem = tile % tiles_per_task;
            const int stile = (int)(rem / P.n_itile), itile = tile_at(P, tq, rem);
            const int c = tq / P.nr, r = tq % P.nr, task = c * 3 + r;
            const int s0 = stile * NT, i = itile * 128 + m;
            const bool col_ok = i < P.ncol;
            // result = 2^(eA + eK - 14 - 8 (S-1)) * sum_l acc_l 2^(8 (S-1-l))
            const int ebase = P.t_exp[task] - 14 - 8 * (S - 1);
            const int* aexp = P.a_exp[c];
            double* pcol = P.Pt + ((long)c * P.Ns + s0) * P.ldp + (long)r * P.ncp + (col_ok ? i : 0);
            const int nt = (int)tile_krange(P, task, itile, ksteps).nt;
            for (int k0 = 0; k0 < nt; k0 += chunk_steps, ++chunk_id) {
                mbar_wait(&tfull_bar, chunk_id & 1);
                tc_fence_after();
                for (int n0 = 0; n0 < NT; n0 += 8) {
                    uint32_t v[S][8];
#pragma unroll
                    for (int lvl = 0; lvl < S; ++lvl) tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + lvl * NT + n0, v[lvl]);
                    tmem_ld_wait();
                    if (col_ok) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int s = s0 + n0 + k;
                            if (s < P.Ns) {
                                long long acc = (long long)(int)v[0][k];
#pragma unroll
                                for (int lvl = 1; lvl < S; ++lvl) acc = acc * 256 + (long long)(int)v[lvl][k];
                                const double add = ldexp((double)acc, aexp[s] + ebase);
                                double* dst = pcol + (long)(n0 + k) * P.ldp;          // 32 lanes -> 32 consecutive doubles of one Pt row
                                *dst = (k0 == 0 && !P.accumulate) ? add : *dst + add;
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int S>
static cudaError_t launch_one(const Params& P, int sm_count, cudaStream_t s, int* perm_scratch) {
    constexpr int NT = Cfg<S>::NTP;
    constexpr int smem = ring_stages<S>() * S * NT * 32;
    cudaError_t e = cudaFuncSetAttribute(ozaki_project_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    Params q = P;
    q.n_stile = (P.Ns + NT - 1) / NT;
    q.n_itile = (P.ncol + 127) / 128;
    const long ntiles = 2L * q.nr * q.n_stile * q.n_itile;
    const int grid = (int)(ntiles < (long)sm_count ? ntiles : (long)sm_count);
    if (q.cull && perm_scratch) {
        // sorted tile order (only matters when tiles differ, i.e. with culling): keys in the second half of the scratch
        int* keys = perm_scratch + (long)2 * q.nr * q.n_itile;
        dim3 g((unsigned)((q.n_itile + 127) / 128), (unsigned)(2 * q.nr));
        tile_steps_kernel<<<g, 128, 0, s>>>(q, (int)(q.kp / 32), keys);
        tile_rank_kernel<<<g, 128, 0, s>>>(keys, q.n_itile, perm_scratch);
        q.perm = perm_scratch;
    } else {
        q.perm = nullptr;
    }
    ozaki_project_kernel<S><<<grid, TS_THREADS, smem, s>>>(q);
    return cudaGetLastError();
}

}  // namespace ozaki

// ------------------------------------------------------------------------------------------------ host entry points
// Accumulator flush interval in contraction indices: level S-1 sums S products of balanced digits (|d| <= 128), so
// S * 2^14 * chunk < 2^31 holds for chunk = 16384 and S <= 7.  GEOBO_B200_CHUNK (multiple of 32, <= 16384) shortens it
// so that tests can exercise the multi-flush code paths on small cubes.
int ozaki_chunk() {
    static int chunk = 0;
    if (!chunk) {
        chunk = 16384;
        const char* e = getenv("GEOBO_B200_CHUNK");
        if (e) {
            const int v = atoi(e);
            if (v >= 32 && v <= 16384 && v % 32 == 0) chunk = v;
        }
    }
    return chunk;
}

int ozaki_tile_n(int slices) { return slices == 4 ? ozaki::Cfg<4>::NT : slices == 5 ? ozaki::Cfg<5>::NT : slices == 6 ? ozaki::Cfg<6>::NT : 0; }
// row-tile size of the sensitivities' digit blocks (N-side operand of the projection and of the AkA products)
int ozaki_tile_np(int slices) { return slices == 4 ? ozaki::Cfg<4>::NTP : slices == 5 ? ozaki::Cfg<5>::NTP : slices == 6 ? ozaki::Cfg<6>::NTP : 0; }

// bytes of the pre-tiled digit blocks of a rows x kp operand with `tr` rows per tile (128: M side, tile_n: N side)
long ozaki_rows_bytes(long rows, long kp, int slices, int tr) { return ((rows + tr - 1) / tr) * (kp / 32) * (long)slices * tr * 32; }

cudaError_t ozaki_slice_rows(const double* A, long rows, long cols, long ld, int slices, int* exps, uint8_t* out, long kp, int tr, cudaStream_t s) {
    ozaki::row_absmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(A, rows, cols, ld, exps);
    const long ksteps = kp / 32, rows_p = (rows + tr - 1) / tr * tr;
    dim3 grid((unsigned)((2 * ksteps + 127) / 128), (unsigned)rows_p);
    switch (slices) {
        case 4: ozaki::slice_rows_tiled_kernel<4><<<grid, 128, 0, s>>>(A, rows, cols, ld, exps, out, ksteps, tr, 0, ksteps); break;
        case 5: ozaki::slice_rows_tiled_kernel<5><<<grid, 128, 0, s>>>(A, rows, cols, ld, exps, out, ksteps, tr, 0, ksteps); break;
        case 6: ozaki::slice_rows_tiled_kernel<6><<<grid, 128, 0, s>>>(A, rows, cols, ld, exps, out, ksteps, tr, 0, ksteps); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t ozaki_row_absmax_accum(const double* A, long rows, long cols, long ld, double* amax, cudaStream_t s) {
    ozaki::row_absmax_accum_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(A, rows, cols, ld, amax);
    return cudaGetLastError();
}

cudaError_t ozaki_exps_from_absmax(const double* amax, long rows, int* exps, cudaStream_t s) {
    ozaki::exps_from_absmax_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(amax, rows, exps);
    return cudaGetLastError();
}

// digits of a column chunk (cols columns of A, ld) into k steps [ks0, ks0 + ceil(cols / 32)) of a layout with kp_total / 32 steps
cudaError_t ozaki_slice_rows_range(const double* A, long rows, long cols, long ld, int slices, const int* exps, uint8_t* out, long kp_total,
                                   int tr, long ks0, cudaStream_t s) {
    const long ksteps = kp_total / 32, ks_count = (cols + 31) / 32, rows_p = (rows + tr - 1) / tr * tr;
    if (ks0 < 0 || ks0 + ks_count > ksteps) return cudaErrorInvalidValue;
    dim3 grid((unsigned)((2 * ks_count + 127) / 128), (unsigned)rows_p);
    switch (slices) {
        case 4: ozaki::slice_rows_tiled_kernel<4><<<grid, 128, 0, s>>>(A, rows, cols, ld, exps, out, ksteps, tr, ks0, ks_count); break;
        case 5: ozaki::slice_rows_tiled_kernel<5><<<grid, 128, 0, s>>>(A, rows, cols, ld, exps, out, ksteps, tr, ks0, ks_count); break;
        case 6: ozaki::slice_rows_tiled_kernel<6><<<grid, 128, 0, s>>>(A, rows, cols, ld, exps, out, ksteps, tr, ks0, ks_count); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// sensitivities: N-side operand of the projection (tiles of ozaki_tile_n rows)
cudaError_t ozaki_slice_sens(const double* A, long rows, long cols, long ld, int slices, int* exps, uint8_t* out, long kp, cudaStream_t s) {
    return ozaki_slice_rows(A, rows, cols, ld, slices, exps, out, kp, ozaki_tile_np(slices), s);
}

cudaError_t ozaki_slice_tables(const double* tables, long ext, int slices, int* exps, uint8_t* out, cudaStream_t s) {
    ozaki::table_absmax_kernel<<<9, 256, 0, s>>>(tables, ext, exps);
    dim3 grid((unsigned)((ozaki::plane_stride(ext) + 255) / 256), 9);
    switch (slices) {
        case 4: ozaki::slice_table_kernel<4><<<grid, 256, 0, s>>>(tables, ext, exps, out); break;
        case 5: ozaki::slice_table_kernel<5><<<grid, 256, 0, s>>>(tables, ext, exps, out); break;
        case 6: ozaki::slice_table_kernel<6><<<grid, 256, 0, s>>>(tables, ext, exps, out); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t ozaki_table_extents(const uint8_t* t8, long ext, int slices, const int64_t n[3], int* cull, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(cull, 0, 18 * sizeof(int), s);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((ext + 255) / 256), 9);
    switch (slices) {
        case 4: ozaki::table_extent_kernel<4><<<grid, 256, 0, s>>>(t8, ext, (int)n[0], (int)n[1], (int)n[2], cull); break;
        case 5: ozaki::table_extent_kernel<5><<<grid, 256, 0, s>>>(t8, ext, (int)n[0], (int)n[1], (int)n[2], cull); break;
        case 6: ozaki::table_extent_kernel<6><<<grid, 256, 0, s>>>(t8, ext, (int)n[0], (int)n[1], (int)n[2], cull); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

long ozaki_table_bytes(long ext, int slices) { return 9L * slices * ozaki::plane_stride(ext); }

cudaError_t ozaki_project(const OzakiArgs& a, int slices, int sm_count, cudaStream_t s) {
    ozaki::Params P;
    P.a8[0] = a.a8[0]; P.a8[1] = a.a8[1]; P.a_exp[0] = a.a_exp[0]; P.a_exp[1] = a.a_exp[1];
    P.t8 = a.t8; P.t_exp = a.t_exp; P.L = a.L; P.Pt = a.Pt;
    P.ext = a.ext; P.C0 = a.C0; P.kp = a.kp; P.ldp = a.ldp; P.ncp = a.ncp;
    P.Ns = a.Ns; P.ncol = a.ncol; P.c0 = a.c0;
    P.nr = a.nr == 2 ? 2 : 3;
    P.cull = a.cull;
    P.sync_ctr = a.sync_ctr;
    P.sync_slack = a.sync_slack;
    P.perm = nullptr;
    P.ks_base = a.ks_base; P.cy0 = a.cy0; P.cy1 = a.cy1; P.accumulate = a.accumulate;
    P.steps_ctr = a.steps_ctr;
    P.xN = a.n[0]; P.yN = a.n[1]; P.zN = a.n[2];
    P.chunk = ozaki_chunk();
    P.n_stile = 0;
    P.n_itile = 0;
    switch (slices) {
        case 4: return ozaki::launch_one<4>(P, sm_count, s, a.perm_scratch);
        case 5: return ozaki::launch_one<5>(P, sm_count, s, a.perm_scratch);
        case 6: return ozaki::launch_one<6>(P, sm_count, s, a.perm_scratch);
    }
    return cudaErrorInvalidValue;
}
