// Shared pieces of the int8 digit-slice (Ozaki-type) tensor-core kernels: tile configuration per slice count,
// balanced-digit extraction and scaling exponents.  See ozaki.cu (fused covariance generation + projection) and
// ozaki_gemm.cu (both operands from pre-tiled digit blocks in memory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ozaki {

constexpr int TABLE_PAD = 64;
__host__ __device__ constexpr long plane_stride(long ext) { return (ext + TABLE_PAD + 15) & ~15L; }   // keeps every digit plane 16-byte aligned

// Tile shapes per slice count S.  Every kernel keeps S int32 accumulators (one per significance level) of NT TMEM columns.
//   * projection (ozaki.cu, A operand generated into tensor memory): NTP sensor rows per tile; the A digits need
//     NABUF x S x 8 TMEM columns (double buffered; deeper buffers with narrower tiles measured slower),
//     S * NTP + NABUF * 8 * S <= 512.  NTP is also the row-tile size of the sensitivities' digit blocks in memory.
//   * GEMM with both operands in shared memory (ozaki_gemm.cu): NT output columns, S * NT <= 512, STAGES-deep ring.
template <int S> struct Cfg;
template <> struct Cfg<4> { static constexpr int NT = 128, NTP = 112, STAGES = 6; };
template <> struct Cfg<5> { static constexpr int NT = 96, NTP = 80, STAGES = 6; };
template <> struct Cfg<6> { static constexpr int NT = 80, NTP = 64, STAGES = 5; };
constexpr int NABUF = 2;

// ------------------------------------------------------------------------------------------------ digit extraction
// t in [-1/2, 1/2]: BALANCED digits (every d_q in [-128, 127], stored as two's-complement bytes) of t rounded to
// nearest at the last digit:  t ~ sum_q d_q 2^-(7 + 8 q).  All digits signed => one instruction descriptor for every
// digit pair, several B digits can share one wide-N MMA, and the worst-case accumulator growth is 4x smaller.
template <int S>
__device__ __forceinline__ void digits(double t, uint8_t (&d)[S]) {
    t += ldexp(1.0, -(7 + 8 * (S - 1) + 1));
    int v[S];
    double x = t * 128.0;
    double f = floor(x);
    v[0] = (int)f;                       // in [-64, 64]
    double r = x - f;
#pragma unroll
    for (int q = 1; q < S; ++q) {
        x = r * 256.0;
        f = floor(x);
        v[q] = (int)f;                   // in [0, 255]
        r = x - f;
    }
#pragma unroll
    for (int q = S - 1; q >= 1; --q)
        if (v[q] >= 128) { v[q] -= 256; v[q - 1] += 1; }
#pragma unroll
    for (int q = 0; q < S; ++q) d[q] = (uint8_t)(int8_t)v[q];
}

// exponent e with |x| <= 2^(e-1)  (so t = x / 2^e lies in [-1/2, 1/2])
__device__ __forceinline__ int scale_exp(double amax) {
    if (!(amax > 0.0) || !isfinite(amax)) return 0;
    int e;
    frexp(amax, &e);      // amax = m 2^e, m in [0.5, 1)
    return e + 1;
}

}  // namespace ozaki
