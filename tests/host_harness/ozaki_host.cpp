// Host build of the digit extraction of csrc/ozaki.cuh (TEST HARNESS, compiled by tests/test_digit_slices.py with g++ and the
// CUDA include directory: outside nvcc the CUDA headers define __device__ / __forceinline__ away, so ozaki::digits<S> and
// ozaki::scale_exp compile unchanged).
#include <math.h>

#include "../../geobo_b200/csrc/ozaki.cuh"

template <int S>
static void run(const double* t, long n, uint8_t* out) {
    for (long i = 0; i < n; ++i) {
        uint8_t d[S];
        ozaki::digits<S>(t[i], d);
        for (int q = 0; q < S; ++q) out[i * S + q] = d[q];
    }
}

extern "C" int host_digits(int S, const double* t, long n, uint8_t* out) {
    if (S == 4) run<4>(t, n, out);
    else if (S == 5) run<5>(t, n, out);
    else if (S == 6) run<6>(t, n, out);
    else return -1;
    return 0;
}

extern "C" int host_scale_exp(double amax) { return ozaki::scale_exp(amax); }
