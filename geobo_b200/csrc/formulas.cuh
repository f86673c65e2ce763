// Scalar fp64 formulas of the path, shared by the device kernels (cov.cu, sens.cu) and by the host harness the CPU tests
// compile (tests/host_harness/formulas_host.cpp): the covariance functions of geobo/kernels.py:81-195 with the block
// selection of create_cov, and the prism-corner potentials of geobo/sensormodel.py:96-133.  Under nvcc these are
// __device__ __forceinline__ exactly as before; under a host compiler the rounding intrinsics map to the plain IEEE
// operations (build the harness with -ffp-contract=off), so the only difference to the device is libm vs CUDA
// exp / log / atan / sin / cos in the last ulp.
#pragma once
#include "../../include/geobo_b200.h"

#if defined(__CUDACC__)
#define GB_DEV __device__ __forceinline__
#else
#include <math.h>
#define GB_DEV static inline
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#endif

struct CovParams {
    int kernel_id;
    double l[3];   // de-duplicated length scales
    double w[3];   // w1 (0-2), w2 (1-2), w3 (0-1)
    double amp;
};

// ------------------------------------------------------------------------------------ formulas
// The operation order follows the NumPy expressions so that results differ from the reference only
// by the last-ulp differences of exp/sin/cos implementations.
GB_DEV double k_exp_same(double D2, double g) {        // kernels.py:88
    return exp(__ddiv_rn(__dmul_rn(-0.5, D2), __dmul_rn(g, g)));
}
GB_DEV double k_exp_cross(double D2, double l1, double l2) {   // kernels.py:99
    const double s = __dadd_rn(__dmul_rn(l1, l1), __dmul_rn(l2, l2));
    return __dmul_rn(sqrt(__ddiv_rn(__dmul_rn(__dmul_rn(2.0, l1), l2), s)), exp(-__ddiv_rn(D2, s)));
}
GB_DEV double k_matern_same(double D2, double g) {     // kernels.py:145-146
    const double nu = __ddiv_rn(__dmul_rn(sqrt(3.0), sqrt(D2)), g);
    return __dmul_rn(__dadd_rn(1.0, nu), exp(-nu));
}
GB_DEV double k_matern_cross(double D2, double l1, double l2) {   // kernels.py:153-156
    const double norm = __ddiv_rn(__dmul_rn(2.0, sqrt(__dmul_rn(l1, l2))), __dsub_rn(__dmul_rn(l1, l1), __dmul_rn(l2, l2)));
    const double sd = sqrt(__dmul_rn(3.0, D2));
    const double a = __dmul_rn(l1, exp(__ddiv_rn(-sd, l1)));
    const double b = __dmul_rn(l2, exp(__ddiv_rn(-sd, l2)));
    return __dmul_rn(norm, __dsub_rn(a, b));
}
#define GB_PI 3.141592653589793
GB_DEV double k_sparse_same(double D2, double g) {     // kernels.py:108-114
    const double d = sqrt(D2);
    if (!(d < g)) return 0.0;
    const double arg = __ddiv_rn(__dmul_rn(2.0 * GB_PI, d), g);
    const double t1 = __dmul_rn(__ddiv_rn(__dadd_rn(2.0, cos(arg)), 3.0), __dsub_rn(1.0, __ddiv_rn(d, g)));
    const double t2 = __dmul_rn(1.0 / (2.0 * GB_PI), sin(arg));
    const double r = __dadd_rn(t1, t2);
    return r < 0.0 ? 0.0 : r;
}
GB_DEV double k_sparse_cross(double D2, double l1, double l2) {   // kernels.py:121-138
    const double d = sqrt(D2);
    if (l1 == l2) l2 = __dadd_rn(l2, __dmul_rn(1e-3, l2));                 // :125-126
    const double lmean = __ddiv_rn(__dadd_rn(l1, l2), 2.0);
    const double lmin = fmin(l1, l2), lmax = fmax(l1, l2);
    const double half_diff = __ddiv_rn(fabs(__dsub_rn(l2, l1)), 2.0);
    const double half_sum = __ddiv_rn(__dadd_rn(l1, l2), 2.0);
    const double c0 = __ddiv_rn(2.0, __dmul_rn(3.0, sqrt(__dmul_rn(l1, l2))));
    double res = 0.0;
    if (d >= half_diff && d <= half_sum) {                                  // branch B wins ties (:135)
        const double den = __dmul_rn(2.0 * GB_PI, __dsub_rn(__dmul_rn(l1, l1), __dmul_rn(l2, l2)));
        const double l1c = __dmul_rn(__dmul_rn(l1, l1), l1), l2c = __dmul_rn(__dmul_rn(l2, l2), l2);
        const double s1 = sin(__ddiv_rn(__dmul_rn(GB_PI, __dsub_rn(l2, __dmul_rn(2.0, d))), l1));
        const double s2 = sin(__ddiv_rn(__dmul_rn(GB_PI, __dsub_rn(l1, __dmul_rn(2.0, d))), l2));
        double v = __dsub_rn(lmean, d);
        v = __dadd_rn(v, __ddiv_rn(__dmul_rn(l1c, s1), den));
        v = __dsub_rn(v, __ddiv_rn(__dmul_rn(l2c, s2), den));
        res = __dmul_rn(c0, v);
    } else if (d <= half_diff) {                                            // branch A (:133), cosine inside the sine
        const double lmax3 = __dmul_rn(__dmul_rn(lmax, lmax), lmax);
        const double f = __ddiv_rn(__dmul_rn(1.0 / GB_PI, lmax3), __dsub_rn(__dmul_rn(lmax, lmax), __dmul_rn(lmin, lmin)));
        const double inner = cos(__ddiv_rn(__dmul_rn(2.0 * GB_PI, d), lmax));
        const double s = sin(__dmul_rn(__ddiv_rn(__dmul_rn(GB_PI, lmin), lmax), inner));
        res = __dmul_rn(c0, __dadd_rn(lmin, __dmul_rn(f, s)));
    }
    return res < 0.0 ? 0.0 : res;
}

// Block (row-block r, column-block c) of create_cov: same-property kernel on the diagonal, otherwise
// w(r,c) * cross(l_c, l_r)  (kernels.py:183-194: column strip c, vstack slot r, gammas = params[[c, r]]).
// The amplitude multiplies the finished block (inversion.py:92: gp_amp * create_cov(...)).
GB_DEV double cov_value(const CovParams& P, int r, int c, double D2) {
    double v;
    if (r == c) {
        const double g = P.l[c];
        v = P.kernel_id == GB_KERNEL_EXP ? k_exp_same(D2, g)
            : P.kernel_id == GB_KERNEL_MATERN32 ? k_matern_same(D2, g) : k_sparse_same(D2, g);
    } else {
        const double l1 = P.l[c], l2 = P.l[r];
        const int lo = min(r, c), hi = max(r, c);
        const double w = (lo == 0 && hi == 1) ? P.w[2] : (lo == 0 && hi == 2) ? P.w[0] : P.w[1];
        v = P.kernel_id == GB_KERNEL_EXP ? k_exp_cross(D2, l1, l2)
            : P.kernel_id == GB_KERNEL_MATERN32 ? k_matern_cross(D2, l1, l2) : k_sparse_cross(D2, l1, l2);
        v = __dmul_rn(w, v);
    }
    return __dmul_rn(P.amp, v);
}

// squared distance of an integer lattice offset, summed x, y, z like kernels.py:46,54-58
GB_DEV double lattice_d2(int dx, int dy, int dz, double sx, double sy, double sz) {
    const double ax = __dmul_rn((double)dx, sx), ay = __dmul_rn((double)dy, sy), az = __dmul_rn((double)dz, sz);
    return __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
}


// ------------------------------------------------------------------------------------ prism-corner potentials
#define GB_ALONG_WAY 1e6

GB_DEV double grav_corner(double x, double y, double z) {   // sensormodel.py:107-110
    const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    const double t1 = __dmul_rn(x, log(__dadd_rn(y, r)));
    const double t2 = __dmul_rn(y, log(__dadd_rn(x, r)));
    const double t3 = __dmul_rn(z, atan(__ddiv_rn(__dmul_rn(x, y), __dadd_rn(__dmul_rn(z, r), 1e-9))));
    return __dsub_rn(__dadd_rn(t1, t2), t3);
}

GB_DEV double magn_corner(double x, double y, double z, double bx, double by, double bz) {
    // sensormodel.py:127-133
    const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    const double normB = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)), __dmul_rn(bz, bz)));
    const double a1 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, by), bz), log(__dadd_rn(x, r)));
    const double a2 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, bz), bx), log(__dadd_rn(y, r)));
    const double a3 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, by), bx), log(__dadd_rn(z, r)));
    const double a4 = __dmul_rn(__dsub_rn(__dmul_rn(bz, bz), __dmul_rn(by, by)), atan(__ddiv_rn(__dmul_rn(x, z), __dmul_rn(y, r))));
    const double a5 = __dmul_rn(__dsub_rn(__dmul_rn(bz, bz), __dmul_rn(bx, bx)), atan(__ddiv_rn(__dmul_rn(y, z), __dmul_rn(x, r))));
    const double sum = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(a1, a2), a3), a4), a5);
    return -__dmul_rn(__ddiv_rn(1.0, normB), sum);
}
