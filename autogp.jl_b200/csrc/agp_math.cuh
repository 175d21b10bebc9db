// FP64 elementary functions of the kernel-tree interpreter, written for throughput on sm_100a:
//   * E independent arguments advance in lock-step, so each polynomial coefficient is fetched once
//     per step and feeds E DFMAs (libdevice's scalar exp/sin re-materialise every 64-bit constant
//     per call: more UMOVs than DFMAs in the SASS of the epilogue);
//   * no data-dependent branches on the fast path; out-of-range arguments (|x| >= 700 for exp,
//     |x| > 1e5 or non-finite for sin) fall back to libdevice;
//   * division by a per-node constant uses the host-computed correctly rounded reciprocal and two
//     exact-remainder corrections (Markstein), which returns the correctly rounded quotient — the
//     same value as Julia's `/` — without the ~30-instruction generic division.
// Accuracy (tests/test_device_math.py compiles this header for the host and compares with 50-digit
// mpmath): exp <= 1 ulp, sin/cos <= 1 ulp on the fast range, division bit-exact.
//
// The header also compiles as plain C++ (AGP_MATH_HOST) so the CPU suite can check it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(AGP_MATH_HOST)
#define AGP_MATH_FN __device__ __forceinline__
#define AGP_MATH_CONST __constant__
#else
#define AGP_MATH_FN static inline
#define AGP_MATH_CONST static const
#endif

namespace agp {

#if defined(__CUDACC__) && !defined(AGP_MATH_HOST)
__device__ __forceinline__ int f64_hi(double x) { return __double2hiint(x); }
__device__ __forceinline__ int f64_lo(double x) { return __double2loint(x); }
__device__ __forceinline__ double f64_make(int hi, int lo) { return __hiloint2double(hi, lo); }
__device__ __forceinline__ double slow_exp(double x) { return exp(x); }
__device__ __forceinline__ double slow_sin(double x) { return sin(x); }
#else
static inline int f64_hi(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline int f64_lo(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
static inline double f64_make(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x; }
static inline double slow_exp(double x) { return exp(x); }
static inline double slow_sin(double x) { return sin(x); }
#endif

// exp(r) on |r| <= ln2/2, degree 11 (highest power first; last two coefficients are 1)
AGP_MATH_CONST double kExpC[12] = {
    2.5022322536502990e-08, 2.7630903488173108e-07, 2.7557514545882439e-06, 2.4801491039099165e-05,
    1.9841269589115497e-04, 1.3888888945916380e-03, 8.3333333334550432e-03, 4.1666666666519754e-02,
    1.6666666666666477e-01, 5.0000000000000122e-01, 1.0, 1.0};

// fdlibm kernel polynomials on z = r^2, |r| <= pi/4:  sin r = r + r z S(z);  cos r = 1 - z/2 + z^2 C(z)
// row 0: S6..S1   row 1: C6..C1
AGP_MATH_CONST double kTrigC[2][6] = {
    {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04,
     8.33333333332248946124e-03, -1.66666666666666324348e-01},
    {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
     -1.38888888888741095749e-03, 4.16666666666666019037e-02}};

constexpr double kShift = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer

template <int E>
AGP_MATH_FN void exp_v(const double (&x)[E], double (&y)[E]) {
    double r[E], p[E];
    int ni[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double t = fma(x[e], 1.4426950408889634e+00, kShift);
        ni[e] = f64_lo(t);
        const double n = t - kShift;
        r[e] = fma(n, -6.9314718055994529e-01, x[e]);
        r[e] = fma(n, -2.3190468138462996e-17, r[e]);
        p[e] = kExpC[0];
    }
#pragma unroll
    for (int k = 1; k < 12; ++k)
#pragma unroll
        for (int e = 0; e < E; ++e) p[e] = fma(p[e], r[e], kExpC[k]);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        y[e] = f64_make(f64_hi(p[e]) + (ni[e] << 20), f64_lo(p[e]));  // * 2^n: result is a normal number on this range
        if ((unsigned)(f64_hi(x[e]) & 0x7fffffff) >= 0x4085e000u) y[e] = slow_exp(x[e]);  // |x| >= 700, inf, nan
    }
}

// y = sin(x)^2 is what the Periodic kernel needs (src/GP.jl:334); the quadrant sign is irrelevant.
// Returns sin(x) up to sign.
template <int E>
AGP_MATH_FN void sin_abs_v(const double (&x)[E], double (&y)[E]) {
    double r[E], z[E], p[E];
    int odd[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double t = fma(x[e], 6.3661977236758138e-01, kShift);
        odd[e] = f64_lo(t) & 1;
        const double n = t - kShift;
        r[e] = fma(n, -1.5707963267948966e+00, x[e]);
        r[e] = fma(n, -6.1232339957367660e-17, r[e]);
        r[e] = fma(n, 1.4973849048591698e-33, r[e]);
        z[e] = r[e] * r[e];
        p[e] = odd[e] ? kTrigC[1][0] : kTrigC[0][0];
    }
#pragma unroll
    for (int k = 1; k < 6; ++k)
#pragma unroll
        for (int e = 0; e < E; ++e) p[e] = fma(p[e], z[e], odd[e] ? kTrigC[1][k] : kTrigC[0][k]);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double s = fma(r[e] * z[e], p[e], r[e]);          // sin r
        const double c = fma(z[e], fma(z[e], p[e], -0.5), 1.0);  // cos r
        y[e] = odd[e] ? c : s;
        // |x| > 1e5 (three-constant reduction no longer exact), inf, nan
        if ((unsigned)(f64_hi(x[e]) & 0x7fffffff) >= 0x40f86a00u) y[e] = slow_sin(x[e]);
    }
}

// x / a with ra = 1/a correctly rounded on the host.  Exact-remainder corrections give the
// correctly rounded quotient when nothing under/overflows; `fast` (host: a normal and of moderate
// magnitude) and the exponent window on x guarantee that, otherwise the generic division runs.
AGP_MATH_FN double div_const(double x, double a, double ra, bool fast) {
    const unsigned ex = (unsigned)(f64_hi(x) & 0x7ff00000);
    if (fast && (ex - 0x1ff00000u) < 0x40000000u) {  // 2^-512 <= |x| < 2^512
        double q = x * ra;
        double rem = fma(-a, q, x);
        q = fma(rem, ra, q);
        rem = fma(-a, q, x);
        return fma(rem, ra, q);
    }
    return x / a;
}

}  // namespace agp
