// FP64 elementary functions of the kernel-tree interpreter, written for throughput on sm_100a.
//
// Why not libdevice: an ncu capture of the Gram fill (profiles/r01_ncu_gramfill_v1.txt) showed the
// FP64 pipe only 31 % busy with the issue slots 60 % busy: 40 % of all executed instructions were
// UMOV / IMAD.MOV re-materialising the 64-bit polynomial constants of exp/sin for every call, and
// another 15 % the branches of their special-case handling, which also keep the compiler from
// interleaving independent evaluations.  Here
//   * E independent arguments advance in lock-step: each coefficient is fetched once per Horner
//     step (constant bank) and feeds E DFMAs;
//   * the fast path has no data-dependent branch (and every array index is a compile-time constant, so
//     the E-wide state stays in registers); out-of-range arguments (|x| >= 700 for exp,
//     |x| > 1e5 or non-finite for sin) are collected in one flag per call and redone by libdevice in
//     an out-of-line fallback;
//   * division by a per-node constant uses the host-computed correctly rounded reciprocal and two
//     exact-remainder corrections (Markstein), which returns the correctly rounded quotient — the
//     value Julia's `/` returns — without the generic division subroutine.
// Accuracy (tests/test_device_math.py compiles this header for the host and compares with 50-digit
// mpmath): exp <= 1 ulp, sin^2 <= 3.5 ulp on the fast range (libdevice's sin squared: ~4.5), division bit-exact.
//
// The header also compiles as plain C++ (AGP_MATH_HOST) so the CPU suite can check it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(AGP_MATH_HOST)
#define AGP_MATH_FN __device__ __forceinline__
#define AGP_MATH_SLOW static __device__ __noinline__
#define AGP_MATH_CONST __constant__
#else
#define AGP_MATH_FN static inline
#define AGP_MATH_SLOW static
#define AGP_MATH_CONST static const
#endif

namespace agp {

#if defined(__CUDACC__) && !defined(AGP_MATH_HOST)
__device__ __forceinline__ int f64_hi(double x) { return __double2hiint(x); }
__device__ __forceinline__ int f64_lo(double x) { return __double2loint(x); }
__device__ __forceinline__ double f64_make(int hi, int lo) { return __hiloint2double(hi, lo); }
#else
static inline int f64_hi(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline int f64_lo(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
static inline double f64_make(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x; }
#endif

// out-of-line fallbacks (rare): libdevice / libm
AGP_MATH_SLOW double slow_exp(double x) { return exp(x); }
AGP_MATH_SLOW double slow_sin(double x) { return sin(x); }
AGP_MATH_SLOW double slow_div(double x, double a) { return x / a; }
AGP_MATH_SLOW double slow_pow(double x, double y) { return pow(x, y); }
AGP_MATH_SLOW double slow_tanh(double x) { return tanh(x); }

// exp(r) on |r| <= ln2/2, degree 11 (highest power first; last two coefficients are 1)
AGP_MATH_CONST double kExpC[12] = {
    2.5022322536502990e-08, 2.7630903488173108e-07, 2.7557514545882439e-06, 2.4801491039099165e-05,
    1.9841269589115497e-04, 1.3888888945916380e-03, 8.3333333334550432e-03, 4.1666666666519754e-02,
    1.6666666666666477e-01, 5.0000000000000122e-01, 1.0, 1.0};

// fdlibm kernel polynomial on z = r^2, |r| <= pi/4:  sin r = r + r z S(z), S6..S1
AGP_MATH_CONST double kSinC[6] = {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                                  -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01};

#define AGP_SHIFT 6755399441055744.0  // 1.5 * 2^52: adding it rounds to the nearest integer

// argument-reduction constants, kept in the constant bank next to the coefficients (as immediates the
// compiler re-materialises each 64-bit value with two moves per use)
//   [0] log2(e)  [1] -ln2 hi  [2] -ln2 lo  [3] 2/pi  [4] -pi/2 hi  [5] -pi/2 mid  [6] +pi/2 lo
AGP_MATH_CONST double kRed[7] = {1.4426950408889634e+00, -6.9314718055994529e-01, -2.3190468138462996e-17, 6.3661977236758138e-01,
                                 -1.5707963267948966e+00, -6.1232339957367660e-17, 1.4973849048591698e-33};

template <int E>
AGP_MATH_FN void exp_v(const double (&x)[E], double (&y)[E]) {
    double r[E], p[E];
    int ni[E];
    unsigned bad = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double t = fma(x[e], kRed[0], AGP_SHIFT);
        ni[e] = f64_lo(t);
        const double n = t - AGP_SHIFT;
        r[e] = fma(n, kRed[1], x[e]);
        r[e] = fma(n, kRed[2], r[e]);
        p[e] = kExpC[0];
        bad |= (unsigned)((unsigned)(f64_hi(x[e]) & 0x7fffffff) >= 0x4085e000u);  // |x| >= 700, inf, nan
    }
#pragma unroll
    for (int k = 1; k < 12; ++k)
#pragma unroll
        for (int e = 0; e < E; ++e) p[e] = fma(p[e], r[e], kExpC[k]);
#pragma unroll
    for (int e = 0; e < E; ++e) y[e] = f64_make(f64_hi(p[e]) + (ni[e] << 20), f64_lo(p[e]));  // * 2^n: a normal number on this range
    if (bad) {
#pragma unroll
        for (int e = 0; e < E; ++e)
            if ((unsigned)(f64_hi(x[e]) & 0x7fffffff) >= 0x4085e000u) y[e] = slow_exp(x[e]);
    }
}

// y = sin(x)^2 — what the Periodic kernel needs (src/GP.jl:334).  Quadrant reduction x = n pi/2 + r,
// |r| <= pi/4; odd quadrants use sin^2(x) = cos^2(r) = 1 - sin^2(r) (no cancellation: sin^2(r) <= 1/2),
// so one polynomial serves both.
template <int E>
AGP_MATH_FN void sin2_v(const double (&x)[E], double (&y)[E]) {
    double r[E], z[E], p[E];
    int odd[E];
    unsigned bad = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double t = fma(x[e], kRed[3], AGP_SHIFT);
        odd[e] = f64_lo(t) & 1;
        const double n = t - AGP_SHIFT;
        // three-constant Cody-Waite reduction (the first step is exact for |x| <= 1e5); r carries the
        // reduced argument to half an ulp, like libdevice's sin on this range
        r[e] = fma(n, kRed[4], x[e]);
        r[e] = fma(n, kRed[5], r[e]);
        r[e] = fma(n, kRed[6], r[e]);
        z[e] = r[e] * r[e];
        p[e] = kSinC[0];
        bad |= (unsigned)((unsigned)(f64_hi(x[e]) & 0x7fffffff) >= 0x40f86a00u);  // |x| > 1e5, inf, nan
    }
#pragma unroll
    for (int k = 1; k < 6; ++k)
#pragma unroll
        for (int e = 0; e < E; ++e) p[e] = fma(p[e], z[e], kSinC[k]);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double s = fma(r[e] * z[e], p[e], r[e]);  // sin r
        const double s2 = s * s;
        y[e] = odd[e] ? 1.0 - s2 : s2;
    }
    if (bad) {
#pragma unroll
        for (int e = 0; e < E; ++e)
            if ((unsigned)(f64_hi(x[e]) & 0x7fffffff) >= 0x40f86a00u) {
                const double s = slow_sin(x[e]);
                y[e] = s * s;
            }
    }
}

// y = x / a with ra = 1/a correctly rounded on the host.  Two exact-remainder corrections give the
// correctly rounded quotient when nothing under/overflows: `fast` (host: a normal and of moderate
// magnitude) and an exponent window on x guarantee that; everything else takes the generic division.
// outside 2^-512 <= |x| < 2^512 and not zero (a zero numerator stays on the fast path: it returns +-0,
// possibly with the other sign of zero, which none of the callers can observe)
AGP_MATH_FN bool div_needs_slow(double x) {
    const unsigned hi = (unsigned)f64_hi(x);
    const unsigned ex = hi & 0x7ff00000u;
    return (ex - 0x1ff00000u) >= 0x40000000u && ((hi & 0x7fffffffu) | (unsigned)f64_lo(x)) != 0u;
}

template <int E>
AGP_MATH_FN void div_const_v(const double (&x)[E], double a, double ra, bool fast, double (&y)[E]) {
    unsigned worst = fast ? 0u : 0xffffffffu;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        double q = x[e] * ra;
        double rem = fma(-a, q, x[e]);
        q = fma(rem, ra, q);
        rem = fma(-a, q, x[e]);
        y[e] = fma(rem, ra, q);
        // exponent window as one unsigned distance; zeros trip it too and are sorted out below (rare)
        const unsigned d = ((unsigned)f64_hi(x[e]) & 0x7ff00000u) - 0x1ff00000u;
        worst = worst > d ? worst : d;
    }
    if (worst >= 0x40000000u) {
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (!fast || div_needs_slow(x[e])) y[e] = slow_div(x[e], a);
    }
}

}  // namespace agp
