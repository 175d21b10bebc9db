// Host build of the interpreter's FP64 elementary functions (autogp.jl_b200/csrc/agp_math.cuh)
// so the CPU suite can measure their accuracy against mpmath.  Test infrastructure only.
#define AGP_MATH_HOST 1
#include "../../autogp.jl_b200/csrc/agp_math.cuh"

extern "C" {
void agp_host_exp(const double* x, double* y, long n) {
    for (long i = 0; i + 4 <= n; i += 4) {
        double a[4] = {x[i], x[i + 1], x[i + 2], x[i + 3]}, b[4];
        agp::exp_v<4>(a, b);
        for (int e = 0; e < 4; ++e) y[i + e] = b[e];
    }
}
void agp_host_sin2(const double* x, double* y, long n) {
    for (long i = 0; i + 4 <= n; i += 4) {
        double a[4] = {x[i], x[i + 1], x[i + 2], x[i + 3]}, b[4];
        agp::sin2_v<4>(a, b);
        for (int e = 0; e < 4; ++e) y[i + e] = b[e];
    }
}
void agp_host_div(const double* x, const double* a, double* y, long n) {
    for (long i = 0; i < n; ++i) {
        double xi[2] = {x[i], x[i]}, yi[2];
        agp::div_const_v<2>(xi, a[i], 1.0 / a[i], true, yi);
        y[i] = yi[0];
    }
}
}
