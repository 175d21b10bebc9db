// Device interpreter for compiled kernel-tree programs: one covariance entry k(t1, t2).
//
// Arithmetic follows the reference's broadcast expressions operation by operation
// (src/GP.jl:137-140, 163-166, 199-203, 241-245, 285-289, 331-336, 375-377, 421-423,
// 481-483, 493-503; scalar forms :135-491).  This translation unit is compiled with
// -fmad=false so no multiply-add is contracted: every +,* rounds once, as in Julia where each
// broadcast statement materialises.  Remaining differences to the CPU path come only from the
// exp/sin/pow/tanh implementations (CUDA libdevice vs openlibm).
//
// The operand stack lives in D registers (shift on push/pop) — D = 4 covers every tree with
// fewer than 16 leaves, D = 8 the rest (host guarantees need <= AGP_MAX_STACK).
#pragma once
#include "agp_program.h"

namespace agp {

__device__ __forceinline__ double sigma_cp(double x, double location, double scale) {
    // src/GP.jl:481-483
    return 0.5 * (1.0 + tanh((location - x) / scale));
}

template <int D>
struct RegStack {
    double s[D];
    __device__ __forceinline__ void push(double v) {
#pragma unroll
        for (int i = D - 1; i > 0; --i) s[i] = s[i - 1];
        s[0] = v;
    }
    // replace the two top entries by r
    __device__ __forceinline__ void reduce(double r) {
        s[0] = r;
#pragma unroll
        for (int i = 1; i < D - 1; ++i) s[i] = s[i + 1];
    }
};

// (t1, t2) = (ts[row], ts[col]) of the UPPER-triangle element; lower-triangle entries are the
// mirror image (Matrix(Symmetric(K)), src/GP.jl:501-502), so callers pass row <= col.
template <int D>
__device__ __forceinline__ double eval_program(const AgpInstr* __restrict__ prog, int m, double t1, double t2, int form) {
    RegStack<D> st;
#pragma unroll
    for (int i = 0; i < D; ++i) st.s[i] = 0.0;
    const double dx = t1 - t2;
    const double adx = fabs(dx);
    for (int q = 0; q < m; ++q) {
        const int op = prog[q].op;
        const double a = prog[q].a, b = prog[q].b, c = prog[q].c;
        switch (op) {
            case AGP_I_CONST:
                st.push(a);
                break;
            case AGP_I_LINEAR: {
                double cc = (t1 - a) * (t2 - a);
                st.push(b + c * cc);
                break;
            }
            case AGP_I_SE: {
                double e = exp(((-0.5 * dx) * dx) / a);
                st.push(b * e);
                break;
            }
            case AGP_I_GE: {
                double e = exp(-pow(adx / a, b));
                st.push(c * e);
                break;
            }
            case AGP_I_PER: {
                double sn = sin(a * adx);
                double e = exp(b * (sn * sn));
                st.push(c * e);
                break;
            }
            case AGP_I_WN:
                st.push(t1 == t2 ? a : 0.0);
                break;
            case AGP_I_PLUS:
                st.reduce(st.s[1] + st.s[0]);
                break;
            case AGP_I_TIMES:
                st.reduce(st.s[1] * st.s[0]);
                break;
            default: {  // AGP_I_CP / AGP_I_CP_SWAP
                double kl = (op == AGP_I_CP) ? st.s[1] : st.s[0];
                double kr = (op == AGP_I_CP) ? st.s[0] : st.s[1];
                double g1 = sigma_cp(t1, a, b);
                double g2 = sigma_cp(t2, a, b);
                double r;
                if (form == 0) {  // vectorised: sig_1 .* k_1 + sig_2 .* k_2   (GP.jl:494-501)
                    double sig1 = g1 * g2;
                    double sig2 = (1.0 - g1) * (1.0 - g2);
                    r = sig1 * kl + sig2 * kr;
                } else {  // scalar: s1*k_l*s2 + (1-s1)*k_r*(1-s2)            (GP.jl:485-491)
                    r = (g1 * kl) * g2 + ((1.0 - g1) * kr) * (1.0 - g2);
                }
                st.reduce(r);
                break;
            }
        }
    }
    return st.s[0];
}

__device__ __forceinline__ double eval_entry(const AgpInstr* __restrict__ prog, int m, int need, double t1, double t2, int form) {
    if (need <= 4) return eval_program<4>(prog, m, t1, t2, form);
    return eval_program<AGP_MAX_STACK>(prog, m, t1, t2, form);
}

}  // namespace agp
