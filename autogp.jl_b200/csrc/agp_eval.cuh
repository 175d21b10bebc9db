// Device interpreter for compiled kernel-tree programs: one covariance entry k(t1, t2).
//
// Arithmetic follows the reference's broadcast expressions operation by operation
// (src/GP.jl:137-140, 163-166, 199-203, 241-245, 285-289, 331-336, 375-377, 421-423,
// 481-483, 493-503; scalar forms :135-491).  This translation unit is compiled with
// -fmad=false so no multiply-add is contracted: every +,* rounds once, as in Julia where each
// broadcast statement materialises.  Remaining differences to the CPU path come only from the
// exp/sin/pow/tanh implementations (agp_math.cuh / CUDA libdevice vs openlibm); divisions by a node
// constant are correctly rounded (agp_math.cuh: div_const_v), i.e. identical to Julia's `/`.
//
// The operand stack lives in D registers (shift on push/pop) — D = 4 covers every tree with
// fewer than 16 leaves, D = 8 the rest (host guarantees need <= AGP_MAX_STACK).
#pragma once
#include "agp_math.cuh"
#include "agp_program.h"

namespace agp {

// sigma(t) = 0.5 (1 + tanh((location - t) / scale)), src/GP.jl:481-483, for E points at once
template <int E>
__device__ __forceinline__ void sigma_cp_v(const double (&t)[E], double location, double scale, double rscale, bool fast, double (&g)[E]) {
    double u[E], q[E];
#pragma unroll
    for (int e = 0; e < E; ++e) u[e] = location - t[e];
    div_const_v<E>(u, scale, rscale, fast, q);
#pragma unroll
    for (int e = 0; e < E; ++e) g[e] = 0.5 * (1.0 + slow_tanh(q[e]));
}

__device__ __forceinline__ double sigma_cp(double x, double location, double scale) {
    // src/GP.jl:481-483 (scalar twin, used by the gradient interpreter)
    return 0.5 * (1.0 + tanh((location - x) / scale));
}

// E independent entries are evaluated per interpreter pass in lock-step (agp_math.cuh): the
// instruction-level parallelism hides the latency of the FP64 chains and amortises the dispatch;
// the stack holds E values per level.
template <int D, int E>
struct RegStack {
    double s[D][E];
    __device__ __forceinline__ void push(const double (&v)[E]) {
#pragma unroll
        for (int i = D - 1; i > 0; --i)
#pragma unroll
            for (int e = 0; e < E; ++e) s[i][e] = s[i - 1][e];
#pragma unroll
        for (int e = 0; e < E; ++e) s[0][e] = v[e];
    }
    // remove the top entry into r
    __device__ __forceinline__ void pop(double (&r)[E]) {
#pragma unroll
        for (int e = 0; e < E; ++e) r[e] = s[0][e];
#pragma unroll
        for (int i = 0; i < D - 1; ++i)
#pragma unroll
            for (int e = 0; e < E; ++e) s[i][e] = s[i + 1][e];
    }
    // replace the two top entries by r
    __device__ __forceinline__ void reduce(const double (&r)[E]) {
#pragma unroll
        for (int e = 0; e < E; ++e) s[0][e] = r[e];
#pragma unroll
        for (int i = 1; i < D - 1; ++i)
#pragma unroll
            for (int e = 0; e < E; ++e) s[i][e] = s[i + 1][e];
    }
};

// (t1[e], t2[e]) = (ts[row], ts[col]) of the UPPER-triangle element; lower-triangle entries are
// the mirror image (Matrix(Symmetric(K)), src/GP.jl:501-502), so callers pass row <= col.
template <int D, int E>
__device__ __forceinline__ void eval_program(const AgpInstr* __restrict__ prog, int m, const double (&t1)[E], const double (&t2)[E], int form,
                                             double (&out)[E]) {
    RegStack<D, E> st;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int e = 0; e < E; ++e) st.s[i][e] = 0.0;
    double dx[E];
#pragma unroll
    for (int e = 0; e < E; ++e) dx[e] = t1[e] - t2[e];
    int opw = prog[0].op;
#pragma unroll 1
    for (int q = 0; q < m;) {
        // keep the per-leaf expressions of dx inside their switch case: hoisted out of the loop they
        // would all stay live across every other node (registers), for programs that may not use them
#pragma unroll
        for (int e = 0; e < E; ++e) asm volatile("" : "+d"(dx[e]));
        const int op = opw & 0xff;
        const bool fast = (opw & AGP_I_FASTDIV) != 0;
        const double a = prog[q].a, b = prog[q].b, c = prog[q].c;
        const int opw_next = (q + 1 < m) ? prog[q + 1].op : -1;
        double v[E];
        if (op >= AGP_I_PLUS) {  // binary node on the two top entries
            if (op == AGP_I_PLUS) {
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = st.s[1][e] + st.s[0][e];
            } else if (op == AGP_I_TIMES) {
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = st.s[1][e] * st.s[0][e];
            } else {  // AGP_I_CP / AGP_I_CP_SWAP; c = 1 / scale
                double g1[E], g2[E];
                sigma_cp_v<E>(t1, a, b, c, fast, g1);
                sigma_cp_v<E>(t2, a, b, c, fast, g2);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    double kl = (op == AGP_I_CP) ? st.s[1][e] : st.s[0][e];
                    double kr = (op == AGP_I_CP) ? st.s[0][e] : st.s[1][e];
                    if (form == 0) {  // vectorised: sig_1 .* k_1 + sig_2 .* k_2   (GP.jl:494-501)
                        double sig1 = g1[e] * g2[e];
                        double sig2 = (1.0 - g1[e]) * (1.0 - g2[e]);
                        v[e] = sig1 * kl + sig2 * kr;
                    } else {  // scalar: s1*k_l*s2 + (1-s1)*k_r*(1-s2)            (GP.jl:485-491)
                        v[e] = (g1[e] * kl) * g2[e] + ((1.0 - g1[e]) * kr) * (1.0 - g2[e]);
                    }
                }
            }
            st.reduce(v);
            opw = opw_next;
            ++q;
            continue;
        }
        switch (op) {
            case AGP_I_CONST:
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = a;
                break;
            case AGP_I_LINEAR:
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    double cc = (t1[e] - a) * (t2[e] - a);
                    v[e] = b + c * cc;
                }
                break;
            case AGP_I_SE: {  // amp * exp((-0.5 * dx * dx) / l^2), src/GP.jl:241-245; c = 1 / l^2
                double w[E], u[E];
#pragma unroll
                for (int e = 0; e < E; ++e) w[e] = (-0.5 * dx[e]) * dx[e];
                div_const_v<E>(w, a, c, fast, u);
                exp_v<E>(u, v);
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = b * v[e];
                break;
            }
            case AGP_I_GE: {  // amp * exp(-(|dx| / l)^gamma), src/GP.jl:285-289; d = 1 / l
                double w[E], u[E];
#pragma unroll
                for (int e = 0; e < E; ++e) w[e] = fabs(dx[e]);
                div_const_v<E>(w, a, prog[q].d, fast, u);
#pragma unroll
                for (int e = 0; e < E; ++e) w[e] = -slow_pow(u[e], b);
                exp_v<E>(w, v);
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = c * v[e];
                break;
            }
            case AGP_I_PER: {  // amp * exp((-2/l^2) * sin((pi/p) * |dx|)^2), src/GP.jl:331-336
                double w[E], u[E];
#pragma unroll
                for (int e = 0; e < E; ++e) w[e] = a * fabs(dx[e]);
                sin2_v<E>(w, u);
#pragma unroll
                for (int e = 0; e < E; ++e) w[e] = b * u[e];
                exp_v<E>(w, v);
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = c * v[e];
                break;
            }
            default:  // AGP_I_WN
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = (t1[e] == t2[e]) ? a : 0.0;
                break;
        }
        // A leaf that is the right operand of the Plus / Times that follows it combines with the stack
        // top directly (same operands, same order, same rounding as push + reduce): one interpreter
        // step and no stack traffic for the usual chains  k1 * k2 + k3 ...
        const int nop = opw_next & 0xff;
        if (opw_next >= 0 && nop == AGP_I_PLUS) {
#pragma unroll
            for (int e = 0; e < E; ++e) st.s[0][e] = st.s[0][e] + v[e];
            opw = (q + 2 < m) ? prog[q + 2].op : -1;
            q += 2;
        } else if (opw_next >= 0 && nop == AGP_I_TIMES) {
#pragma unroll
            for (int e = 0; e < E; ++e) st.s[0][e] = st.s[0][e] * v[e];
            opw = (q + 2 < m) ? prog[q + 2].op : -1;
            q += 2;
        } else {
            st.push(v);
            opw = opw_next;
            ++q;
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) out[e] = st.s[0][e];
}

// Deep trees (stack depth > 4): two entries at a time, out of line (one copy of the big-stack interpreter).
static __device__ __noinline__ void eval_pair_deep(const AgpInstr* __restrict__ prog, int m, double t1a, double t1b, double t2a, double t2b, int form,
                                            double* o0, double* o1) {
    const double a1[2] = {t1a, t1b}, a2[2] = {t2a, t2b};
    double o2[2];
    eval_program<AGP_MAX_STACK, 2>(prog, m, a1, a2, form, o2);
    *o0 = o2[0];
    *o1 = o2[1];
}

// E entries at once (E even).  Shallow programs (register stack of 2 covers e.g.
// Plus(Times(SE,Periodic),Linear) after Sethi-Ullman ordering) and programs up to stack depth 4 run
// E-wide; deeper trees fall back to pairs to bound register use.
template <int E>
__device__ __forceinline__ void eval_entries(const AgpInstr* __restrict__ prog, int m, int need, const double (&t1)[E], const double (&t2)[E],
                                             int form, double (&out)[E]) {
    if (need <= 2) {
        eval_program<2, E>(prog, m, t1, t2, form, out);
    } else if (need <= 4) {
        eval_program<4, E>(prog, m, t1, t2, form, out);
    } else {
#pragma unroll
        for (int h = 0; h < E; h += 2) {
            double o0, o1;
            eval_pair_deep(prog, m, t1[h], t1[h + 1], t2[h], t2[h + 1], form, &o0, &o1);
            out[h] = o0;
            out[h + 1] = o1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Reverse-mode derivative of one covariance entry with respect to every kernel parameter
// (vectorised form of src/GP.jl:137-503): forward sweep over the compiled program keeping every
// node value, backward sweep pushing the adjoint to the leaves, where `acc(j, d)` receives
// d = adjoint * dk/dparams[j] with j indexing the particle's params[] slice in wire order
// (Julia fieldnames order).  This is what ReverseDiff computes through eval_cov for Gen.hmc /
// Gen.map_optimize (src/inference_utils.jl:63-67, src/Greedy.jl:95, 370).
// ------------------------------------------------------------------------------------------
constexpr int AGP_GRAD_MAX_NODES = 64;

// Operand node indices of every binary node (s1 = left-hand stack operand, s0 = right-hand): a property of the
// program, so it is derived once per CTA (one thread) instead of per entry.
__device__ __forceinline__ void grad_operands(const AgpInstr* __restrict__ prog, int m, unsigned char* opa, unsigned char* opb) {
    unsigned char stack[AGP_MAX_STACK + 1];
    int sp = 0;
    for (int q = 0; q < m; ++q) {
        if ((prog[q].op & 0xff) > AGP_I_WN) {
            const int ib = stack[--sp], ia = stack[--sp];
            opa[q] = (unsigned char)ia;
            opb[q] = (unsigned char)ib;
        } else {
            opa[q] = opb[q] = 0;
        }
        stack[sp++] = (unsigned char)q;
    }
}

// E entries at once (independent chains hide the local-memory and FP64 latencies); acc(j, d) receives the SUM over
// the E entries of  seed_e * dk_e / dparams[j].
template <int E, class Acc>
__device__ __forceinline__ void eval_entries_grad(const AgpInstr* __restrict__ prog, int m, const unsigned char* __restrict__ opa,
                                                  const unsigned char* __restrict__ opb, const double (&t1)[E], const double (&t2)[E],
                                                  const double (&seed)[E], Acc&& acc) {
    double val[AGP_GRAD_MAX_NODES][E], adj[AGP_GRAD_MAX_NODES][E];
    // per-leaf intermediates of the forward sweep that the backward sweep needs again (the exponential before the
    // amplitude; sin / cos of the Periodic argument; the power of the GammaExponential; tanh of the ChangePoint)
    double ex[AGP_GRAD_MAX_NODES][E], u1s[AGP_GRAD_MAX_NODES][E], u2s[AGP_GRAD_MAX_NODES][E];
    double dx[E], adx[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        dx[e] = t1[e] - t2[e];
        adx[e] = fabs(dx[e]);
    }
    for (int q = 0; q < m; ++q) {
        const int op = prog[q].op & 0xff;
        const double a = prog[q].a, b = prog[q].b, c = prog[q].c;
        double v[E];
        switch (op) {
            case AGP_I_CONST:
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = a;
                break;
            case AGP_I_LINEAR:
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = b + c * ((t1[e] - a) * (t2[e] - a));
                break;
            case AGP_I_SE: {
                double xin[E], e1[E];
#pragma unroll
                for (int e = 0; e < E; ++e) xin[e] = ((-0.5 * dx[e]) * dx[e]) / a;
                exp_v<E>(xin, e1);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    ex[q][e] = e1[e];
                    v[e] = b * e1[e];
                }
                break;
            }
            case AGP_I_GE: {
                double xin[E], e1[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double w = pow(adx[e] / a, b);
                    u1s[q][e] = w;
                    xin[e] = -w;
                }
                exp_v<E>(xin, e1);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    ex[q][e] = e1[e];
                    v[e] = c * e1[e];
                }
                break;
            }
            case AGP_I_PER: {
                double xin[E], e1[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    double sn, cs;
                    sincos(a * adx[e], &sn, &cs);
                    u1s[q][e] = sn;
                    u2s[q][e] = cs;
                    xin[e] = b * (sn * sn);
                }
                exp_v<E>(xin, e1);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    ex[q][e] = e1[e];
                    v[e] = c * e1[e];
                }
                break;
            }
            case AGP_I_WN:
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = (t1[e] == t2[e]) ? a : 0.0;
                break;
            case AGP_I_PLUS: {
                const int ia = opa[q], ib = opb[q];
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = val[ia][e] + val[ib][e];
                break;
            }
            case AGP_I_TIMES: {
                const int ia = opa[q], ib = opb[q];
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = val[ia][e] * val[ib][e];
                break;
            }
            default: {  // ChangePoint
                const int il = (op == AGP_I_CP) ? opa[q] : opb[q], ir = (op == AGP_I_CP) ? opb[q] : opa[q];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double th1 = tanh((a - t1[e]) / b), th2 = tanh((a - t2[e]) / b);
                    u1s[q][e] = th1;
                    u2s[q][e] = th2;
                    const double g1 = 0.5 * (1.0 + th1), g2 = 0.5 * (1.0 + th2);
                    v[e] = (g1 * g2) * val[il][e] + ((1.0 - g1) * (1.0 - g2)) * val[ir][e];
                }
                break;
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) val[q][e] = v[e];
    }
    // (the program is a tree: every node has exactly one parent, which ASSIGNS the child's adjoint in the backward sweep
    // before the child is visited — no zero-initialisation, no read-modify-write of the adjoint array)
#pragma unroll
    for (int e = 0; e < E; ++e) adj[m - 1][e] = seed[e];
    for (int q = m - 1; q >= 0; --q) {
        const int op = prog[q].op & 0xff, off = prog[q].pad;
        const double a = prog[q].a, b = prog[q].b, c = prog[q].c;
        double g[E];
#pragma unroll
        for (int e = 0; e < E; ++e) g[e] = adj[q][e];
        switch (op) {
            case AGP_I_CONST: {
                double s0 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) s0 += g[e];
                acc(off, s0);
                break;
            }
            case AGP_I_WN: {
                double s0 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) s0 += (t1[e] == t2[e]) ? g[e] : 0.0;
                acc(off, s0);
                break;
            }
            case AGP_I_LINEAR: {  // bias + amp (t1 - c0)(t2 - c0): params (intercept c0, bias, amplitude)
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double u1 = t1[e] - a, u2 = t2[e] - a;
                    s0 += g[e] * (-c * (u1 + u2));
                    s1 += g[e];
                    s2 += g[e] * (u1 * u2);
                }
                acc(off, s0);
                acc(off + 1, s1);
                acc(off + 2, s2);
                break;
            }
            case AGP_I_SE: {  // amp exp(-dx^2 / (2 l^2)): params (lengthscale l, amplitude); a = l^2, d = l
                const double al = a * prog[q].d;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    s0 += g[e] * (val[q][e] * (dx[e] * dx[e]) / al);
                    s1 += g[e] * ex[q][e];
                }
                acc(off, s0);
                acc(off + 1, s1);
                break;
            }
            case AGP_I_GE: {  // amp exp(-(|dx| / l)^gamma): params (l, gamma, amp)
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double u = adx[e] / a, w = u1s[q][e];
                    s0 += g[e] * (val[q][e] * b * w / a);
                    s1 += (u > 0.0) ? g[e] * (-val[q][e] * w * log(u)) : 0.0;
                    s2 += g[e] * ex[q][e];
                }
                acc(off, s0);
                acc(off + 1, s1);
                acc(off + 2, s2);
                break;
            }
            case AGP_I_PER: {  // amp exp(b s^2), s = sin(a |dx|), a = pi / p, b = -2 / l^2: params (l, p, amp); d = l, reserved = p
                const double dbdl = -2.0 * b / prog[q].d;       // db/dl = 4 / l^3 = -2 b / l
                const double dadp = -a / prog[q].reserved;      // da/dp = -pi / p^2 = -a / p
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double sn = u1s[q][e], cs = u2s[q][e];
                    s0 += g[e] * (val[q][e] * (sn * sn) * dbdl);
                    s1 += g[e] * (val[q][e] * b * 2.0 * sn * cs * adx[e] * dadp);
                    s2 += g[e] * ex[q][e];
                }
                acc(off, s0);
                acc(off + 1, s1);
                acc(off + 2, s2);
                break;
            }
            case AGP_I_PLUS: {
                const int ia = opa[q], ib = opb[q];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    adj[ia][e] = g[e];
                    adj[ib][e] = g[e];
                }
                break;
            }
            case AGP_I_TIMES: {
                const int ia = opa[q], ib = opb[q];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double va = val[ia][e], vb = val[ib][e];
                    adj[ia][e] = g[e] * vb;
                    adj[ib][e] = g[e] * va;
                }
                break;
            }
            default: {  // ChangePoint: params (location, scale)
                const int il = (op == AGP_I_CP) ? opa[q] : opb[q], ir = (op == AGP_I_CP) ? opb[q] : opa[q];
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double kl = val[il][e], kr = val[ir][e];
                    const double u1 = (a - t1[e]) / b, u2 = (a - t2[e]) / b;
                    const double th1 = u1s[q][e], th2 = u2s[q][e];
                    const double g1 = 0.5 * (1.0 + th1), g2 = 0.5 * (1.0 + th2);
                    const double dg1 = 0.5 * (1.0 - th1 * th1) / b, dg2 = 0.5 * (1.0 - th2 * th2) / b;  // d sigma / d location
                    adj[il][e] = g[e] * (g1 * g2);
                    adj[ir][e] = g[e] * ((1.0 - g1) * (1.0 - g2));
                    const double dk1 = g2 * kl - (1.0 - g2) * kr, dk2 = g1 * kl - (1.0 - g1) * kr;   // dk / d sigma(t1), d sigma(t2)
                    s0 += g[e] * (dk1 * dg1 + dk2 * dg2);
                    s1 += g[e] * (-(dk1 * dg1 * u1 + dk2 * dg2 * u2));
                }
                acc(off, s0);
                acc(off + 1, s1);
                break;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------
// The same derivative with a LIFO tape instead of per-node arrays (round 2).  The version above keeps the value, the
// adjoint and three leaf intermediates of EVERY node in dynamically indexed per-thread arrays: 75 local-memory
// instructions per covariance entry, FP64 pipe 20 % busy (profiles/r01_ncu_launches_grad_v8.csv).  A
// postfix program is a tree, so both sweeps are stack machines:
//   forward   the operand stack of eval_program (D shift registers); a node pushes onto the TAPE only what its own
//             backward step needs: SE its exponential; GE its power and exponential; Periodic sin, cos, exponential;
//             Times its two operand values; ChangePoint its operand values and the two tanh — nothing for Constant,
//             Linear, WhiteNoise, Plus;
//   backward  the program read from its last node to its first visits every node after its parent and the subtree of the
//             operand that was evaluated LAST before the one evaluated first — the mirror image of the forward order — so
//             the adjoints live on a register stack of the same depth D (same recurrence as the Sethi-Ullman need of the
//             forward sweep) and the tape is popped in exactly the reverse order of the pushes.
// Tape traffic for Plus(Times(SE, Periodic), Linear): 6 pushes + 6 pops per entry.
// ------------------------------------------------------------------------------------------
constexpr int AGP_GRAD_TAPE = 4 * AGP_GRAD_MAX_NODES;  // a node pushes at most four values
// The hot variant's tape covers every program of up to AGP_GRAD_MAX_NODES nodes.  Larger kernels (structure learning
// with max_depth = -1 proposes them now and then) take a second instantiation with a tape of AGP_GRAD_TAPE_BIG levels;
// the host counts the pushes of every program (agp_grad_tape_need) and picks the variant per batch.
// (AGP_GRAD_TAPE_BIG: agp_program.h, shared with the host's check)

template <int D, int E, int TAPE, class Acc>
__device__ __forceinline__ void eval_program_grad(const AgpInstr* __restrict__ prog, int m, const double (&t1)[E], const double (&t2)[E],
                                                  const double (&seed)[E], Acc&& acc) {
    double tape[TAPE][E];
    int tp = 0;
    RegStack<D, E> st;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int e = 0; e < E; ++e) st.s[i][e] = 0.0;
    double dx[E], adx[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        dx[e] = t1[e] - t2[e];
        adx[e] = fabs(dx[e]);
    }
    auto tpush = [&](const double (&x)[E]) {
#pragma unroll
        for (int e = 0; e < E; ++e) tape[tp][e] = x[e];
        ++tp;
    };
    auto tpop = [&](double (&x)[E]) {
        --tp;
#pragma unroll
        for (int e = 0; e < E; ++e) x[e] = tape[tp][e];
    };
#pragma unroll 1
    for (int q = 0; q < m; ++q) {
        const int op = prog[q].op & 0xff;
        const double a = prog[q].a, b = prog[q].b, c = prog[q].c;
        double v[E];
        if (op >= AGP_I_PLUS) {
            if (op == AGP_I_PLUS) {
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = st.s[1][e] + st.s[0][e];
            } else if (op == AGP_I_TIMES) {
                tpush(st.s[1]);
                tpush(st.s[0]);
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = st.s[1][e] * st.s[0][e];
            } else {  // ChangePoint; the SWAP form has its right-hand kernel on s[1]
                double th1[E], th2[E];
                tpush(op == AGP_I_CP ? st.s[1] : st.s[0]);  // k_left
                tpush(op == AGP_I_CP ? st.s[0] : st.s[1]);  // k_right
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    th1[e] = tanh((a - t1[e]) / b);
                    th2[e] = tanh((a - t2[e]) / b);
                    const double kl = (op == AGP_I_CP) ? st.s[1][e] : st.s[0][e], kr = (op == AGP_I_CP) ? st.s[0][e] : st.s[1][e];
                    const double g1 = 0.5 * (1.0 + th1[e]), g2 = 0.5 * (1.0 + th2[e]);
                    v[e] = (g1 * g2) * kl + ((1.0 - g1) * (1.0 - g2)) * kr;
                }
                tpush(th1);
                tpush(th2);
            }
            st.reduce(v);
            continue;
        }
        switch (op) {
            case AGP_I_CONST:
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = a;
                break;
            case AGP_I_LINEAR:
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = b + c * ((t1[e] - a) * (t2[e] - a));
                break;
            case AGP_I_SE: {
                double xin[E], e1[E];
#pragma unroll
                for (int e = 0; e < E; ++e) xin[e] = ((-0.5 * dx[e]) * dx[e]) / a;
                exp_v<E>(xin, e1);
                tpush(e1);
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = b * e1[e];
                break;
            }
            case AGP_I_GE: {
                double w[E], xin[E], e1[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    w[e] = pow(adx[e] / a, b);
                    xin[e] = -w[e];
                }
                exp_v<E>(xin, e1);
                tpush(w);
                tpush(e1);
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = c * e1[e];
                break;
            }
            case AGP_I_PER: {
                double sn[E], cs[E], xin[E], e1[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    sincos(a * adx[e], &sn[e], &cs[e]);
                    xin[e] = b * (sn[e] * sn[e]);
                }
                exp_v<E>(xin, e1);
                tpush(sn);
                tpush(cs);
                tpush(e1);
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = c * e1[e];
                break;
            }
            default:  // AGP_I_WN
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = (t1[e] == t2[e]) ? a : 0.0;
                break;
        }
        st.push(v);
    }
    // backward sweep: the adjoint stack starts with the seed on the root
    RegStack<D, E> ad;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int e = 0; e < E; ++e) ad.s[i][e] = 0.0;
    ad.push(seed);
#pragma unroll 1
    for (int q = m - 1; q >= 0; --q) {
        const int op = prog[q].op & 0xff, off = prog[q].pad;
        const double a = prog[q].a, b = prog[q].b, c = prog[q].c;
        double g[E];
        ad.pop(g);
        switch (op) {
            case AGP_I_CONST: {
                double s0 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) s0 += g[e];
                acc(off, s0);
                break;
            }
            case AGP_I_WN: {
                double s0 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) s0 += (t1[e] == t2[e]) ? g[e] : 0.0;
                acc(off, s0);
                break;
            }
            case AGP_I_LINEAR: {  // bias + amp (t1 - c0)(t2 - c0): params (intercept c0, bias, amplitude)
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double u1 = t1[e] - a, u2 = t2[e] - a;
                    s0 += g[e] * (-c * (u1 + u2));
                    s1 += g[e];
                    s2 += g[e] * (u1 * u2);
                }
                acc(off, s0);
                acc(off + 1, s1);
                acc(off + 2, s2);
                break;
            }
            case AGP_I_SE: {  // amp exp(-dx^2 / (2 l^2)): params (lengthscale l, amplitude); a = l^2, d = l
                const double al = a * prog[q].d;
                double e1[E], s0 = 0.0, s1 = 0.0;
                tpop(e1);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    s0 += g[e] * ((b * e1[e]) * (dx[e] * dx[e]) / al);
                    s1 += g[e] * e1[e];
                }
                acc(off, s0);
                acc(off + 1, s1);
                break;
            }
            case AGP_I_GE: {  // amp exp(-(|dx| / l)^gamma): params (l, gamma, amp)
                double e1[E], w[E], s0 = 0.0, s1 = 0.0, s2 = 0.0;
                tpop(e1);
                tpop(w);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double u = adx[e] / a, val = c * e1[e];
                    s0 += g[e] * (val * b * w[e] / a);
                    s1 += (u > 0.0) ? g[e] * (-val * w[e] * log(u)) : 0.0;
                    s2 += g[e] * e1[e];
                }
                acc(off, s0);
                acc(off + 1, s1);
                acc(off + 2, s2);
                break;
            }
            case AGP_I_PER: {  // amp exp(b s^2), s = sin(a |dx|), a = pi / p, b = -2 / l^2: params (l, p, amp); d = l, reserved = p
                const double dbdl = -2.0 * b / prog[q].d;       // db/dl = 4 / l^3 = -2 b / l
                const double dadp = -a / prog[q].reserved;      // da/dp = -pi / p^2 = -a / p
                double e1[E], cs[E], sn[E], s0 = 0.0, s1 = 0.0, s2 = 0.0;
                tpop(e1);
                tpop(cs);
                tpop(sn);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double val = c * e1[e];
                    s0 += g[e] * (val * (sn[e] * sn[e]) * dbdl);
                    s1 += g[e] * (val * b * 2.0 * sn[e] * cs[e] * adx[e] * dadp);
                    s2 += g[e] * e1[e];
                }
                acc(off, s0);
                acc(off + 1, s1);
                acc(off + 2, s2);
                break;
            }
            case AGP_I_PLUS:
                ad.push(g);  // the operand evaluated first ...
                ad.push(g);  // ... and, on top, the operand evaluated last: its subtree is what the sweep meets next
                break;
            case AGP_I_TIMES: {
                double v0[E], v1[E], g1[E], g0[E];
                tpop(v0);
                tpop(v1);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    g1[e] = g[e] * v0[e];
                    g0[e] = g[e] * v1[e];
                }
                ad.push(g1);
                ad.push(g0);
                break;
            }
            default: {  // ChangePoint: params (location, scale)
                double th2[E], th1[E], kr[E], kl[E], gl[E], gr[E], s0 = 0.0, s1 = 0.0;
                tpop(th2);
                tpop(th1);
                tpop(kr);
                tpop(kl);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double u1 = (a - t1[e]) / b, u2 = (a - t2[e]) / b;
                    const double s1g = 0.5 * (1.0 + th1[e]), s2g = 0.5 * (1.0 + th2[e]);
                    const double dg1 = 0.5 * (1.0 - th1[e] * th1[e]) / b, dg2 = 0.5 * (1.0 - th2[e] * th2[e]) / b;  // d sigma / d location
                    gl[e] = g[e] * (s1g * s2g);
                    gr[e] = g[e] * ((1.0 - s1g) * (1.0 - s2g));
                    const double dk1 = s2g * kl[e] - (1.0 - s2g) * kr[e], dk2 = s1g * kl[e] - (1.0 - s1g) * kr[e];  // dk / d sigma(t1), d sigma(t2)
                    s0 += g[e] * (dk1 * dg1 + dk2 * dg2);
                    s1 += g[e] * (-(dk1 * dg1 * u1 + dk2 * dg2 * u2));
                }
                acc(off, s0);
                acc(off + 1, s1);
                // the operand on s[0] in the forward sweep was evaluated last: its adjoint goes on top
                if (op == AGP_I_CP) {
                    ad.push(gl);
                    ad.push(gr);
                } else {
                    ad.push(gr);
                    ad.push(gl);
                }
                break;
            }
        }
    }
}

// dispatch on the operand-stack depth the program needs, as eval_entries does
template <int E, int TAPE = AGP_GRAD_TAPE, class Acc>
__device__ __forceinline__ void eval_entries_grad_tape(const AgpInstr* __restrict__ prog, int m, int need, const double (&t1)[E],
                                                       const double (&t2)[E], const double (&seed)[E], Acc&& acc) {
    if (TAPE != AGP_GRAD_TAPE) {  // the big variant: one interpreter copy, the deepest stack
        eval_program_grad<AGP_MAX_STACK, E, TAPE>(prog, m, t1, t2, seed, acc);
        return;
    }
    if (need <= 2) eval_program_grad<2, E, TAPE>(prog, m, t1, t2, seed, acc);
    else if (need <= 4) eval_program_grad<4, E, TAPE>(prog, m, t1, t2, seed, acc);
    else eval_program_grad<AGP_MAX_STACK, E, TAPE>(prog, m, t1, t2, seed, acc);
}
}  // namespace agp
