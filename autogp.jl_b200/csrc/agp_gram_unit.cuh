// One Gram work unit: 64 rows x 128 columns of lower tile (i,k) of particle p  <-  K(ts_i, ts_k) [+ noise I]; identity in
// the padding rows.  Shared by agp_gramfill_kernel (agp_fused.cu: one CTA per unit, its own launch) and by the GRAM
// items of the persistent kernel (agp_chol_gram.cu: the same units as work-queue items, so that the kernel-tree
// evaluation overlaps the factorisation of earlier block columns).  Thread -> column, 32 rows, E entries per interpreter
// pass; rows are written as coalesced 1 KB segments.  Reference semantics: src/GP.jl:137-503, 666-668.
#pragma once
#include "agp_eval.cuh"
#include "agp_kernels.cuh"

namespace agp {

constexpr int GU_FT = 256;  // threads per unit
constexpr int GU_M = 64;    // rows of a unit
constexpr int GU_N = TB;    // columns of a unit

// ts_r / ts_c: the unit's row / column time points in shared memory; prog_s: the particle's program (shared memory when
// !LONGPROG, else ignored: the interpreter then reads v.prog with generic loads).
template <int E, bool LONGPROG>
__device__ __forceinline__ void gram_unit(const BatchView& v, int p, int i, int k, int h, const double* ts_r, const double* ts_c,
                                          const AgpInstr* prog_s, int tid) {
    constexpr int UM = GU_M, UN = GU_N;
    const bool diag = (i == k);
    const int row0 = i * TB + h * UM, col0 = k * TB;
    const int ld = v.ld;
    double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    const int poff = v.prog_off[p];
    const int pm = v.prog_off[p + 1] - poff;
    const int need = v.prog_need[p];
    const double noise = v.noise[p];
    const int n = v.n;
    const int c = tid & (UN - 1), rbase = tid >> 7;
    const int gc = col0 + c;
    const double tcol = ts_c[c];
    const bool plain = !diag && row0 + UM <= n && col0 + UN <= n;
    const bool no_kernel_rows = v.aug_identity && row0 >= v.nt * TB;
#pragma unroll 1
    for (int eb = 0; eb < 32 / E; ++eb) {
        double t1[E], t2[E], val[E];
        const int rlast = rbase + 2 * (E * eb + E - 1);
        // nothing to evaluate: the strictly-upper part of a diagonal tile (zeros), and every appended row of an
        // identity-augmented batch ([I 0]: agp_lml_grad_batch — three quarters of that matrix)
        const bool skip = (diag && c > rlast + h * UM) || no_kernel_rows;
        if (!skip) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                t1[j] = tcol;  // upper-triangle element (gc, gr): gc <= gr
                t2[j] = ts_r[rbase + 2 * (E * eb + j)];
            }
            if (LONGPROG) eval_entries<E>(v.prog + poff, pm, need, t1, t2, 0, val);
            else eval_entries<E>(prog_s, pm, need, t1, t2, 0, val);
        }
        if (plain) {
            // interior tile (all rows and columns are observations, no diagonal): nothing to decide per entry
            double* dst = Lp + (long long)(row0 + rbase + 2 * E * eb) * ld + gc;
#pragma unroll
            for (int j = 0; j < E; ++j) dst[(long long)(2 * j) * ld] = val[j];
            continue;
        }
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int r = rbase + 2 * (E * eb + j);
            const int gr = row0 + r;
            // rows/columns: [0, n) observations | [n, nt*TB) padding | [nt*TB, nt*TB + n_pred) appended
            // prediction points | padding.  Padding rows are independent unit-variance dummies
            // (identity block: log 1 = 0 in the log det, 0 in the quadratic form).
            double out = 0.0;
            if (!(diag && c > r + h * UM)) {
                const int lt = v.nt * TB;
                if (gr < n) {  // gc <= gr < n
                    out = val[j];
                    if (gr == gc) out = out + noise;  // + noise*I, src/GP.jl:667
                } else if (v.aug_identity && gr >= lt) {
                    out = (gr - lt == gc) ? 1.0 : 0.0;  // [I 0]: the appended rows solve to L^{-T}, -K^{-1}, -alpha
                } else if (gr >= lt && gr < lt + v.n_pred && (gc < n || gc >= lt)) {
                    out = val[j];  // K(t, t*) and K(t*, t*), no noise (src/GP.jl:743-747)
                } else {
                    out = (gr == gc) ? 1.0 : 0.0;
                }
            }
            Lp[(long long)gr * ld + gc] = out;
        }
    }
}

}  // namespace agp
