/*
 * CPU ORACLE (plain C) — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no Julia in this image and
 * the reference holds no golden vectors for this path; see oracle/autogp_oracle.py header).
 *
 * Independent scalar restatement of the AutoGP.jl hot path, element by element:
 *   eval_cov(node, t1, t2)            src/GP.jl:135, 161, 194-197, 236-239, 279-283, 324-329,
 *                                     371-373, 417-419, 485-491 (form = 1), or the per-entry
 *                                     arithmetic of the vectorised methods :137-503 (form = 0)
 *   compute_cov_matrix                src/GP.jl:674-684
 *   mvnormal logpdf (mu = 0)          src/Model.jl:136 -> Distributions MvNormal logpdf:
 *                                     -(n log 2pi + logdet)/2 - |U^-T x|^2/2, K = U'U (dpotrf 'U',
 *                                     restated here as an unblocked column Cholesky)
 * Programs arrive in the wire format of include/agp_b200.h (postfix `unroll` order,
 * src/GP.jl:111-113; GPConfig codes :1101-1108 + 9 = WhiteNoise) and are evaluated with a
 * plain operand stack — deliberately a different evaluation strategy from the CUDA library.
 *
 * Build: make -C oracle   (-ffp-contract=off: no FMA contraction, like Julia's unfused code)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define MAXSTACK 256

static double eval_entry(const int32_t* ops, const int32_t* off, int32_t m, const double* prm, double t1, double t2, int form) {
    double st[MAXSTACK];
    int sp = 0;
    for (int q = 0; q < m; ++q) {
        const double* p = prm + off[q];
        switch (ops[q]) {
            case 1: st[sp++] = p[0]; break;                                    /* Constant */
            case 2: {                                                          /* Linear */
                double c = (t1 - p[0]) * (t2 - p[0]);
                st[sp++] = p[1] + p[2] * c;
                break;
            }
            case 3: {                                                          /* SquaredExponential */
                double dx = t1 - t2;
                double c = exp(-.5 * dx * dx / (p[0] * p[0]));
                st[sp++] = p[1] * c;
                break;
            }
            case 4: {                                                          /* GammaExponential */
                double dt = fabs(t1 - t2);
                double c = exp(-pow(dt / p[0], p[1]));
                st[sp++] = p[2] * c;
                break;
            }
            case 5: {                                                          /* Periodic */
                double freq = M_PI / p[1];
                double dx = fabs(t1 - t2);
                double s = sin(freq * dx);
                double c = exp((-2. / (p[0] * p[0])) * (s * s));
                st[sp++] = p[2] * c;
                break;
            }
            case 9: st[sp++] = (t1 == t2) ? p[0] : 0.0; break;                 /* WhiteNoise */
            case 6: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;             /* Plus */
            case 7: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;             /* Times */
            case 8: {                                                          /* ChangePoint */
                double kr = st[--sp], kl = st[sp - 1];
                double s1 = .5 * (1 + tanh((p[0] - t1) / p[1]));
                double s2 = .5 * (1 + tanh((p[0] - t2) / p[1]));
                double r;
                if (form == 0) {
                    double sig1 = s1 * s2, sig2 = (1 - s1) * (1 - s2);
                    r = sig1 * kl + sig2 * kr;
                } else {
                    r = s1 * kl * s2 + (1 - s1) * kr * (1 - s2);
                }
                st[sp - 1] = r;
                break;
            }
            default: return NAN;
        }
    }
    return st[0];
}

/* K (column-major n x n, both triangles) = eval_cov + noise*I.  form 0 mirrors the upper
 * triangle onto the lower like Matrix(Symmetric(K)) (GP.jl:501-502). */
void oracle_gram(const int32_t* ops, const int32_t* off, int32_t m, const double* prm, const double* ts, int32_t n, double noise, int form,
                 double* K) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            double v;
            if (form == 0 && i > j) v = eval_entry(ops, off, m, prm, ts[j], ts[i], form);
            else v = eval_entry(ops, off, m, prm, ts[i], ts[j], form);
            if (i == j) v += noise;
            K[(size_t)j * n + i] = v;
        }
}

/* "CPU-best" context baseline, NOT the reference's path: columns [j0, j1) of the upper triangle of
 * K + noise*I evaluated entry by entry with no n x n temporaries (what one fused dot-broadcast over the whole tree would
 * give; the reference makes one temporary per node, src/GP.jl:243).  The caller deals column ranges out over host
 * threads (ctypes releases the GIL; this image has no libgomp) and factors the result with LAPACK. */
void oracle_gram_upper_cols(const int32_t* ops, const int32_t* off, int32_t m, const double* prm, const double* ts, int32_t n, double noise,
                            int32_t j0, int32_t j1, double* K) {
    for (int j = j0; j < j1; ++j) {
        double* cj = K + (size_t)j * n;
        for (int i = 0; i <= j; ++i) cj[i] = eval_entry(ops, off, m, prm, ts[i], ts[j], 0);
        cj[j] += noise;
    }
}

/* Upper Cholesky in place on column-major K (only the upper triangle is read), LAPACK info. */
static int chol_upper(double* K, int n) {
    for (int j = 0; j < n; ++j) {
        double* cj = K + (size_t)j * n;
        for (int i = 0; i < j; ++i) {   /* U(i,j) = (K(i,j) - sum_{k<i} U(k,i) U(k,j)) / U(i,i) */
            const double* ci = K + (size_t)i * n;
            double s = cj[i];
            for (int k = 0; k < i; ++k) s -= ci[k] * cj[k];
            cj[i] = s / ci[i];
        }
        double d = cj[j];
        for (int k = 0; k < j; ++k) d -= cj[k] * cj[k];
        if (!(d > 0.0)) return j + 1;
        cj[j] = sqrt(d);
    }
    return 0;
}

/* returns info; *lml = logpdf(mvnormal(0, K + noise I), xs) */
int oracle_lml(const int32_t* ops, const int32_t* off, int32_t m, const double* prm, const double* ts, const double* xs, int32_t n, double noise,
               double* lml) {
    if (n == 0) { *lml = 0.0; return 0; }
    double* K = (double*)malloc((size_t)n * n * sizeof(double));
    double* z = (double*)malloc((size_t)n * sizeof(double));
    oracle_gram(ops, off, m, prm, ts, n, noise, 0, K);
    int info = chol_upper(K, n);
    if (info == 0) {
        double logdet = 0.0, q = 0.0;
        for (int j = 0; j < n; ++j) {   /* z = U^-T x: forward substitution on U' */
            const double* cj = K + (size_t)j * n;
            double s = xs[j];
            for (int k = 0; k < j; ++k) s -= cj[k] * z[k];
            z[j] = s / cj[j];
            q += z[j] * z[j];
            logdet += log(cj[j]);
        }
        *lml = -0.5 * ((double)n * log(2.0 * M_PI) + 2.0 * logdet) - 0.5 * q;
    } else {
        *lml = NAN;
    }
    free(K);
    free(z);
    return info;
}
