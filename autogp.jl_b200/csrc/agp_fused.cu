// Kernels around the persistent Cholesky kernel (agp_chol_kernel.cu): Gram fill, predictive extraction, LML gradient.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "agp_eval.cuh"
#include "agp_gram_unit.cuh"
#include "agp_kernels.cuh"
#include "agp_ptx.cuh"

namespace agp {

namespace {
constexpr int FT = 256;   // threads per CTA: 8 warps
constexpr int UM = 64;    // rows of a Gram / gradient work unit
constexpr int UN = TB;    // columns of a Gram / gradient work unit
constexpr int PROG_SMEM = 64;
}  // namespace

// ------------------------------------------------------------------------------------------
// Gram fill: L tile (i,k), i >= k  <-  K(ts_i, ts_k) [+ noise I]; identity in the padding rows.
// grid (2 * nt(nt+1)/2, P): one CTA = 64 rows x 128 columns; thread -> column, 32 rows, E
// entries per interpreter pass; rows are written as coalesced 1 KB segments.
// ------------------------------------------------------------------------------------------
#ifndef AGP_GF_E
#define AGP_GF_E 4
#endif
#ifndef AGP_GRAD_E
#define AGP_GRAD_E 2
#endif
#ifndef AGP_GRAD_ARRAYS
#define AGP_GRAD_ARRAYS 0  // 1: the round-1 reverse-mode interpreter (per-node arrays) instead of the tape form, for A/B builds
#endif
#ifndef AGP_GF_MINB
#define AGP_GF_MINB 2
#endif
constexpr int GF_E = AGP_GF_E, GF_MINB = AGP_GF_MINB;

// LONGPROG = false: every program of the batch fits the shared-memory cache (the interpreter then reads
// instructions with LDS instead of generic loads); the host picks the variant per batch.
template <int E, int MINB, bool LONGPROG>
__global__ void __launch_bounds__(FT, MINB) agp_gramfill_kernel(BatchView v, int tile_id0) {
    __shared__ __align__(128) double ts_r[UM];
    __shared__ __align__(128) double ts_c[UN];
    __shared__ AgpInstr prog_s[PROG_SMEM];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int p = blockIdx.y;
    // linear id -> lower-triangular tile (i >= k) and row half
    const int t = tile_id0 + (blockIdx.x >> 1), h = blockIdx.x & 1;
    int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    const int k = t - i * (i + 1) / 2;
    const int row0 = i * TB + h * UM, col0 = k * TB;

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (UM + UN) * 8);
        tma_bulk_g2s(ts_r, v.ts + row0, UM * 8, &bar);
        tma_bulk_g2s(ts_c, v.ts + col0, UN * 8, &bar);
    }
    const int poff = v.prog_off[p];
    const int pm = v.prog_off[p + 1] - poff;
    if (!LONGPROG) {
        const double* src = reinterpret_cast<const double*>(v.prog + poff);
        double* dst = reinterpret_cast<double*>(prog_s);
        for (int w = tid; w < pm * AGP_INSTR_DOUBLES; w += FT) dst[w] = src[w];
    }
    mbar_wait(&bar, 0);
    __syncthreads();

    gram_unit<E, LONGPROG>(v, p, i, k, h, ts_r, ts_c, prog_s, tid);
}

// ------------------------------------------------------------------------------------------
// Predictive distribution out of the augmented factorisation (src/GP.jl:731-758): with the m
// prediction points appended as extra rows, the panel solves leave L_21 = K_21 L_11^{-T}, the
// forward solve leaves y_2 = 0 - L_21 z = -K_21 K_11^{-1} xs, and the partial items leave the Schur
// complement K_22 - L_21 L_21^T in the trailing lower tiles.
//   mean[p][a]   = -y[nt*TB + a]
//   cov[p][a][b] = S[max(a,b)][min(a,b)] + noise_pred[p] * (a == b)      (symmetric by construction)
// ------------------------------------------------------------------------------------------
__global__ void agp_predict_extract_kernel(BatchView v, const double* __restrict__ noise_pred, double* __restrict__ mean_out,
                                           double* __restrict__ cov_out) {
    const int p = blockIdx.y;
    const int m = v.n_pred, o = v.nt * TB, ld = v.ld;
    const double* Lp = v.L + (long long)p * v.mat_stride;
    const double np_ = noise_pred[p];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)m * m; e += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(e / m), b = (int)(e % m);
        const int hi = a > b ? a : b, lo = a > b ? b : a;
        double s = Lp[(long long)(o + hi) * ld + o + lo];
        if (a == b) s = s + np_;
        cov_out[(long long)p * m * m + e] = s;
    }
    if (blockIdx.x == 0)
        for (int a = threadIdx.x; a < m; a += blockDim.x) mean_out[(long long)p * m + a] = -v.y[(long long)p * ld + o + a];
}

void launch_predict_extract(const BatchView& v, int P, const double* noise_pred, double* mean_out, double* cov_out, cudaStream_t s) {
    if (P <= 0 || v.n_pred <= 0) return;
    long long e = (long long)v.n_pred * v.n_pred;
    int bx = (int)((e + 255) / 256);
    if (bx > 1024) bx = 1024;
    agp_predict_extract_kernel<<<dim3(bx, P), 256, 0, s>>>(v, noise_pred, mean_out, cov_out);
}

// ------------------------------------------------------------------------------------------
// Gradient of the log marginal likelihood (Gen mvnormal logpdf_grad + ReverseDiff
// through eval_cov in the reference):   dLML/dtheta = 1/2 sum_ik A_ik dK_ik/dtheta,
// A = alpha alpha^T - K^{-1}.  After the identity-augmented factorisation the trailing lower tiles
// hold -K^{-1} and the appended forward-solve entries hold -alpha, so
//   A_ik = y[lt+i] y[lt+k] + L[lt+i][lt+k].
// One CTA = 64 rows x 128 columns of a lower tile; every thread walks its entries through the
// reverse-mode interpreter and keeps per-parameter sums; the CTA reduces them in a fixed order and
// a second kernel adds the per-CTA partials in a fixed order (bitwise reproducible, no atomics).
// ------------------------------------------------------------------------------------------
// BIG = false: the hot variant (programs of up to PROG_SMEM nodes cached in shared memory, up to AGP_GRAD_MAX_PARAMS
// parameters).  BIG = true: any program length (instructions read from global memory), a tape of AGP_GRAD_TAPE_BIG
// levels, and the parameters [j0, j0 + AGP_GRAD_MAX_PARAMS) of every kernel per launch — the host launches one window
// after the other; the noise gradient rides in window 0.
template <bool BIG>
__global__ void __launch_bounds__(FT) agp_grad_kernel(BatchView v, const int* __restrict__ param_off, double* __restrict__ partial, int j0) {
    __shared__ AgpInstr prog_s[BIG ? 1 : PROG_SMEM];
#if AGP_GRAD_ARRAYS
    __shared__ unsigned char opa_s[PROG_SMEM], opb_s[PROG_SMEM];
#endif
    __shared__ double red[FT / 32][AGP_GRAD_MAX_PARAMS + 1];
    __shared__ double ts_rs[UM], nal_rs[UM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = blockIdx.y;
    const int t = blockIdx.x >> 1, h = blockIdx.x & 1;
    int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    const int k = t - i * (i + 1) / 2;
    const int row0 = i * TB + h * UM, col0 = k * TB;
    const int ld = v.ld, n = v.n, lt = v.nt * TB;
    const double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    const double* __restrict__ nal = v.y + (long long)p * ld + lt;  // -alpha
    const int poff = v.prog_off[p];
    const int pm = v.prog_off[p + 1] - poff;
    if (!BIG) {
        const double* src = reinterpret_cast<const double*>(v.prog + poff);
        double* dst = reinterpret_cast<double*>(prog_s);
        for (int w = tid; w < pm * AGP_INSTR_DOUBLES; w += FT) dst[w] = src[w];  // host guarantees pm <= PROG_SMEM
    }
    const AgpInstr* __restrict__ prog_p = BIG ? v.prog + poff : prog_s;
    __syncthreads();
#if AGP_GRAD_ARRAYS
    if (tid == 0) grad_operands(prog_s, pm, opa_s, opb_s);
    __syncthreads();
#else
    const int need = v.prog_need[p];
#endif
    // parameters of this launch's window; slot np receives the noise gradient (BIG: window 0 only)
    const int np_all = param_off[p + 1] - param_off[p];  // !BIG: host guarantees np_all <= AGP_GRAD_MAX_PARAMS
    const int np = BIG ? max(0, min(AGP_GRAD_MAX_PARAMS, np_all - j0)) : np_all;
    if (BIG && np == 0 && j0 > 0) return;  // this kernel has no parameter in the window
    double g[AGP_GRAD_MAX_PARAMS + 1];
    for (int j = 0; j <= np; ++j) g[j] = 0.0;
    const int c = tid & (UN - 1), rbase = tid >> 7;
    const int gc = col0 + c;
    // the unit's row time points and -alpha entries: read once per CTA
    if (tid < UM) {
        const int gr = row0 + tid;
        ts_rs[tid] = (gr < n) ? v.ts[gr] : 0.0;
        nal_rs[tid] = (gr < n) ? nal[gr] : 0.0;
    }
    __syncthreads();
    if (gc < n) {
        const double tcol = v.ts[gc];
        const double nac = nal[gc];
        constexpr int GE_ = AGP_GRAD_E;  // entries per interpreter pass
        // -K^{-1} entries of a pass are loaded one pass ahead: their DRAM latency (the two most stalled instructions of
        // the first tape version, 10 % of all samples each) hides behind the interpreter
        auto load_kinv = [&](int e0, double (&kin)[GE_]) {
#pragma unroll
            for (int u = 0; u < GE_; ++u) {
                const int gr = row0 + rbase + 2 * (e0 + u);
                kin[u] = (gr < n && gc <= gr) ? Lp[(long long)(lt + gr) * ld + lt + gc] : 0.0;
            }
        };
        double kin[GE_];
        load_kinv(0, kin);
#pragma unroll 1
        for (int e0 = 0; e0 < 32; e0 += GE_) {
            double t1[GE_], t2[GE_], wgt[GE_], knext[GE_];
            if (e0 + GE_ < 32) load_kinv(e0 + GE_, knext);
            bool any = false;
#pragma unroll
            for (int u = 0; u < GE_; ++u) {
                const int r = rbase + 2 * (e0 + u), gr = row0 + r;
                const bool ok = gr < n && gc <= gr;
                t1[u] = tcol;
                t2[u] = ok ? ts_rs[r] : tcol;  // an entry outside the lower triangle runs as a (k(t,t), weight 0) dummy
                double A = 0.0;
                if (ok) {
                    A = nal_rs[r] * nac + kin[u];
                    if (gr == gc) g[np] += A;  // dK/dnoise = I
                }
                wgt[u] = ok ? ((gr == gc) ? A : 2.0 * A) : 0.0;  // off-diagonal entries count twice (symmetry)
                any = any || ok;
            }
#pragma unroll
            for (int u = 0; u < GE_; ++u) kin[u] = knext[u];
            if (!any) continue;
#if AGP_GRAD_ARRAYS
            eval_entries_grad<GE_>(prog_s, pm, opa_s, opb_s, t1, t2, wgt, [&](int j, double d) { g[j] += d; });
#else
            if (BIG) {
                eval_entries_grad_tape<GE_, AGP_GRAD_TAPE_BIG>(prog_p, pm, need, t1, t2, wgt, [&](int j, double d) {
                    const unsigned jj = (unsigned)(j - j0);
                    if (jj < (unsigned)np) g[jj] += d;
                });
            } else {
                eval_entries_grad_tape<GE_>(prog_p, pm, need, t1, t2, wgt, [&](int j, double d) { g[j] += d; });
            }
#endif
        }
    }
    for (int j = 0; j <= np; ++j) {
        const double sres = warp_sum(g[j]);
        if (lane == 0) red[warp][j] = sres;
    }
    __syncthreads();
    if (tid <= np) {
        double sres = 0.0;
#pragma unroll
        for (int w = 0; w < FT / 32; ++w) sres += red[w][tid];
        partial[((long long)p * gridDim.x + blockIdx.x) * (AGP_GRAD_MAX_PARAMS + 1) + tid] = sres;
    }
}

// window j0 of the parameters (0 for the hot variant); the noise gradient is slot np of window 0
__global__ void agp_grad_reduce_kernel(const double* __restrict__ partial, int blocks, const int* __restrict__ param_off, double* __restrict__ grad_out,
                                       double* __restrict__ gnoise_out, int j0) {
    const int p = blockIdx.x, j = threadIdx.x;
    const int np = max(0, min(AGP_GRAD_MAX_PARAMS, param_off[p + 1] - param_off[p] - j0));
    if (j > np || (j == np && j0 > 0)) return;
    double sres = 0.0;
    for (int b = 0; b < blocks; ++b) sres += partial[((long long)p * blocks + b) * (AGP_GRAD_MAX_PARAMS + 1) + j];
    if (j < np) grad_out[param_off[p] + j0 + j] = 0.5 * sres;
    else gnoise_out[p] = 0.5 * sres;
}

int grad_blocks_per_particle(const BatchView& v) { return v.nt * (v.nt + 1); }

int launch_grad(const BatchView& v, int P, const int* param_off, double* partial, double* grad_out, double* gnoise_out, int max_params, bool big,
                cudaStream_t s) {
    if (P <= 0 || v.nt <= 0) return 0;
    const int blocks = grad_blocks_per_particle(v);
    if (!big) {
        agp_grad_kernel<false><<<dim3(blocks, P), FT, 0, s>>>(v, param_off, partial, 0);
        agp_grad_reduce_kernel<<<P, AGP_GRAD_MAX_PARAMS + 1, 0, s>>>(partial, blocks, param_off, grad_out, gnoise_out, 0);
        return 2;
    }
    int launches = 0;
    for (int j0 = 0; j0 == 0 || j0 < max_params; j0 += AGP_GRAD_MAX_PARAMS, launches += 2) {  // the windows reuse `partial` in stream order
        agp_grad_kernel<true><<<dim3(blocks, P), FT, 0, s>>>(v, param_off, partial, j0);
        agp_grad_reduce_kernel<<<P, AGP_GRAD_MAX_PARAMS + 1, 0, s>>>(partial, blocks, param_off, grad_out, gnoise_out, j0);
    }
    return launches;
}

void launch_gramfill(const BatchView& v, int P, int row_tile0, cudaStream_t s) {
    if (P <= 0 || v.nt_total <= row_tile0) return;
    const int tile_id0 = row_tile0 * (row_tile0 + 1) / 2;  // lower tiles are numbered row by row
    const int tiles = v.nt_total * (v.nt_total + 1) / 2 - tile_id0;
    dim3 grid(2 * tiles, P);  // 2 row halves per tile
    // four entries per interpreter pass in lock-step (agp_math.cuh), two CTAs per SM
    if (v.max_prog_len <= PROG_SMEM) agp_gramfill_kernel<GF_E, GF_MINB, false><<<grid, FT, 0, s>>>(v, tile_id0);
    else agp_gramfill_kernel<GF_E, GF_MINB, true><<<grid, FT, 0, s>>>(v, tile_id0);
}

}  // namespace agp
