// Stand-alone Gram matrix kernel (site 1 of the drop-in boundary: GP.compute_cov_matrix_vectorized /
// compute_cov_matrix / eval_cov(node, ts), src/GP.jl:61, 666-684): K(ts,ts) + noise I, column-major,
// both triangles, HBM-write bound for cheap kernel trees.  The fused LML path lives in agp_fused.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "agp_eval.cuh"
#include "agp_kernels.cuh"
#include "agp_ptx.cuh"

namespace agp {

constexpr int PROG_SMEM = 64;      // instructions cached in shared memory

// ------------------------------------------------------------------------------------------
// stand-alone Gram kernel: column-major K, both triangles; 64x64 tile pairs mirrored via smem
// ------------------------------------------------------------------------------------------
constexpr int GT = 64;
constexpr int GR_THREADS = 256;

template <bool LONGPROG>
__global__ void __launch_bounds__(GR_THREADS) agp_gram_kernel(const AgpInstr* __restrict__ prog_g, int m, int need, const double* __restrict__ ts,
                                                             int n, double noise, int form, double* __restrict__ K) {
    __shared__ double tile[GT][GT + 1];
    __shared__ double tsi[GT], tsj[GT];
    __shared__ AgpInstr prog_s[PROG_SMEM];
    // linear block id -> upper-triangular tile pair (bi <= bj)
    const int ntile = (n + GT - 1) / GT;
    int bj = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((long long)(bj + 1) * (bj + 2) / 2 <= (long long)blockIdx.x) ++bj;
    while ((long long)bj * (bj + 1) / 2 > (long long)blockIdx.x) --bj;
    const int bi = blockIdx.x - bj * (bj + 1) / 2;
    (void)ntile;
    const int tid = threadIdx.x;
    const int i0 = bi * GT, j0 = bj * GT;
    if (tid < GT) {
        tsi[tid] = (i0 + tid < n) ? ts[i0 + tid] : 0.0;
        tsj[tid] = (j0 + tid < n) ? ts[j0 + tid] : 0.0;
    }
    if (!LONGPROG) {  // the program fits the shared-memory cache: the interpreter reads it with LDS
        const double* src = reinterpret_cast<const double*>(prog_g);
        double* dst = reinterpret_cast<double*>(prog_s);
        for (int q = tid; q < m * AGP_INSTR_DOUBLES; q += GR_THREADS) dst[q] = src[q];
    }
    __syncthreads();
    const int il = tid & (GT - 1);
    const int gi = i0 + il;
#pragma unroll 1
    for (int e4 = 0; e4 < GT / 16; ++e4) {
        double t1[4], t2[4], val[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jl = (tid >> 6) + (e4 * 4 + j) * 4;
            // Symmetric(K): entry (i,j) takes the upper-triangle element (min,max)
            const bool up = gi <= j0 + jl;
            t1[j] = up ? tsi[il] : tsj[jl];
            t2[j] = up ? tsj[jl] : tsi[il];
        }
        if (LONGPROG) eval_entries<4>(prog_g, m, need, t1, t2, form, val);
        else eval_entries<4>(prog_s, m, need, t1, t2, form, val);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jl = (tid >> 6) + (e4 * 4 + j) * 4;
            const int gj = j0 + jl;
            double v = 0.0;
            if (gi < n && gj < n) {
                v = val[j];
                if (gi == gj) v = v + noise;
                K[(long long)gj * n + gi] = v;  // column j, rows contiguous
            }
            tile[jl][il] = v;
        }
    }
    if (bi != bj) {
        __syncthreads();
        // mirrored block: rows j0.., columns i0..  (element (gj, gi) = value(gi, gj))
        const int jl2 = tid & (GT - 1);
#pragma unroll 1
        for (int e = 0; e < GT / 4; ++e) {
            const int il2 = (tid >> 6) + e * 4;
            const int gi = i0 + il2, gj = j0 + jl2;
            if (gi < n && gj < n) K[(long long)gi * n + gj] = tile[jl2][il2];
        }
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
void launch_gram(const AgpInstr* prog, int m, int need, const double* ts, int n, double noise, int form, double* K, cudaStream_t s) {
    if (n <= 0) return;
    int nt = (n + GT - 1) / GT;
    int blocks = nt * (nt + 1) / 2;
    if (m <= PROG_SMEM) agp_gram_kernel<false><<<blocks, GR_THREADS, 0, s>>>(prog, m, need, ts, n, noise, form, K);
    else agp_gram_kernel<true><<<blocks, GR_THREADS, 0, s>>>(prog, m, need, ts, n, noise, form, K);
}

// ------------------------------------------------------------------------------------------
// dLML/dnoise alone (agp_lml_grad_noise_batch):  dK/dnoise = I, so
//   dLML/dnoise = 1/2 tr(alpha alpha^T - K^{-1}) = 1/2 (|alpha|^2 - |L^{-1}|_F^2).
// After the factorisation + blocked trtri of the identity-augmented matrix (no lauum pass, no kernel-tree
// walk), row lt + r of L holds row r of L^{-T} (zero left of tile column r / 128) and y[lt + r] = -alpha_r.
// One CTA sums NG_ROWS rows, one warp per row at a time, lanes along the row (coalesced); fixed summation
// order: lane-strided partial sums -> warp shuffle tree -> per-warp -> per-CTA -> second kernel over CTAs.
// ------------------------------------------------------------------------------------------
constexpr int NG_THREADS = 256, NG_ROWS = 32;

__global__ void __launch_bounds__(NG_THREADS) agp_noise_grad_kernel(BatchView v, double* __restrict__ partial) {
    __shared__ double red[NG_THREADS / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = blockIdx.y;
    const int ld = v.ld, n = v.n, lt = v.nt * TB;
    const double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    const double* __restrict__ nal = v.y + (long long)p * ld + lt;
    double acc = 0.0;
    for (int rr = warp; rr < NG_ROWS; rr += NG_THREADS / 32) {
        const int r = blockIdx.x * NG_ROWS + rr;
        if (r >= n) break;
        const double* row = Lp + (long long)(lt + r) * ld;
        double sq = 0.0;
        for (int c = (r / TB) * TB + lane; c < n; c += 32) {
            const double x = row[c];
            sq += x * x;
        }
        acc -= sq;
        if (lane == 0) {
            const double a = nal[r];
            acc += a * a;
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double sres = 0.0;
#pragma unroll
        for (int w = 0; w < NG_THREADS / 32; ++w) sres += red[w];
        partial[(long long)p * gridDim.x + blockIdx.x] = sres;
    }
}

__global__ void agp_noise_grad_reduce_kernel(const double* __restrict__ partial, int blocks, double* __restrict__ gnoise_out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double sres = 0.0;
    for (int b = 0; b < blocks; ++b) sres += partial[(long long)p * blocks + b];
    gnoise_out[p] = 0.5 * sres;
}

int noise_grad_blocks_per_particle(const BatchView& v) { return (v.n + NG_ROWS - 1) / NG_ROWS; }

void launch_noise_grad(const BatchView& v, int P, double* partial, double* gnoise_out, cudaStream_t s) {
    if (P <= 0 || v.n <= 0) return;
    const int blocks = noise_grad_blocks_per_particle(v);
    agp_noise_grad_kernel<<<dim3(blocks, P), NG_THREADS, 0, s>>>(v, partial);
    agp_noise_grad_reduce_kernel<<<P, 1, 0, s>>>(partial, blocks, gnoise_out);
}

// ------------------------------------------------------------------------------------------
// Predictive marginals (agp_predict_marginals_batch): mean and the DIAGONAL of the conditional covariance — what
// `predict`'s quantiles read (Distributions.quantile, src/GP.jl:1006-1012: mean and sqrt(diag(cov))).  m values per particle leave the GPU
// instead of m^2.
// ------------------------------------------------------------------------------------------
__global__ void agp_predict_marginals_kernel(BatchView v, const double* __restrict__ noise_pred, double* __restrict__ mean_out,
                                             double* __restrict__ var_out) {
    const int p = blockIdx.y;
    const int m = v.n_pred, o = v.nt * TB, ld = v.ld;
    const double* Lp = v.L + (long long)p * v.mat_stride;
    const double np_ = noise_pred[p];
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < m; a += gridDim.x * blockDim.x) {
        mean_out[(long long)p * m + a] = -v.y[(long long)p * ld + o + a];
        var_out[(long long)p * m + a] = Lp[(long long)(o + a) * ld + o + a] + np_;
    }
}

void launch_predict_extract_marginals(const BatchView& v, int P, const double* noise_pred, double* mean_out, double* var_out, cudaStream_t s) {
    if (P <= 0 || v.n_pred <= 0) return;
    agp_predict_marginals_kernel<<<dim3((v.n_pred + 255) / 256, P), 256, 0, s>>>(v, noise_pred, mean_out, var_out);
}

// ------------------------------------------------------------------------------------------
// Appended rows of an identity-augmented batch (agp_lml_grad_batch): row lt + r = e_r' over the observation columns,
// zeros over the trailing block, up to the end of the row's diagonal tile (only lower tiles are ever read).  Pure
// stores: 16 bytes per thread and instruction, one CTA per row — the Gram-fill kernel spends as long on these 392
// tiles per particle as on the 136 it has to evaluate.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) agp_augfill_kernel(BatchView v) {
    const int p = blockIdx.y, r = blockIdx.x;
    const int lt = v.nt * TB;
    const int ncol = lt + (r / TB + 1) * TB;  // multiple of 128
    double2* __restrict__ row = reinterpret_cast<double2*>(v.L + (long long)p * v.mat_stride + (long long)(lt + r) * v.ld);
    for (int c2 = threadIdx.x; c2 < ncol / 2; c2 += 256) {
        double2 out = make_double2(0.0, 0.0);
        if (c2 == (r >> 1)) {
            if (r & 1) out.y = 1.0;
            else out.x = 1.0;
        }
        row[c2] = out;
    }
}

void launch_augfill(const BatchView& v, int P, cudaStream_t s) {
    const int rows = (v.nt_total - v.nt) * TB;
    if (P <= 0 || rows <= 0) return;
    agp_augfill_kernel<<<dim3(rows, P), 256, 0, s>>>(v);
}

// ------------------------------------------------------------------------------------------
// Joint posterior of the summands of a sum kernel (agp_predict_sum_batch; infer_gp_sum, src/GP.jl:904-993).
// The batch was uploaded with the kernel  k_1 + ... + k_M  and (M + 1) copies of the m prediction points appended:
// appended row  g m + a  stands for F_g(t*_a) for g < M and for X(t*_a) for g = M.  agp_gramfill_kernel has filled
// every appended row with the SUM kernel; this kernel rewrites what differs (one CTA per appended row):
//   row F_g :  columns of the observations   <- k_g(t_c, t*_a)          Cov[F_g(T*), X(T)]   = Ktp[g]'   (:959-960)
//              trailing columns of F_g       <- k_g(t*_a', t*_a)         Cov[F_g(T*)]         = Kpp[g]    (:952)
//              trailing columns of F_g', g' < g  <- 0                    independent summands
//   row X*  :  trailing columns of F_g'      <- k_g'(t*_a', t*_a)        Cov[X(T*), F_g'(T*)] = Kpp[g']'  (:955-956)
// (the observation columns and the X*/X* block of row X* already hold the sum kernel, :964-967).  Entries are
// evaluated as the upper-triangle element of the joint matrix over z = [ts; ts_pred] (smaller index first), as
// compute_cov_matrix_vectorized does (:923).
// ------------------------------------------------------------------------------------------
constexpr int CF_THREADS = 256;

__device__ __forceinline__ void component_segment(const AgpInstr* __restrict__ prog, int pm, int need, const double* __restrict__ tsrc, int count,
                                                  int a, double ta, bool minmax, double* __restrict__ dst) {
    for (int j = 2 * threadIdx.x; j < count; j += 2 * CF_THREADS) {
        const int j1 = (j + 1 < count) ? j + 1 : j;
        const double u0 = tsrc[j], u1 = tsrc[j1];
        const bool f0 = minmax && j > a, f1 = minmax && j1 > a;   // the column's point comes after the row's in z
        const double t1[2] = {f0 ? ta : u0, f1 ? ta : u1};
        const double t2[2] = {f0 ? u0 : ta, f1 ? u1 : ta};
        double val[2];
        eval_entries<2>(prog, pm, need, t1, t2, 0, val);
        dst[j] = val[0];
        if (j1 != j) dst[j1] = val[1];
    }
}

__global__ void __launch_bounds__(CF_THREADS) agp_component_fill_kernel(BatchView v, ComponentView cv) {
    const int p = blockIdx.y, rp = blockIdx.x;
    const int M = cv.M, me = cv.m_each;
    const int g = rp / me, a = rp - g * me;
    const int lt = v.nt * TB, n = v.n;
    double* __restrict__ row = v.L + (long long)p * v.mat_stride + (long long)(lt + rp) * v.ld;
    const double* __restrict__ tp = v.ts + lt;  // the prediction points (first copy)
    const double ta = tp[a];
    if (g < M) {
        const int c = p * M + g;
        const AgpInstr* prog = cv.prog + cv.off[c];
        const int pm = cv.off[c + 1] - cv.off[c], need = cv.need[c];
        component_segment(prog, pm, need, v.ts, n, a, ta, false, row);                         // Cov[F_g(t*_a), X(T)]
        for (int j = threadIdx.x; j < g * me; j += CF_THREADS) row[lt + j] = 0.0;                // earlier summands: independent
        component_segment(prog, pm, need, tp, a + 1, a, ta, false, row + lt + g * me);         // Cov[F_g(t*_a), F_g(t*_a')], a' <= a
    } else {
        for (int gq = 0; gq < M; ++gq) {
            const int c = p * M + gq;
            component_segment(cv.prog + cv.off[c], cv.off[c + 1] - cv.off[c], cv.need[c], tp, me, a, ta, true, row + lt + gq * me);
        }
    }
}

void launch_component_fill(const BatchView& v, int P, const ComponentView& cv, cudaStream_t s) {
    if (P <= 0 || cv.M <= 0 || cv.m_each <= 0) return;
    agp_component_fill_kernel<<<dim3((cv.M + 1) * cv.m_each, P), CF_THREADS, 0, s>>>(v, cv);
}

}  // namespace agp
