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

}  // namespace agp
