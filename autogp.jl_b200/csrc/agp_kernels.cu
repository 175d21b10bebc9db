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

}  // namespace agp
