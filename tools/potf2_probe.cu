// Where do the ~55 us of one POTF2 item (128x128 diagonal tile, agp_fused.cu do_potf2) go?  A stand-alone copy of its
// phases on one CTA with clock64() stamps per phase, and variants of the serial phase (the 32x32 diagonal block factored by
// one warp in registers):
//   VAR 0  the shipped phases        VAR 1  trailing update as one 32x32 register-blocked DMMA tile per warp + unrolled store
//   VAR 2  VAR 1 + pivot column broadcast through shared memory instead of shuffles
// (a float-seeded rsqrt with two Newton steps was measured too: 12.6k clk per diagonal block against 9.7k for rsqrt())
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o potf2_probe tools/potf2_probe.cu && ./potf2_probe
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int TB = 128, FT = 256, BS = 33, BLK = 32 * BS;
__device__ __forceinline__ int blk_off(int bi, int bj) { return (bi * (bi + 1) / 2 + bj) * BLK; }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int RSQ>
__device__ __forceinline__ double inv_sqrt(double d) {
    if (RSQ == 0) return rsqrt(d);
    // float seed (rel. error 2^-22) + two Newton steps y <- y (1.5 - (d/2) y^2): 2^-44, 2^-88 -> rounding-limited
    const float f = __double2float_rn(d);
    double y = (double)rsqrtf(f);
    const double h = 0.5 * d;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    if (!(f > 1e-30f && f < 1e30f)) y = rsqrt(d);  // outside the float range (or NaN): library path
    return y;
}

template <int VAR>
__global__ void __launch_bounds__(FT) potf2_kernel(const double* __restrict__ A, double* __restrict__ Lout, double* __restrict__ dinv,
                                                  long long* stamps, int reps) {
    extern __shared__ __align__(16) double Ab[];
    __shared__ double ys[TB], Ri[TB];
    __shared__ __align__(16) double colbuf[64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = FT / 32, WORKERS = (NW - 1) * 32;
    for (int rep = 0; rep < reps; ++rep) {
        int sp = 0;
        auto stamp = [&]() { if (tid == 0 && rep == reps - 1) stamps[sp] = clock64(); ++sp; };
        __syncthreads();
        stamp();
        for (int idx = tid; idx < TB * TB; idx += FT) {
            int r = idx >> 7, c = idx & (TB - 1);
            if ((c >> 5) <= (r >> 5)) Ab[blk_off(r >> 5, c >> 5) + (r & 31) * BS + (c & 31)] = (c <= r) ? A[r * TB + c] : 0.0;
        }
        if (tid < TB) ys[tid] = 0.01 * tid;
        __syncthreads();
        stamp();  // 1: tile loaded
        if (warp == NW - 1) {
            for (int jb = 0; jb < 4; ++jb) {
                const int j0 = jb * 32;
                const double* Dg = Ab + blk_off(jb, jb);
                named_bar_sync(1, FT);
                double x[32];
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    double sacc = 0.0;
#pragma unroll
                    for (int m = 0; m < r; ++m) sacc = fma(Dg[r * BS + m], x[m], sacc);
                    const double rhs = (r == lane) ? 1.0 : 0.0;
                    x[r] = (r < lane) ? 0.0 : (rhs - sacc) * Ri[j0 + r];
                }
                double* out = dinv + jb * 1024;
#pragma unroll
                for (int r = 0; r < 32; ++r) out[r * 32 + lane] = x[r];
                if (lane == 0 && rep == reps - 1) stamps[20 + jb] = clock64();
            }
        } else {
            for (int jb = 0; jb < 4; ++jb) {
                const int j0 = jb * 32;
                double* Dg = Ab + blk_off(jb, jb);
                if (warp == 0 && VAR >= 2) {
                    // pivot column published through shared memory: one broadcast LDS.128 feeds two updates (the
                    // shuffle form needs two SHFL per update, and they all issue from this one warp)
                    double a[32];
                    const double* rowp = Dg + lane * BS;
#pragma unroll
                    for (int c = 0; c < 32; ++c) a[c] = rowp[c];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        double d = __shfl_sync(0xffffffffu, a[j], j);
                        if (!(d > 0.0)) d = 1.0;
                        const double inv = rsqrt(d);
                        const double l = (lane == j) ? d * inv : a[j] * inv;
                        a[j] = l;
                        if (lane == 0) Ri[j0 + j] = inv;
                        if (j < 31) {
                            colbuf[(j & 1) * 32 + lane] = l;
                            __syncwarp();
                            const double* cb = colbuf + (j & 1) * 32;
#pragma unroll
                            for (int c = j + 1; c < 32; ++c) {
                                if (((c & 1) == 0) && c + 1 < 32) {
                                    const double2 lc = *reinterpret_cast<const double2*>(cb + c);
                                    a[c] = fma(-l, lc.x, a[c]);
                                    a[c + 1] = fma(-l, lc.y, a[c + 1]);
                                    ++c;
                                } else {
                                    a[c] = fma(-l, cb[c], a[c]);
                                }
                            }
                        }
                    }
                    double* roww = Dg + lane * BS;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (c <= lane) roww[c] = a[c];
                } else if (warp == 0) {
                    double a[32];
                    const double* rowp = Dg + lane * BS;
#pragma unroll
                    for (int c = 0; c < 32; ++c) a[c] = rowp[c];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        double d = __shfl_sync(0xffffffffu, a[j], j);
                        if (!(d > 0.0)) d = 1.0;
                        const double inv = inv_sqrt<0>(d);
                        const double l = (lane == j) ? d * inv : a[j] * inv;
                        a[j] = l;
                        if (lane == 0) Ri[j0 + j] = inv;
#pragma unroll
                        for (int c = j + 1; c < 32; ++c) {
                            const double lc = __shfl_sync(0xffffffffu, l, c);
                            a[c] = fma(-l, lc, a[c]);
                        }
                    }
                    double* roww = Dg + lane * BS;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (c <= lane) roww[c] = a[c];
                }
                stamp();  // phase 1 done (thread 0 is in warp 0)
                named_bar_sync(1, FT);
                const int R = TB + 1 - (j0 + 32);
                if (tid >= 32 && tid - 32 < R) {
                    const int i = j0 + 32 + (tid - 32);
                    double* rowp = (i < TB) ? Ab + blk_off(i >> 5, jb) + (i & 31) * BS : ys + j0;
                    double a[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) a[c] = rowp[c];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const double l = a[j] * Ri[j0 + j];
                        a[j] = l;
#pragma unroll
                        for (int c = j + 1; c < 32; ++c) a[c] = fma(-l, Dg[c * BS + j], a[c]);
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) rowp[c] = a[c];
                }
                named_bar_sync(2, WORKERS);
                stamp();  // phase 2 done
                const int T = TB - (j0 + 32);
                if (T > 0 && VAR >= 1) {
                    // one warp per 32x32 block of the trailing lower triangle: 4x4 DMMA tiles, fragments reused across the
                    // tile row / column, sixteen independent accumulator chains
                    const int nb32 = T >> 5, nblk = nb32 * (nb32 + 1) / 2;
                    const int g = lane >> 2, c4 = lane & 3;
                    for (int blk = warp; blk < nblk; blk += NW - 1) {
                        const int bi = (blk >= 3) ? 2 : (blk >= 1 ? 1 : 0);
                        const int bc = blk - bi * (bi + 1) / 2;
                        const bool dg = (bi == bc);
                        const int Bi = jb + 1 + bi, Bc = jb + 1 + bc;  // 32-block row / column inside the tile
                        const double* ap = Ab + blk_off(Bi, jb) + g * BS + c4;
                        const double* bp = Ab + blk_off(Bc, jb) + g * BS + c4;
                        double acc[4][4][2];
#pragma unroll
                        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                            for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
#pragma unroll
                        for (int kk = 0; kk < 32; kk += 4) {
                            double a[4], b[4];
#pragma unroll
                            for (int mb = 0; mb < 4; ++mb) a[mb] = ap[mb * 8 * BS + kk];
#pragma unroll
                            for (int nb = 0; nb < 4; ++nb) b[nb] = bp[nb * 8 * BS + kk];
#pragma unroll
                            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                                for (int nb = 0; nb < 4; ++nb)
                                    if (!(dg && nb > mb)) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
                        }
                        double* cp = Ab + blk_off(Bi, Bc) + g * BS + 2 * c4;
#pragma unroll
                        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                            for (int nb = 0; nb < 4; ++nb)
                                if (!(dg && nb > mb)) {
                                    cp[mb * 8 * BS + nb * 8] -= acc[mb][nb][0];
                                    cp[mb * 8 * BS + nb * 8 + 1] -= acc[mb][nb][1];
                                }
                    }
                } else if (T > 0) {
                    const int nb8 = T >> 3, nblk = nb8 * (nb8 + 1) / 2;
                    const int g = lane >> 2, c4 = lane & 3;
                    for (int blk = warp; blk < nblk; blk += NW - 1) {
                        int bi = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
                        while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
                        while (bi * (bi + 1) / 2 > blk) --bi;
                        const int bc = blk - bi * (bi + 1) / 2;
                        const int ri = j0 + 32 + bi * 8 + g, rc = j0 + 32 + bc * 8 + g;
                        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                        const double* ap = Ab + blk_off(ri >> 5, jb) + (ri & 31) * BS + c4;
                        const double* bp = Ab + blk_off(rc >> 5, jb) + (rc & 31) * BS + c4;
#pragma unroll
                        for (int kk = 0; kk < 32; kk += 8) {
                            dmma884(c0, c1, ap[kk], bp[kk]);
                            dmma884(d0, d1, ap[kk + 4], bp[kk + 4]);
                        }
                        const int cc = j0 + 32 + bc * 8 + 2 * c4;
                        double* cp = Ab + blk_off(ri >> 5, cc >> 5) + (ri & 31) * BS + (cc & 31);
                        cp[0] -= c0 + d0;
                        cp[1] -= c1 + d1;
                    }
                }
                named_bar_sync(3, WORKERS);
                stamp();  // phase 3 done
            }
        }
        __syncthreads();
        sp = 14;
        stamp();  // 14: all warps (incl. the inverting warp) done
        if (VAR >= 1) {
#pragma unroll 1
            for (int base = 0; base < TB * TB; base += FT * 8) {
                double tmp[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = base + u * FT + tid, r = idx >> 7, c = idx & (TB - 1);
                    tmp[u] = (c <= r) ? Ab[blk_off(r >> 5, c >> 5) + (r & 31) * BS + (c & 31)] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) Lout[base + u * FT + tid] = tmp[u];
            }
        } else
        for (int idx = tid; idx < TB * TB; idx += FT) {
            int r = idx >> 7, c = idx & (TB - 1);
            Lout[idx] = (c <= r) ? Ab[blk_off(r >> 5, c >> 5) + (r & 31) * BS + (c & 31)] : 0.0;
        }
        __syncthreads();
        stamp();  // 15: stored
    }
}

template <int VAR>
void run(const double* dA, double* dL, double* dinv, long long* dst, const std::vector<double>& A) {
    const int smem = 10 * BLK * 8;
    cudaFuncSetAttribute(potf2_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaMemset(dst, 0, 32 * 8);
    potf2_kernel<VAR><<<1, FT, smem>>>(dA, dL, dinv, dst, 5);
    cudaError_t e = cudaDeviceSynchronize();
    long long st[32];
    std::vector<double> L(TB * TB);
    cudaMemcpy(st, dst, sizeof(st), cudaMemcpyDeviceToHost);
    cudaMemcpy(L.data(), dL, TB * TB * 8, cudaMemcpyDeviceToHost);
    double err = 0, nrm = 0;
    for (int r = 0; r < TB; ++r)
        for (int c = 0; c <= r; ++c) {
            double s = 0;
            for (int k = 0; k <= c; ++k) s += L[r * TB + k] * L[c * TB + k];
            err = fmax(err, fabs(s - A[r * TB + c]));
            nrm = fmax(nrm, fabs(A[r * TB + c]));
        }
    printf("VAR=%d (%s): total %lld clk (load %lld, store %lld), |LL'-A|/|A| = %.2e\n", VAR, cudaGetErrorString(e), st[15] - st[0], st[1] - st[0],
           st[15] - st[14], err / nrm);
    for (int jb = 0; jb < 4; ++jb)
        printf("   jb=%d: phase1 %lld  phase2 %lld  phase3 %lld   (inverse of block done at +%lld after its phase 1)\n", jb,
               st[2 + 3 * jb] - (jb == 0 ? st[1] : st[1 + 3 * jb]), st[3 + 3 * jb] - st[2 + 3 * jb], st[4 + 3 * jb] - st[3 + 3 * jb],
               st[20 + jb] - st[2 + 3 * jb]);
    printf("   wait for the last inverse + final barrier: %lld clk\n", st[14] - st[13]);
}

int main() {
    std::vector<double> A(TB * TB);
    for (int r = 0; r < TB; ++r)
        for (int c = 0; c < TB; ++c) {
            double dx = (r - c) / 64.0;
            A[r * TB + c] = exp(-0.5 * dx * dx / 0.09) + 0.3 * exp(-2.0 * pow(sin(3.0 * fabs(dx)), 2)) + (r == c ? 0.05 : 0.0);
        }
    double *dA, *dL, *dinv;
    long long* dst;
    cudaMalloc(&dA, TB * TB * 8); cudaMalloc(&dL, TB * TB * 8); cudaMalloc(&dinv, 4096 * 8); cudaMalloc(&dst, 32 * 8);
    cudaMemcpy(dA, A.data(), TB * TB * 8, cudaMemcpyHostToDevice);
    run<0>(dA, dL, dinv, dst, A);
    run<1>(dA, dL, dinv, dst, A);
    run<2>(dA, dL, dinv, dst, A);
    return 0;
}
