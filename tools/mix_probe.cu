// Do DMMA (tensor subpipe) and DFMA (fp64 pipe) contend on B200?  Half the warps of each CTA
// run DMMA chains, the other half DFMA chains; compare with each running alone.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode: 0 = all warps DMMA, 1 = all DFMA, 2 = even warps DMMA / odd warps DFMA (same SMSP mix: warps w, w+4 share an SMSP)
// 3 = warps 0..7 DMMA, 8..15 DFMA
__global__ void __launch_bounds__(512) mix(double* out, int iters, int mode, double a, double b) {
    int w = threadIdx.x >> 5;
    bool do_mma = mode == 0 || (mode == 2 && ((w >> 2) & 1) == 0) || (mode == 3 && w < 8);
    bool do_fma = mode == 1 || (mode == 2 && ((w >> 2) & 1) == 1) || (mode == 3 && w >= 8);
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    if (do_mma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
    } else if (do_fma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], a, b); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* out; cudaMalloc(&out, 8 * 512 * 148);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 4; ++mode) {
        mix<<<148, 512>>>(out, iters, mode, 1.0000001, 1e-9);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        mix<<<148, 512>>>(out, iters, mode, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double nw_mma = mode == 0 ? 16 : (mode == 1 ? 0 : 8), nw_fma = mode == 1 ? 16 : (mode == 0 ? 0 : 8);
        double mma_tf = 2.0 * 256 * 16 * iters * nw_mma * 148 / ms * 1e-9;
        double fma_tf = 2.0 * 32 * 32 * iters * nw_fma * 148 / ms * 1e-9;
        printf("mode %d: %.3f ms  DMMA %.2f TF/s  DFMA %.2f TF/s\n", mode, ms, mma_tf, fma_tf);
    }
    return 0;
}
