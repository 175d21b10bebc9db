// What limits the DMMA main loop?  Variants of the 64x128 CTA tile loop of agp_fused.cu:
//   mode 0: fragments from shared memory (LDS.128) + DMMA, no barriers, no global loads
//   mode 1: + __syncthreads per K-chunk
//   mode 2: + cp.async refills from global (L2-resident buffer), full pipeline
// for 1 and 2 CTAs per SM.  Prints achieved TFLOP/s.   nvcc -O3 -arch=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(s)), "l"(g)); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

constexpr int UM = 64, UN = 128, KC = 16, NSTAGE = 4, FT = 256;
constexpr int STAGE_D = (UM + UN) * KC;
__device__ __forceinline__ int swz(int row, int chunk) { return row * KC + ((chunk ^ ((row & 1) << 2)) << 1); }

template <int MODE>
__global__ void __launch_bounds__(FT, 2) loop_kernel(const double* __restrict__ A, int ld, int nchunk, double* out) {
    extern __shared__ __align__(128) double stages[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, c4 = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    for (int w = tid; w < NSTAGE * STAGE_D; w += FT) stages[w] = 1e-3 * (w & 7);
    __syncthreads();
    const double* Ag = A + (size_t)(blockIdx.x % 64) * 192 * ld;
    auto load_stage = [&](int st, int chunk) {
        double* Bs = stages + st * STAGE_D;
        double* As = Bs + UN * KC;
        const int kk0 = chunk * KC;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int w = tid + e * FT, row = w >> 3, ch = w & 7;
            cp_async16(Bs + swz(row, ch), Ag + (size_t)row * ld + kk0 + ch * 2);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int w = tid + e * FT, row = w >> 3, ch = w & 7;
            cp_async16(As + swz(row, ch), Ag + (size_t)(128 + row) * ld + kk0 + ch * 2);
        }
    };
    if (MODE == 2) {
        for (int st = 0; st < NSTAGE - 1; ++st) { load_stage(st, st); cp_commit(); }
    }
    for (int ch = 0; ch < nchunk; ++ch) {
        if (MODE == 2) cp_wait<NSTAGE - 2>();
        if (MODE >= 1) __syncthreads();
        const double* Bs = stages + (ch % NSTAGE) * STAGE_D;
        const double* As = Bs + UN * KC;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            double2 a[4], b[4];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) a[mb] = *reinterpret_cast<const double2*>(As + swz(wm * 32 + mb * 8 + g, ks * 4 + c4));
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) b[nb] = *reinterpret_cast<const double2*>(Bs + swz(wn * 32 + nb * 8 + g, ks * 4 + c4));
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].x, b[nb].x);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].y, b[nb].y);
            if (MODE == 2 && ks == 0) {
                int nxt = ch + NSTAGE - 1;
                if (nxt < nchunk) load_stage(nxt % NSTAGE, nxt);
                cp_commit();
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) s += acc[mb][nb][0] + acc[mb][nb][1];
    if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, const double* A, int ld, double* out, int ctas_per_sm) {
    const int nchunk = 1024;  // K = 16384 -> but we wrap reads within ld
    int smem = NSTAGE * STAGE_D * 8;
    cudaFuncSetAttribute(loop_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    loop_kernel<MODE><<<grid, FT, smem>>>(A, ld, nchunk, out);
    cudaEventRecord(e0);
    loop_kernel<MODE><<<grid, FT, smem>>>(A, ld, nchunk, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * grid * (double)UM * UN * KC * nchunk;
    printf("%-28s ctas/sm=%d: %.3f ms  %.2f TFLOP/s  (%s)\n", name, ctas_per_sm, ms, fl / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int ld = 16384 + 64;
    double* A; cudaMalloc(&A, (size_t)64 * 192 * ld * 8);
    cudaMemset(A, 0, (size_t)64 * 192 * ld * 8);
    double* out; cudaMalloc(&out, 8);
    for (int c = 1; c <= 2; ++c) {
        run<0>("LDS+DMMA", A, ld, out, c);
        run<1>("LDS+DMMA+barrier", A, ld, out, c);
        run<2>("LDS+DMMA+barrier+cp.async", A, ld, out, c);
    }
    return 0;
}
