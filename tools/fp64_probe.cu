// FP64 issue-rate probe for B200 (sm_100a): DFMA vs mma.sync DMMA vs libdevice transcendentals.
// Measures the FP64 roofline denominator that MEASURED_PEAKS.json lacks.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8 f64: A 4 regs, B 2 regs, C 4 regs
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double av, double bv) {
    double c[NACC][4];
    double a[4] = {av, av + 1, av + 2, av + 3};
    double b[2] = {bv, bv + 1};
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
__global__ void __launch_bounds__(256) transc_kernel(double* out, int iters, double x0, double g) {
    double x = x0 + 1e-3 * threadIdx.x;
    double s = 0;
    for (int it = 0; it < iters; ++it) {
        double v;
        if (OP == 0) v = exp(-x);
        else if (OP == 1) v = sin(x * 37.0);
        else if (OP == 2) v = pow(x, g);
        else if (OP == 3) v = tanh(x - 0.5);
        else if (OP == 4) v = sqrt(x);
        else v = 1.0 / x;
        s += v;
        x += 1e-7;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();  // warm
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * 256 * sms * 16));
    const int iters = 20000;
    for (int cps = 1; cps <= 8; cps *= 2) {
        int grid = sms * cps;
        {
            float ms = time_ms([&] { dfma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * iters * 256.0 * grid;
            printf("DFMA ilp16 ctas/sm=%d: %.3f ms %.2f TFLOP/s\n", cps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma884_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 16 * iters * 8.0 * grid;
            printf("DMMA m8n8k4 acc16 ctas/sm=%d: %.3f ms %.2f TFLOP/s\n", cps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma1688_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * 8 * 8 * 8 * iters * 8.0 * grid;
            printf("DMMA m16n8k8 acc8 ctas/sm=%d: %.3f ms %.2f TFLOP/s\n", cps, ms, fl / ms * 1e-9);
        }
    }
    // small-occupancy cases: 8 warps/SM with 32 accumulators (as in a 64x32 warp tile)
    {
        float ms = time_ms([&] { dmma884_kernel<32><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 256 * 32 * iters * 8.0 * sms;
        printf("DMMA m8n8k4 acc32 1cta/sm 8 warps: %.3f ms %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { dmma884_kernel<32><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 256 * 32 * iters * 4.0 * sms;
        printf("DMMA m8n8k4 acc32 1cta/sm 4 warps: %.3f ms %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { dfma_kernel<64><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 64 * iters * 256.0 * sms;
        printf("DFMA ilp64 1cta/sm 8 warps: %.3f ms %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    const char* names[] = {"exp", "sin", "pow", "tanh", "sqrt", "rcp"};
    const int titers = 2000;
    int grid = sms * 8;
    float tms[6];
    tms[0] = time_ms([&] { transc_kernel<0><<<grid, 256>>>(out, titers, 0.3, 1.3); }, 3);
    tms[1] = time_ms([&] { transc_kernel<1><<<grid, 256>>>(out, titers, 0.3, 1.3); }, 3);
    tms[2] = time_ms([&] { transc_kernel<2><<<grid, 256>>>(out, titers, 0.3, 1.3); }, 3);
    tms[3] = time_ms([&] { transc_kernel<3><<<grid, 256>>>(out, titers, 0.3, 1.3); }, 3);
    tms[4] = time_ms([&] { transc_kernel<4><<<grid, 256>>>(out, titers, 0.3, 1.3); }, 3);
    tms[5] = time_ms([&] { transc_kernel<5><<<grid, 256>>>(out, titers, 0.3, 1.3); }, 3);
    for (int i = 0; i < 6; ++i) {
        double evals = (double)titers * 256.0 * grid;
        printf("f64 %s: %.3f ms %.2f Geval/s\n", names[i], tms[i], evals / tms[i] * 1e-6);
    }
    // HBM write bandwidth (Gram stage bound): 4 GiB fill
    {
        size_t bytes = (size_t)4 << 30;
        double* buf; CK(cudaMalloc(&buf, bytes));
        float ms = time_ms([&] { CK(cudaMemsetAsync(buf, 0, bytes)); }, 5);
        printf("memset 4GiB: %.3f ms %.1f GB/s\n", ms, bytes / ms * 1e-6);
        CK(cudaFree(buf));
    }
    return 0;
}
