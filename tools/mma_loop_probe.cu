// What limits the DMMA main loop?  Variants of the 64x128 CTA tile loop of agp_fused.cu:
//   mode 0: fragments from shared memory (LDS.128) + DMMA, no barriers, no global loads
//   mode 1: + __syncthreads per K-chunk
//   mode 2: + cp.async refills from global (L2-resident buffer), full pipeline
// for 1 and 2 CTAs per SM.  Prints achieved TFLOP/s.   nvcc -O3 -arch=sm_100a
#include <cstdio>
#include <cuda.h>
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

// variant: K-chunk of 32 (two 128-byte segments per row), 2 stages, one barrier per 128 DMMAs per warp
constexpr int KC2 = 32;
constexpr int STAGE2_D = (UM + UN) * KC2;
__device__ __forceinline__ int swz2(int row, int chunk) { return row * KC2 + ((chunk ^ ((row & 1) << 2)) << 1); }  // 16 chunks per row
template <int NST>
__global__ void __launch_bounds__(FT, 2) loop_kernel_kc32(const double* __restrict__ A, int ld, int nchunk, double* out) {
    extern __shared__ __align__(128) double stages[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, c4 = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    const double* Ag = A + (size_t)(blockIdx.x % 64) * 192 * ld;
    auto load_stage = [&](int st, int chunk) {
        double* Bs = stages + st * STAGE2_D;
        double* As = Bs + UN * KC2;
        const int kk0 = chunk * KC2;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int w = tid + e * FT, row = w >> 4, ch = w & 15;
            cp_async16(Bs + swz2(row, ch), Ag + (size_t)row * ld + kk0 + ch * 2);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int w = tid + e * FT, row = w >> 4, ch = w & 15;
            cp_async16(As + swz2(row, ch), Ag + (size_t)(128 + row) * ld + kk0 + ch * 2);
        }
    };
    for (int st = 0; st < NST - 1; ++st) { load_stage(st, st); cp_commit(); }
    for (int ch = 0; ch < nchunk; ++ch) {
        cp_wait<NST - 2>();
        __syncthreads();
        const double* Bs = stages + (ch % NST) * STAGE2_D;
        const double* As = Bs + UN * KC2;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double2 a[4], b[4];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) a[mb] = *reinterpret_cast<const double2*>(As + swz2(wm * 32 + mb * 8 + g, ks * 4 + c4));
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) b[nb] = *reinterpret_cast<const double2*>(Bs + swz2(wn * 32 + nb * 8 + g, ks * 4 + c4));
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].x, b[nb].x);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].y, b[nb].y);
            if (ks == 0) {
                int nxt = ch + NST - 1;
                if (nxt < nchunk) load_stage(nxt % NST, nxt);
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

template <int NST>
void run_kc32(const char* name, const double* A, int ld, double* out, int ctas_per_sm) {
    const int nchunk = 512;
    int smem = NST * STAGE2_D * 8;
    cudaFuncSetAttribute(loop_kernel_kc32<NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    loop_kernel_kc32<NST><<<grid, FT, smem>>>(A, ld, nchunk, out);
    cudaEventRecord(e0);
    loop_kernel_kc32<NST><<<grid, FT, smem>>>(A, ld, nchunk, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * grid * (double)UM * UN * KC2 * nchunk;
    printf("%-28s ctas/sm=%d: %.3f ms  %.2f TFLOP/s  (%s)\n", name, ctas_per_sm, ms, fl / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

// ---- mode TMA: 2-D TMA tensor copies (128B hardware swizzle) + full/empty mbarriers, no CTA barrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a [rows][128 B] tile written with SWIZZLE_128B
__device__ __forceinline__ int swz128(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

template <int NST>
__global__ void __launch_bounds__(FT, 2) loop_kernel_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int nchunk, double* out) {
    extern __shared__ __align__(1024) unsigned char smem_tma[];
    constexpr int STAGE_B = (UM + UN) * 128;  // bytes: B tile 16 KB then A tile 8 KB
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_tma + NST * STAGE_B);
    uint64_t* empty = full + NST;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, c4 = lane & 3;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, FT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    __syncthreads();
    double acc[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    const int rowbase = (blockIdx.x % 64) * 192;
    auto produce = [&](int chunk) {  // thread 0 only
        const int s = chunk % NST;
        if (chunk >= NST) mbar_wait(empty + s, ((chunk / NST) - 1) & 1);
        mbar_expect_tx(full + s, STAGE_B);
        tma_load_2d(smem_tma + s * STAGE_B, &mapB, chunk * KC, rowbase, full + s);
        tma_load_2d(smem_tma + s * STAGE_B + UN * 128, &mapA, chunk * KC, rowbase + 128, full + s);
    };
    if (tid == 0)
        for (int c = 0; c < NST - 1 && c < nchunk; ++c) produce(c);
    for (int ch = 0; ch < nchunk; ++ch) {
        const int s = ch % NST;
        if (tid == 0 && ch + NST - 1 < nchunk) produce(ch + NST - 1);
        mbar_wait(full + s, (ch / NST) & 1);
        const unsigned char* Bs = smem_tma + s * STAGE_B;
        const unsigned char* As = Bs + UN * 128;
        double2 a[2][4], b[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) a[ks][mb] = *reinterpret_cast<const double2*>(As + swz128(wm * 32 + mb * 8 + g, 2 * c4 + ks));
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) b[ks][nb] = *reinterpret_cast<const double2*>(Bs + swz128(wn * 32 + nb * 8 + g, 2 * c4 + ks));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);  // this warp has read the stage
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[ks][mb].x, b[ks][nb].x);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[ks][mb].y, b[ks][nb].y);
        }
    }
    double sres = 0;
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) sres += acc[mb][nb][0] + acc[mb][nb][1];
    if (sres == 123.456) out[0] = sres;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NST>
void run_tma(const char* name, double* A, int ld, int rows, double* out, int ctas_per_sm) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    EncodeFn encode = (EncodeFn)fn;
    CUtensorMap mapA, mapB;
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t boxA[2] = {16, UM}, boxB[2] = {16, UN}, estr[2] = {1, 1};
    CUresult r1 = encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, dims, strides, boxA, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, dims, strides, boxB, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("tensor map encode failed %d %d\n", (int)r1, (int)r2); return; }
    const int nchunk = 1024;
    int smem = NST * (UM + UN) * 128 + 2 * NST * 8 + 1024;
    cudaFuncSetAttribute(loop_kernel_tma<NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    loop_kernel_tma<NST><<<grid, FT, smem>>>(mapA, mapB, nchunk, out);
    cudaEventRecord(e0);
    loop_kernel_tma<NST><<<grid, FT, smem>>>(mapA, mapB, nchunk, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * grid * (double)UM * UN * KC * nchunk;
    printf("%-28s ctas/sm=%d: %.3f ms  %.2f TFLOP/s  (%s)\n", name, ctas_per_sm, ms, fl / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
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
        run_tma<4>("TMA 2D + mbarrier, 4 stages", A, ld, 64 * 192, out, c);
    }
    return 0;
}
