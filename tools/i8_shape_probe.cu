// Issue rate of tcgen05.mma kind::i8 (cta_group::1, M = 128, K = 32) by N and by accumulator rotation, operands resident in
// shared memory (garbage values: only the rate matters).   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/i8_shape_probe tools/i8_shape_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// mode: number of distinct A tiles / B tiles / accumulators rotated through
__global__ void __launch_bounds__(128, 1) probe(int N, int reps, int nacc, int na, int nb, long long* clocks) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    for (int w = tid; w < 196608 / 16; w += blockDim.x) reinterpret_cast<int4*>(smem)[w] = make_int4(0x01010101, 0x01010101, 0x01010101, 0x01010101);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t tmem = tmem_s;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t s0 = smem_u32(smem);
        // lean issue loop: descriptors precomputed, a k-step is one 64-bit add on the descriptor; 8 products per trip
        uint64_t da[8], db[8];
        uint32_t dd[8];
        for (int e = 0; e < 8; ++e) {
            da[e] = make_desc(s0 + (e % na) * 16384);
            db[e] = make_desc(s0 + 98304 + (e % nb) * 16384);
            dd[e] = tmem + (uint32_t)((e % nacc) * N);
        }
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) mma_i8(dd[e], da[e] + 2 * k4, db[e] + 2 * k4, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        clocks[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u));
}
int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* dC;
    cudaMalloc(&dC, sms * 8);
    const int smem = 196608 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 20000;
    struct Cfg { int N, nacc, na, nb; } cfgs[] = {{64, 1, 1, 1}, {64, 8, 1, 1}, {64, 8, 6, 6}, {128, 1, 1, 1}, {128, 4, 1, 1}, {128, 4, 6, 6}, {256, 1, 1, 1}, {256, 2, 1, 1}, {256, 2, 6, 3},
                                                  {32, 8, 1, 1}, {96, 5, 1, 1}, {16, 8, 1, 1}};
    for (auto c : cfgs) {
        probe<<<sms, 128, smem>>>(c.N, 200, c.nacc, c.na, c.nb, dC);
        cudaDeviceSynchronize();
        probe<<<sms, 128, smem>>>(c.N, reps, c.nacc, c.na, c.nb, dC);
        cudaError_t e = cudaDeviceSynchronize();
        long long clk[256];
        cudaMemcpy(clk, dC, sms * 8, cudaMemcpyDeviceToHost);
        double m = 0;
        for (int i = 0; i < sms; ++i) m += (double)clk[i] / sms;
        printf("N = %3d, %d accumulators, %d A tiles, %d B tiles: %.1f clocks per M128 N%d K32 instruction = %.0f MAC/clk/SM  (%s)\n", c.N, c.nacc, c.na, c.nb, m / (reps * 4.0), c.N,
               128.0 * c.N * 32 / (m / (reps * 4.0)), cudaGetErrorString(e));
    }
    return 0;
}
