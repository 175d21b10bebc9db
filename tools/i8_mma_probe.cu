// Probe for DESIGN.md §9(c): the int8 tensor path of sm_100a (tcgen05.mma kind::i8, TMEM accumulators) that an
// Ozaki-style error-free split of the FP64 trailing update would run on (tools/ozaki_accuracy.py: 8 slices = 36 int8
// products per FP64 product give FP64-grade results on the benchmark factor).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/i8_mma_probe tools/i8_mma_probe.cu
//   tools/i8_mma_probe            (GPU box)
//
// (1) correctness of the hand-built descriptors: D (int32, 128 x 128, TMEM) = A (int8, 128 x 128, K-major, 128-byte
//     swizzle) x B^T against the host;  (2) issue rate with the operands RESIDENT in shared memory: one CTA per SM, one
//     thread issues M128 N128 K32 instructions back to back — the ceiling any slice pipeline could reach;  (3) the same
//     with every 128-deep product's two 16 KB operand tiles streamed from an L2-resident buffer by 1-D bulk copies through
//     a ring of stages — what the slice stream (s bytes per FP64 element and operand) costs next to the tensor time.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major operand tile, 128 rows x 128 bytes, SWIZZLE_128B: 8-row groups of 1024 bytes (stride byte offset), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::i8: D = S32 (c_format 2), A and B signed 8 bit, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int TILE = 128 * 128;  // bytes of one operand tile

// mode 0: one 128-deep product, D written to `out`.  mode 1: `reps` x 4 instructions on resident operands.
// mode 2: `reps` 128-deep products, each with fresh A and B tiles streamed from `stream` (a buffer of `ntiles` tiles
// that stays in L2) through a ring of NST stages.
constexpr int NST = 4;
__global__ void __launch_bounds__(128, 1) probe(const int8_t* __restrict__ a_img, const int8_t* __restrict__ b_img, int32_t* __restrict__ out,
                                                const int8_t* __restrict__ stream, int ntiles, int mode, int reps, long long* clocks) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar_mma, full[NST], empty[NST];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    int8_t* As = reinterpret_cast<int8_t*>(smem);  // stage st: A at st * 2 TILE, B behind it
    if (tid == 0) {
        mbar_init(&bar_mma, 1);
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    // resident operands (modes 0, 1): host-swizzled images copied with ordinary stores, then handed to the async proxy
    if (mode != 2) {
        for (int w = tid; w < TILE / 16; w += blockDim.x) {
            reinterpret_cast<int4*>(As)[w] = reinterpret_cast<const int4*>(a_img)[w];
            reinterpret_cast<int4*>(As + TILE)[w] = reinterpret_cast<const int4*>(b_img)[w];
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t tmem = tmem_base_s;

    long long t0 = 0, t1 = 0;
    if (mode != 2) {
        if (tid == 0) {
            const uint32_t sa = smem_u32(As), sb = smem_u32(As + TILE);
            t0 = clock64();
            for (int r = 0; r < reps; ++r)
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) mma_i8(tmem, make_desc(sa + k4 * 32), make_desc(sb + k4 * 32), (r | k4) != 0 ? 1u : 0u);
            mma_commit(&bar_mma);
            mbar_wait(&bar_mma, 0);
            t1 = clock64();
        }
    } else {
        // producer (warp 1, one thread) and MMA issuer (warp 0, one thread)
        if (tid == 32) {
            for (int r = 0; r < reps; ++r) {
                const int st = r % NST;
                if (r >= NST) mbar_wait(&empty[st], ((r / NST) - 1) & 1);
                mbar_expect_tx(&full[st], 2 * TILE);
                const size_t t = ((size_t)r * 2 + (size_t)blockIdx.x * 5) % (size_t)(ntiles - 1);
                bulk_g2s(As + (size_t)st * 2 * TILE, stream + t * TILE, 2 * TILE, &full[st]);
            }
        } else if (tid == 0) {
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                const int st = r % NST;
                mbar_wait(&full[st], (r / NST) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                const uint32_t sa = smem_u32(As + (size_t)st * 2 * TILE), sb = sa + TILE;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) mma_i8(tmem, make_desc(sa + k4 * 32), make_desc(sb + k4 * 32), (r | k4) != 0 ? 1u : 0u);
                mma_commit(&empty[st]);  // the stage is free once these four instructions have read it
            }
            mma_commit(&bar_mma);
            mbar_wait(&bar_mma, 0);
            t1 = clock64();
        }
    }
    if (tid == 0 && clocks) clocks[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    if (mode == 0) {
        // thread t of warp w reads lane 32 w + t (= row of D), 32 columns at a time
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                  "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                  "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                  "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 32; ++j) out[(size_t)tid * 128 + c0 + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128u));
}

// host image of a [128 rows][128 bytes] K-major tile under the 128-byte swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
static void swizzle_tile(const int8_t* src, int8_t* dst) {
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 8; ++c)
            for (int b = 0; b < 16; ++b) dst[r * 128 + ((c ^ (r & 7)) << 4) + b] = src[r * 128 + c * 16 + b];
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    std::vector<int8_t> A(TILE), B(TILE), Ai(TILE), Bi(TILE);
    srand(7);
    for (int i = 0; i < TILE; ++i) {
        A[i] = (int8_t)(rand() % 255 - 127);
        B[i] = (int8_t)(rand() % 255 - 127);
    }
    swizzle_tile(A.data(), Ai.data());
    swizzle_tile(B.data(), Bi.data());
    int8_t *dA, *dB, *dS;
    int32_t* dD;
    long long* dC;
    const int ntiles = 2048;  // 32 MB of operand tiles: stays in the 126 MB L2
    CK(cudaMalloc(&dA, TILE));
    CK(cudaMalloc(&dB, TILE));
    CK(cudaMalloc(&dD, 128 * 128 * 4));
    CK(cudaMalloc(&dS, (size_t)ntiles * TILE));
    CK(cudaMalloc(&dC, sms * sizeof(long long)));
    CK(cudaMemcpy(dA, Ai.data(), TILE, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bi.data(), TILE, cudaMemcpyHostToDevice));
    CK(cudaMemset(dS, 1, (size_t)ntiles * TILE));
    const int smem = NST * 2 * TILE + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));

    // (1) correctness
    probe<<<1, 128, smem>>>(dA, dB, dD, dS, ntiles, 0, 1, nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> D(128 * 128);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 128; ++j) {
            int32_t ref = 0;
            for (int k = 0; k < 128; ++k) ref += (int32_t)A[i * 128 + k] * (int32_t)B[j * 128 + k];
            bad += (ref != D[i * 128 + j]);
        }
    printf("(1) D = A B^T, int8 x int8 -> int32, M128 N128 K128 through tcgen05.mma kind::i8: %lld of 16384 entries differ from the host\n", bad);

    // (2), (3) rates
    const double mac_per_instr = 128.0 * 128 * 32;
    for (int mode = 1; mode <= 2; ++mode) {
        const int reps = 20000;
        probe<<<sms, 128, smem>>>(dA, dB, dD, dS, ntiles, mode, 200, dC);  // warm
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        probe<<<sms, 128, smem>>>(dA, dB, dD, dS, ntiles, mode, reps, dC);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> clk(sms);
        CK(cudaMemcpy(clk.data(), dC, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        double cmean = 0;
        for (int i = 0; i < sms; ++i) cmean += (double)clk[i] / sms;
        const double instr = (double)reps * 4;
        const double tops = 2.0 * mac_per_instr * instr * sms / (ms * 1e-3) * 1e-12;
        printf("(%d) %s: %d SMs x %d products of 128^3: %.3f ms, %.1f clocks per M128 N128 K32 instruction (%.0f MAC/clk/SM), %.0f TOP/s",
               mode + 1, mode == 1 ? "operands resident in shared memory" : "operand tiles streamed from L2 (32 KB per product, ring of 4 stages)", sms, reps, ms,
               cmean / instr, mac_per_instr * instr / cmean, tops);
        if (mode == 2) printf(", %.2f TB/s of operand stream", (double)reps * 2 * TILE * sms / (ms * 1e-3) * 1e-12);
        printf("\n");
        // what it would mean for the FP64 trailing update: 36 int8 products per FP64 product of the same shape
        const double us_per_128cubed = ms * 1e3 / reps;  // per SM, one 128^3 int8 product
        printf("    -> one FP64-equivalent 128^3 product (36 slice products) = %.2f us of tensor time per SM; the DMMA path needs %.2f us (64 FMA/clk/SM)\n",
               36 * us_per_128cubed, 128.0 * 128 * 128 / 64.0 / (khz * 1e-3));
    }
    return 0;
}
