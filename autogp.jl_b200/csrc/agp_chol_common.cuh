// Shared pieces of the persistent Cholesky kernel (agp_chol_kernel.cu: the persistent loop with the DIAG / PANEL items
// inlined) and the POTF2 item (agp_chol_potf2.cu), which is compiled separately and called through the plain ABI.
//
// Why two translation units: the contraction keeps 64 accumulator + 32 fragment registers per thread in flight and
// fits the 128-register budget of two CTAs per SM with nothing to spare.  As long as both item functions were
// out-of-line functions of ONE translation unit, ptxas' interprocedural register allocation took registers away from
// the contraction whenever the other function changed (round 1: "every variant disturbs do_update's register
// allocation"; round 2: a rewritten POTF2 cost the main loop 100 spill instructions per chunk, 2.4x slower, whichever
// way the functions were split or inlined).  With POTF2 behind a true ABI call (relocatable device code, linked by
// nvlink) the kernel is allocated without knowing it: callee-saved registers are pushed once per POTF2 item (5 % of
// the items), and the contraction's allocation no longer depends on it.  Measured alternatives: every item function
// separately compiled: robust too, but 5 % slower (124 KB of register pushes and pops per item and CTA).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "agp_kernels.cuh"
#include "agp_ptx.cuh"

// Switches that reproduce the round-1 stage race (profiles/r02_race_experiments.txt); all 0 = product.  (The panel main
// loop is now the "simple" form of that record — no per-warp `if (active)` blocks, the diagonal tiles have their own item —
// so AGP_X_NO_RELEASE_FENCE alone brings the failure back.)
#ifndef AGP_X_POTF2_CLK
#define AGP_X_POTF2_CLK 0         // diagnostics: clock totals of the two phases of the POTF2 micro-panels into trace slot 4
#endif
#ifndef AGP_X_NO_RELEASE_FENCE
#define AGP_X_NO_RELEASE_FENCE 0  // drop the cross-proxy fence between a warp's reads of a stage and the stage's release (THE BUG: 1 bad run in 4)
#endif

namespace agp {

constexpr int FT = 256;   // threads per CTA: 8 warps
constexpr int UM = 64;    // item rows
constexpr int UN = TB;    // item columns (one block column)
constexpr int KC = 16;    // K-chunk per pipeline stage (doubles) = one 128-byte row
constexpr int NSTAGE = 4;
constexpr int STAGE_D = (UM + UN) * KC;  // doubles per stage
constexpr int REGION_D = 13056;  // doubles in the region: the ring of four operand stages, the X rows + two W stages, or the packed diagonal tile
constexpr int PROG_SMEM = 64;
// tail of the shared-memory image (doubles): zs[TB] ys[TB] Ri[TB] red[16]
constexpr int TAIL_D = 3 * TB + 16;
constexpr int FUSED_SMEM = (REGION_D + TAIL_D) * 8 + 64 + 64;  // + ctl[16] ints + full[4], empty[4] mbarriers

static_assert(NSTAGE * STAGE_D <= REGION_D, "pipeline stages must fit in the region");
static_assert(2 * (FUSED_SMEM + 1024) <= 228 * 1024, "two CTAs per SM");

// double offset of 16-byte chunk `chunk` (0..7) of row `row` in a [rows][128 B] tile written by TMA with SWIZZLE_128B
__device__ __forceinline__ int swz128(int row, int chunk) { return row * KC + ((chunk ^ (row & 7)) << 1); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}

// thread 0 only: wait until *flag >= need.  A wait that exceeds the limit (2 s by default) raises the scheduler error
// flag so every CTA drains instead of hanging the device.
__device__ __forceinline__ bool wait_ge(const int* flag, int need, int* err, unsigned long long limit_ns) {
    if (ld_acquire_gpu(flag) >= need) return true;
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    for (;;) {
        __nanosleep(64);
        if (ld_acquire_gpu(flag) >= need) return true;
        if ((++spins & 255u) == 0) {
            if (ld_relaxed_gpu(err) != 0) return false;
            if (globaltimer_ns() - t0 > limit_ns) {
                atomicExch(err, 1);
                return false;
            }
        }
    }
}

// mbarrier wait that cannot hang the device: gives up (and raises the scheduler error flag) after the same limit
// as the dependency waits
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int* err, unsigned long long limit_ns) {
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
        if ((spins & 1023u) == 1023u) {
            if (ld_relaxed_gpu(err) != 0) return false;
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > limit_ns) {
                atomicExch(err, 1);
                return false;
            }
        }
    }
}

// diagnostics: thread 0 stamps phase boundaries of item `idx` when tracing is on
__device__ __forceinline__ void stamp(const SchedView& q, int idx, int slot) {
    if (q.trace != nullptr && threadIdx.x == 0) q.trace[(long long)idx * 8 + slot] = (long long)globaltimer_ns();
}

// all threads: release this item's global writes, then bump the counter
__device__ __forceinline__ void signal_done(int* counter) {
    fence_proxy_async();  // this item's shared-memory traffic is ordered before the next item's TMA copies into the same buffers
    fence_proxy_async_global();  // this thread's generic-proxy stores to L are ordered before the TMA (async-proxy) reads of the CTAs the counter releases
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1);
    }
}

// Shared-memory image of one CTA.  Every device function rebuilds this view from the extern
// array itself (never through a pointer argument) so the compiler keeps the shared address
// space and emits LDS/STS instead of generic loads.
struct Smem {
    double* region;
    double* zs;
    double* ys;
    double* Ri;
    double* red;
    int* ctl;  // [0] item index, [1] wait result, [2] potf2 info, [4] pipeline chunks issued so far by this CTA (mbarrier phases)
    uint64_t* full;   // [NSTAGE] stage filled (TMA transaction bytes)
    uint64_t* empty;  // [NSTAGE] stage read by all 8 warps
};

__device__ __forceinline__ Smem smem_view() {
    extern __shared__ __align__(1024) unsigned char smem_raw[];  // TMA destinations with SWIZZLE_128B need 1 KB alignment
    Smem s;
    s.region = reinterpret_cast<double*>(smem_raw);
    s.zs = s.region + REGION_D;
    s.ys = s.zs + TB;
    s.Ri = s.ys + TB;
    s.red = s.Ri + TB;
    s.ctl = reinterpret_cast<int*>(s.red + 16);
    s.full = reinterpret_cast<uint64_t*>(s.ctl + 16);
    s.empty = s.full + NSTAGE;
    return s;
}

// The item's fields, decoded from the queue entry (every phase function decodes them again instead of receiving them:
// whatever the caller keeps in registers across a call is taken away from the callee's 128, see above)
struct ItemFields {
    int type, p, k, i, h, j0, j1, need_k, need_i, extra_flag, extra_need;
    bool diag, partial, yinit;
};
__device__ __forceinline__ ItemFields decode_item(const SchedView& q, int idx) {
    const int4 it = __ldg(q.items + 2 * idx), dep = __ldg(q.items + 2 * idx + 1);
    ItemFields f;
    f.type = it.x & 0xff;
    f.h = (it.x >> 8) & 1;
    f.p = it.y;
    f.k = it.z;
    f.i = it.w;
    f.diag = f.type == ITEM_DIAG;
    f.partial = (it.x & ITEM_PARTIAL) != 0;
    f.yinit = (it.x & ITEM_YINIT) != 0;
    f.j0 = dep.x & 0xffff;
    f.j1 = dep.x >> 16;
    f.need_k = dep.y & 0xffff;
    f.need_i = dep.y >> 16;
    f.extra_flag = dep.z;
    f.extra_need = dep.w;
    return f;
}

// ITEM_POTF2: Cholesky of the diagonal tile, inverses of its diagonal 32x32 blocks (agp_chol_potf2.cu; the DIAG / PANEL
// items are inlined into the kernel's own translation unit)
__device__ bool do_potf2(const BatchView& v, const SchedView& q, int idx);
// ITEM_DIAG: the lower triangle of a diagonal tile minus its contraction, dealt out as 16x16 blocks (agp_chol_diag.cu)
__device__ bool do_diag(const BatchView& v, const SchedView& q, const TmaMaps& maps, int idx);
// ITEM_GRAM: one Gram work unit, K(ts_i, ts_k) [+ noise I] into 64 rows of tile (i,k) (agp_chol_gram.cu)
__device__ bool do_gram(const BatchView& v, const SchedView& q, int idx);
// ITEM_SLICE: int8 digit planes of a finished tile half for the hybrid schedule (agp_chol_slice.cu)
__device__ bool do_slice(const BatchView& v, const SchedView& q, int idx);

}  // namespace agp
