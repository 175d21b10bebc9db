// ITEM_GRAM of agp_chol_kernel: one Gram work unit (agp_gram_unit.cuh) as a work-queue item.  Compiled as its own
// translation unit and called through the plain ABI, like POTF2 (agp_chol_common.cuh says why).
//
// Why the Gram fill is a queue item: as a launch of its own it runs serially in front of the factorisation (1.0 ms of a
// 7.9 ms step at n = 2048 x 64) although it is bound by instruction issue and uses less than half of the FP64 pipe,
// while the factorisation leaves a quarter of that pipe idle (solve / POTF2 phases, dependency waits).  As items the
// units are popped by whatever CTA is free a few hundred items ahead of the first item that reads their tile, and run
// next to the co-resident CTA's DMMA main loop.  Every DIAG / PANEL item that is the first to touch a tile half waits for
// that half's flag (queue builder: fuse_gram_items in agp_api.cu).
#include "agp_chol_common.cuh"
#include "agp_gram_unit.cuh"

namespace agp {

static_assert(GU_FT == FT && GU_M == UM && GU_N == UN, "a Gram unit is one item's tile half");
static_assert((UM + UN) * 8 + PROG_SMEM * (int)sizeof(AgpInstr) <= REGION_D * 8, "time points + program cache fit in the region");

__device__ bool do_gram(const BatchView& v, const SchedView& q, int idx) {
    const Smem s = smem_view();
    const int4 it = __ldg(q.items + 2 * idx);
    const int flag = __ldg(&q.items[2 * idx + 1].z);
    const int p = it.y, k = it.z, i = it.w, h = (it.x >> 8) & 1;
    const int tid = threadIdx.x;
    double* ts_r = s.region;
    double* ts_c = ts_r + UM;
    AgpInstr* prog_s = reinterpret_cast<AgpInstr*>(ts_c + UN);
    stamp(q, idx, 1);
    if (tid < UM) ts_r[tid] = __ldg(v.ts + i * TB + h * UM + tid);
    else if (tid < UM + UN) ts_c[tid - UM] = __ldg(v.ts + k * TB + tid - UM);
    {
        // the host only builds GRAM items when every program of the batch fits the cache (max_prog_len <= PROG_SMEM)
        const int poff = __ldg(v.prog_off + p);
        const int pm = __ldg(v.prog_off + p + 1) - poff;
        const double* src = reinterpret_cast<const double*>(v.prog + poff);
        double* dst = reinterpret_cast<double*>(prog_s);
        for (int w = tid; w < pm * AGP_INSTR_DOUBLES; w += FT) dst[w] = __ldg(src + w);
    }
    __syncthreads();
    gram_unit<4, false>(v, p, i, k, h, ts_r, ts_c, prog_s, tid);
    signal_done(q.head + flag);
    return true;
}

}  // namespace agp
