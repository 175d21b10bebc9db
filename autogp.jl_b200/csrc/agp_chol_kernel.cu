// Persistent dataflow kernel of the fused GP log-marginal-likelihood path (sm_100a).
//
// One launch factorises every particle's K + noise*I (blocked left-looking Cholesky, block
// column width 128) and finishes the LML.  CTAs (two per SM) pop work items from an in-order
// queue; every item's producers sit EARLIER in the queue, so a consumer that spins on a
// dependency counter always waits for a CTA that is already running: no deadlock, no
// co-residency requirement, no tail waves between the stages of a block column, and the two
// CTAs of an SM drift apart so one CTA's solve / factorisation phases fill the gaps of the other's
// DMMA main loop.  The contraction operands arrive by 2-D TMA tensor copies through a ring of
// four stages guarded by full/empty mbarriers (no CTA barrier in the main loop).
//
// agp_gramfill_kernel runs first: it evaluates every particle's kernel-tree program over the lower
// 128x128 tiles and leaves K(ts,ts) + noise*I in L (ts slices staged by 1-D TMA bulk copies, all
// warps of the SM in the FP64-ALU-bound interpreter at once).  The persistent kernel then starts
// each tile's accumulators from -K, so the Gram work never sits between two DMMA main loops.
//
//   ITEM_DIAG  (p,k,h)    64 rows of the diagonal tile:  K(ts_k,ts_k) + noise I - sum_j L_kj L_kj^T
//   ITEM_POTF2 (p,k)      Cholesky of the 128x128 diagonal tile (+ observation row): L_kk, z_k,
//                         log det, z'z, LAPACK info, inverses of the 32x32 diagonal blocks
//   ITEM_PANEL (p,k,i,h)  64 rows of tile (i,k): K - contraction on FP64 tensor cores, then the
//                         triangular solve against L_kk IN SHARED MEMORY (the unfactored tile never
//                         touches HBM), then y_i -= L_ik z_k (forward solve folded in)
//
// Reference semantics: src/GP.jl:137-503, 666-668 (Gram), src/Model.jl:134-136 (noise, mvnormal),
// Distributions' MvNormal logpdf = -(n log 2pi + logdet)/2 - |U^{-T} x|^2/2 with K = U'U.
#include "agp_chol_common.cuh"

namespace agp {

__global__ void __launch_bounds__(FT, 2) agp_chol_kernel(BatchView v, SchedView q, const __grid_constant__ TmaMaps maps) {
    const Smem s = smem_view();
    if (threadIdx.x == 0) {
        // the two TMA descriptors are fetched now, not on the first copy of the first item
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&maps.a) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&maps.b) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&maps.w) : "memory");
        for (int st = 0; st < NSTAGE; ++st) {
            mbar_init(s.full + st, 1);
            mbar_init(s.empty + st, FT / 32);
        }
        mbar_fence_init();
        s.ctl[4] = 0;
    }
    __syncthreads();

    for (;;) {
        if (threadIdx.x == 0) s.ctl[0] = atomicAdd(q.head, 1);
        __syncthreads();
        const int idx = s.ctl[0];
        if (idx >= q.n_items) break;
        const int type = __ldg(&q.items[2 * idx].x) & 0xff;
        bool ok;
        stamp(q, idx, 0);
        if (type == ITEM_POTF2) {
            ok = do_potf2(v, q, idx);
        } else {
            int r = update_contract(v, q, maps, idx);
            if (r == 2) r = update_solve(v, q, maps, idx);
            ok = r != 0;
        }
        if (!ok) break;
        stamp(q, idx, 5);
        if (q.trace != nullptr && threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            q.trace[(long long)idx * 8 + 6] = (long long)smid;
            q.trace[(long long)idx * 8 + 7] = (long long)blockIdx.x;
        }
    }
}

cudaError_t configure_fused() {
    return cudaFuncSetAttribute(agp_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
}

void launch_chol(const BatchView& v, const SchedView& q, const TmaMaps& maps, int ctas, cudaStream_t s) {
    if (q.n_items <= 0) return;
    if (ctas > q.n_items) ctas = q.n_items;
    agp_chol_kernel<<<ctas, FT, FUSED_SMEM, s>>>(v, q, maps);
}

bool make_tma_maps(double* L, int ld, long long rows, double* W, long long w_rows, TmaMaps* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return false;
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box_a[2] = {(cuuint32_t)KC, (cuuint32_t)UM}, box_b[2] = {(cuuint32_t)KC, (cuuint32_t)UN}, estr[2] = {1, 1};
    const CUresult r1 = encode(&out->a, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, L, dims, strides, box_a, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const CUresult r2 = encode(&out->b, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, L, dims, strides, box_b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    // W = L_kk^{-1} tiles: [w_rows][128], boxes of 16 columns x 128 rows (one k-chunk of the triangular product)
    const cuuint64_t wdims[2] = {(cuuint64_t)TB, (cuuint64_t)w_rows};
    const cuuint64_t wstrides[1] = {(cuuint64_t)TB * 8};
    const CUresult r3 = encode(&out->w, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, W, wdims, wstrides, box_b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS && r3 == CUDA_SUCCESS;
}

}  // namespace agp
