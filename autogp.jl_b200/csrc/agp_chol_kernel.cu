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
// The Gram matrix K(ts,ts) + noise*I is evaluated tile half by tile half by the kernel-tree interpreter
// (agp_gram_unit.cuh) and left in L: either by ITEM_GRAM items of this kernel's own queue, popped a few
// hundred items ahead of the first item that reads the tile (plain LML runs), or by agp_gramfill_kernel in
// front of this launch (continuations with appended rows).  Every tile's accumulators then start from -K,
// so the Gram work never sits between two DMMA main loops of one item.
//
//   ITEM_GRAM  (p,k,i,h)  64 rows of tile (i,k)  <-  K(ts_i, ts_k) [+ noise I]
//   ITEM_DIAG  (p,k,h)    half of the 36 lower 16x16 blocks of the diagonal tile:  K(ts_k,ts_k) + noise I - sum_j L_kj L_kj^T
//                         (agp_chol_diag.cu)
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

namespace {

constexpr int XS = 136;  // X / L_kk-panel row stride: 8 mod 16 doubles -> conflict-free LDS.128
static_assert(UM * XS + 4 * 32 * 32 <= REGION_D, "X rows + the solve's ring of four 32x32 operand blocks");

// ------------------------------------------------------------------------------------------
// ITEM_PANEL
// ------------------------------------------------------------------------------------------
// Contraction range [j0, j1) in block columns.  A finishing item (j1 == k) takes the panel through the
// triangular solve.  A PARTIAL item only stores: the tile in L receives K - sum_{j<j1}; a later item with
// j0 = j1 picks it up from there (the accumulators always start from minus the tile), which takes the long
// early part of the contraction of the panel below the next diagonal tile off the per-particle critical path.
__device__ __forceinline__ bool do_update(const BatchView& v, const SchedView& q, const TmaMaps& maps, int idx) {
    const Smem s = smem_view();
    const ItemFields f = decode_item(q, idx);
    const int p = f.p, k = f.k, i = f.i, h = f.h, j0 = f.j0, j1 = f.j1, need_k = f.need_k, need_i = f.need_i, extra_flag = f.extra_flag, extra_need = f.extra_need;
    const bool partial = f.partial, yinit = f.yinit;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int row0 = i * TB + h * UM, col0 = k * TB;
    const int ld = v.ld;
    double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    double* stages = s.region;

    // the operand tile rows k and i must be final over the contraction range (counter values chosen by
    // the queue builder); a continuation item also needs the partial tile of its predecessor
    if (need_k > 0 || need_i > 0 || extra_flag >= 0) {
        if (tid == 0) {
            bool ok = true;
            if (need_k > 0) ok = wait_ge(q.rowdone + p * q.nt_stride + k, need_k, q.err, q.wait_timeout_ns);
            if (ok && need_i > 0) ok = wait_ge(q.rowdone + p * q.nt_stride + i, need_i, q.err, q.wait_timeout_ns);
            if (ok && extra_flag >= 0) ok = wait_ge(q.head + extra_flag, extra_need, q.err, q.wait_timeout_ns);
            s.ctl[1] = ok ? 1 : 0;
        }
        __syncthreads();
        if (!s.ctl[1]) return false;
    }
    stamp(q, idx, 1);

    // --- contraction: acc = sum_{j<k} L_ij L_kj^T -----------------------------------------
    const int wm = warp >> 2, wn = warp & 3;  // 2 (m) x 4 (n) warps, warp tile 32x32
    const int g = lane >> 2, c4 = lane & 3;
    // The Gram tile K(ts_i, ts_k) was written into L by a GRAM item or by agp_gramfill_kernel.  Block column 0 needs no
    // contraction: the panel goes straight to shared memory.  Otherwise the accumulators start from -K, so that after
    // the contraction  acc = -(K - sum_j L_ij L_kj^T).
    const int nchunk = ((j1 - j0) * TB) / KC;
    // Operand pipeline: one thread issues 2-D TMA tensor copies (B: 128 rows of tile row k, A: the item's 64 rows;
    // 16 columns = 128 bytes per row, hardware 128-byte swizzle) into a ring of NSTAGE stages; full[] carries the
    // transaction bytes, empty[] the eight warps' "fragments are in registers".  No CTA-wide barrier and no copy
    // instructions in the MMA warps' LSU queue.  The barriers live for the whole kernel: G0 counts the chunks
    // this CTA has issued so far, which gives every stage use its phase parity.
    const int G0 = s.ctl[4];
    const int ccol = j0 * TB, brow = p * ld + col0, arow = p * ld + row0;
    auto produce = [&](int c) {  // thread 0 only
        const int G = G0 + c, st = G % NSTAGE;
        if (G >= NSTAGE && !mbar_wait_bounded(s.empty + st, ((G / NSTAGE) - 1) & 1, q.err, q.wait_timeout_ns)) return;
        double* Bs = stages + st * STAGE_D;
        mbar_expect_tx(s.full + st, (UN + UM) * KC * 8);
        tma_load_2d(Bs, &maps.b, ccol + c * KC, brow, s.full + st);
        tma_load_2d(Bs + UN * KC, &maps.a, ccol + c * KC, arow, s.full + st);
    };
    if (tid == 0) {
        fence_proxy_async_all();  // after the acquire of the dependency counters, before this item's first async-proxy reads of L
        for (int c = 0; c < NSTAGE - 1 && c < nchunk; ++c) produce(c);
    }
    // (after the first copies are in flight, so the two L2 round trips overlap)
    double acc[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const int r = wm * 32 + mb * 8 + g, c = wn * 32 + nb * 8 + 2 * c4;
            const double2 kv = __ldcg(reinterpret_cast<const double2*>(Lp + (long long)(row0 + r) * ld + col0 + c));
            acc[mb][nb][0] = -kv.x;
            acc[mb][nb][1] = -kv.y;
        }

    for (int ch = 0; ch < nchunk; ++ch) {
        const int G = G0 + ch, st = G % NSTAGE;
        if (tid == 0 && ch + NSTAGE - 1 < nchunk) produce(ch + NSTAGE - 1);
        if (!mbar_wait_bounded(s.full + st, (G / NSTAGE) & 1, q.err, q.wait_timeout_ns)) return false;
        const double* Bs = stages + st * STAGE_D;
        const double* As = Bs + UN * KC;
        // lane c4 takes the 16-byte chunks 2 c4 + ks of a row (a permutation of k shared by A and B): with the
        // 128-byte swizzle the eight lanes of an LDS.128 phase then hit eight different chunk columns
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            double2 a[4], b[4];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) a[mb] = *reinterpret_cast<const double2*>(As + swz128(wm * 32 + mb * 8 + g, 2 * c4 + ks));
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) b[nb] = *reinterpret_cast<const double2*>(Bs + swz128(wn * 32 + nb * 8 + g, 2 * c4 + ks));
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].x, b[nb].x);
            if (ks == 1) {
                // Release of the stage.  The LDS above are generic-proxy reads, the next use of the stage is written by the
                // async proxy (TMA): every lane orders its own reads before the release with a cross-proxy fence, exactly as
                // CUTLASS does before consumer_release when a TMA-fed buffer is read with ordinary loads.  Without it the
                // TMA box of chunk ch + NSTAGE can land while a late LDS of chunk ch is still queued behind the co-resident
                // CTA's shared-memory traffic (round 1's "one wrong row / 8x8 block in 1 of 10^4 items").  The fence waits for this
                // warp's outstanding shared-memory loads; behind the first half of the k-step's DMMAs (which cannot issue
                // before those loads have returned) it never stalls.
#if !AGP_X_NO_RELEASE_FENCE
                fence_proxy_async();
#endif
                __syncwarp();
                if (lane == 0) mbar_arrive(s.empty + st);
            }
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].y, b[nb].y);
        }
    }
    __syncthreads();  // every warp is done with the operand stages: X overwrites them below
    // (behind the barrier rather than in front of it: the item's first read of ctl[4] is ordered before this write by the
    // stage barriers alone — nchunk is 0 or >= 8 > NSTAGE — which compute-sanitizer's racecheck does not model)
    if (tid == 0) s.ctl[4] = G0 + nchunk;
    stamp(q, idx, 2);

    // --- X = -acc:  a store-only item writes it back to L, a finishing panel keeps it in shared memory ----
    double* Xs = s.region;            // [UM][XS]
    double* Ls = s.region + UM * XS;  // [32][XS]
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const int r = wm * 32 + mb * 8 + g, c = wn * 32 + nb * 8 + 2 * c4;
            const double2 x2 = make_double2(-acc[mb][nb][0], -acc[mb][nb][1]);
            if (!partial) *reinterpret_cast<double2*>(Xs + r * XS + c) = x2;
            else *reinterpret_cast<double2*>(Lp + (long long)(row0 + r) * ld + col0 + c) = x2;
        }
    if (partial) {
        signal_done(q.ppre + p * q.nt_stride + i);
        return true;
    }
    double* yp = v.y + (long long)p * ld;
    const int n = v.n;

    // --- triangular solve against L_kk in shared memory ------------------------------------
    stamp(q, idx, 3);
    if (tid == 0) s.ctl[1] = wait_ge(q.fdone + p, k + 1, q.err, q.wait_timeout_ns) ? 1 : 0;
    __syncthreads();  // also publishes X
    if (!s.ctl[1]) return false;
    stamp(q, idx, 4);

    const int o = col0;
    const double* dinv = v.dinv + ((long long)p * q.nt_stride + k) * 4096;  // per block column: POTF2(k+1) may run while column k is still being solved
    if (tid < TB) s.zs[tid] = __ldcg(v.z + (long long)p * ld + o + tid);
    double* xrow = Xs + (warp * 8 + g) * XS;  // this lane's row (fragment row g of the warp's 8 rows)
    double y_old = 0.0;
    if (lane < 8) {
        const int gr = row0 + warp * 8 + lane;
        y_old = yinit ? ((gr < n) ? v.xs[gr] : 0.0) : __ldcg(yp + gr);
    }

    // Blocked substitution over the four 32-column blocks of L_kk:  X_jb = (C_jb - sum_{m<jb} X_m L[jb,m]^T) inv(L[jb,jb])^T.
    // The ten 32x32 operand blocks (six of L_kk, four inverted diagonal blocks) stream through a ring of four
    // shared-memory slots, three blocks ahead of the math, so their L2 latency is paid once, not per block.
    // Each warp owns 8 rows of X, so the block-to-block dependency is warp-local.
    constexpr int SB = 32 * 32;  // doubles per slot, 16-byte chunks XOR-swizzled (conflict-free LDS.128)
    auto blk_swz = [](int row, int chunk) { return row * 32 + ((chunk ^ ((row & 1) << 2)) << 1); };
    auto load_block = [&](int b) {
        // b -> (jb, m): 0:(0,0) 1:(1,0) 2:(1,1) 3:(2,0) 4:(2,1) 5:(2,2) 6:(3,0) 7:(3,1) 8:(3,2) 9:(3,3); m == jb: inverse block
        const int jb = (b >= 6) ? 3 : (b >= 3) ? 2 : (b >= 1) ? 1 : 0;
        const int mb = b - jb * (jb + 1) / 2;
        double* dst = Ls + (b & 3) * SB;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int w = tid + e * FT;
            const int r = w >> 4, ch = w & 15;
            const double* src = (mb < jb) ? Lp + (long long)(o + jb * 32 + r) * ld + o + mb * 32 + ch * 2 : dinv + jb * 1024 + r * 32 + ch * 2;
            cp_async16(dst + blk_swz(r, ch), src);
        }
    };
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        load_block(b);
        cp_async_commit();
    }
    // two accumulator sets (even / odd k of each LDS.128 pair): 8 independent DMMA chains
    double acc0[4][2], acc1[4][2];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) acc0[nb][0] = acc0[nb][1] = acc1[nb][0] = acc1[nb][1] = 0.0;
    int jb = 0, mb = 0;
#pragma unroll 1
    for (int b = 0; b < 10; ++b) {
        cp_async_wait<2>();
        __syncthreads();  // block b has landed for everyone, and everyone is done with the slot of block b-1
        if (b + 3 < 10) load_block(b + 3);
        cp_async_commit();
        const double* Bs = Ls + (b & 3) * SB;
        if (mb == jb) {
            // T = C_jb - S  (own rows only), then X_jb = T inv(L_jb,jb)^T
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                double2* ptr = reinterpret_cast<double2*>(xrow + jb * 32 + nb * 8 + 2 * c4);
                double2 t = *ptr;
                t.x -= acc0[nb][0] + acc1[nb][0];
                t.y -= acc0[nb][1] + acc1[nb][1];
                *ptr = t;
                acc0[nb][0] = acc0[nb][1] = acc1[nb][0] = acc1[nb][1] = 0.0;
            }
            __syncwarp();
        }
        const double* xa = xrow + mb * 32;
        if (mb == jb) {
            // the inverted diagonal block is lower triangular (exact zeros above): output columns of unit nb take nothing from
            // the k-blocks behind it — 10 of the 16 unit products, 15 % of the solve's DMMAs
#pragma unroll
            for (int kk = 0; kk < 32; kk += 8) {
                const double2 a = *reinterpret_cast<const double2*>(xa + kk + 2 * c4);
#pragma unroll
                for (int nb = kk >> 3; nb < 4; ++nb) {
                    const double2 bb = *reinterpret_cast<const double2*>(Bs + blk_swz(nb * 8 + g, (kk >> 1) + c4));
                    dmma884(acc0[nb][0], acc0[nb][1], a.x, bb.x);
                    dmma884(acc1[nb][0], acc1[nb][1], a.y, bb.y);
                }
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < 32; kk += 8) {
                const double2 a = *reinterpret_cast<const double2*>(xa + kk + 2 * c4);
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const double2 bb = *reinterpret_cast<const double2*>(Bs + blk_swz(nb * 8 + g, (kk >> 1) + c4));
                    dmma884(acc0[nb][0], acc0[nb][1], a.x, bb.x);
                    dmma884(acc1[nb][0], acc1[nb][1], a.y, bb.y);
                }
            }
        }
        if (mb == jb) {
            __syncwarp();
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                *reinterpret_cast<double2*>(xrow + jb * 32 + nb * 8 + 2 * c4) =
                    make_double2(acc0[nb][0] + acc1[nb][0], acc0[nb][1] + acc1[nb][1]);
                acc0[nb][0] = acc0[nb][1] = acc1[nb][0] = acc1[nb][1] = 0.0;
            }
            __syncwarp();
            ++jb;
            mb = 0;
        } else {
            ++mb;
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // store L_ik rows (coalesced) and fold the forward solve: y_i -= L_ik z_k
    double dot_mine = 0.0;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const int r = warp * 8 + rr;
        const double* xr = Xs + r * XS;
        double sacc = 0.0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int c = lane + e * 32;
            double x = xr[c];
            Lp[(long long)(row0 + r) * ld + o + c] = x;
            sacc = fma(x, s.zs[c], sacc);
        }
        sacc = warp_sum(sacc);
        if (lane == rr) dot_mine = sacc;
    }
    if (lane < 8) yp[row0 + warp * 8 + lane] = y_old - dot_mine;
    signal_done(q.rowdone + p * q.nt_stride + i);
    return true;
}


}  // namespace

// AGP_CHOL_MIN_CTAS: 2 in the shipped instantiation (two CTAs per SM, 128 registers); the Makefile compiles the kernel's
// translation units a second time with -DAGP_CHOL_MIN_CTAS=1, no register cap and every external name suffixed _solo — the
// instantiation small batches run on, one CTA per SM (chol_ctas in agp_api.cu; csrc/experiments/README.md: -3 .. -5 %)
#ifndef AGP_CHOL_MIN_CTAS
#define AGP_CHOL_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(FT, AGP_CHOL_MIN_CTAS) agp_chol_kernel(const __grid_constant__ BatchView v, const __grid_constant__ SchedView q, const __grid_constant__ TmaMaps maps) {
    const Smem s = smem_view();
    if (threadIdx.x == 0) {
        // the two TMA descriptors are fetched now, not on the first copy of the first item
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&maps.a) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&maps.b) : "memory");
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
        } else if (type == ITEM_GRAM) {
            ok = do_gram(v, q, idx);
        } else if (type == ITEM_SLICE) {
            ok = do_slice(v, q, idx);
        } else if (type == ITEM_DIAG) {
            ok = do_diag(v, q, maps, idx);
        } else {
            ok = do_update(v, q, maps, idx);
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

bool make_tma_maps(double* L, int ld, long long rows, TmaMaps* out) {
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
    return r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS;
}

}  // namespace agp
