// Phase A of the DIAG / PANEL work items of agp_chol_kernel: dependency waits, FP64 tensor-core contraction fed by
// 2-D TMA tensor copies, store.  Compiled as its own translation unit (agp_chol_common.cuh says why).
#include "agp_chol_common.cuh"

namespace agp {

// ------------------------------------------------------------------------------------------
// ITEM_DIAG / ITEM_PANEL
// ------------------------------------------------------------------------------------------
// Contraction range [j0, j1) in block columns.  A finishing item (j1 == k) hands the diagonal tile to
// potf2 or takes a panel through the triangular product with W = L_kk^{-1}.  A PARTIAL item only stores: the tile in L
// receives K - sum_{j<j1}; either a later item with j0 = j1 picks it up from there (the accumulators
// always start from minus the tile), which takes the long early part of the contraction of the
// next diagonal tile and of the panel below it off the per-particle critical path.
//
// The item is three functions.  contract() and solve_store() are register-critical leaves (64 accumulator registers +
// 32 fragment registers per thread): they are compiled out of line with nothing but a handful of scalars live around
// them, so that their register allocation does not depend on what else the kernel contains (round 1 and the first
// builds of round 2 lost up to 2x in the main loop to spills whenever an unrelated part of the kernel changed).
// acc = -(tile) ... - sum over the chunks; then X = -acc to shared memory (final panel) or to L (DIAG / PARTIAL items).
__device__ __forceinline__ bool contract(const TmaMaps& maps, const SchedView& q, double* __restrict__ ctile, int ld, int nchunk, int ccol, int brow, int arow, int mode) {
    const Smem s = smem_view();
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const bool diag = (mode & MODE_DIAG) != 0;
    const int h = (mode & MODE_H) ? 1 : 0;
    const int wm = warp >> 2, wn = warp & 3;  // 2 (m) x 4 (n) warps, warp tile 32x32
    const int g = lane >> 2, c4 = lane & 3;
    // warp tiles strictly above the diagonal of a diagonal tile are never read
    const bool active = !diag || (wn * 32 <= h * UM + wm * 32 + 31);
    double* stages = s.region;
    // Operand pipeline: one thread issues 2-D TMA tensor copies (B: 128 rows of tile row k, A: the item's 64 rows;
    // 16 columns = 128 bytes per row, hardware 128-byte swizzle) into a ring of NSTAGE stages; full[] carries the
    // transaction bytes, empty[] the eight warps' "fragments are in registers".  No CTA-wide barrier and no copy
    // instructions in the MMA warps' LSU queue.  The barriers live for the whole kernel: G0 counts the chunks
    // this CTA has issued so far, which gives every stage use its phase parity.
    const int G0 = s.ctl[4];
    auto produce = [&](int c) {  // thread 0 only
        const int G = G0 + c, st = G % NSTAGE;
        if (G >= NSTAGE && !mbar_wait_bounded(s.empty + st, ((G / NSTAGE) - 1) & 1, q.err, q.wait_timeout_ns)) return;
        double* Bs = stages + st * STAGE_D;
        mbar_expect_tx(s.full + st, (diag ? UN : UN + UM) * KC * 8);
        tma_load_2d(Bs, &maps.b, ccol + c * KC, brow, s.full + st);
        if (!diag) tma_load_2d(Bs + UN * KC, &maps.a, ccol + c * KC, arow, s.full + st);
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
            const double2 kv = __ldcg(reinterpret_cast<const double2*>(ctile + (long long)r * ld + c));
            acc[mb][nb][0] = -kv.x;
            acc[mb][nb][1] = -kv.y;
        }

    for (int ch = 0; ch < nchunk; ++ch) {
        const int G = G0 + ch, st = G % NSTAGE;
        if (tid == 0 && ch + NSTAGE - 1 < nchunk) produce(ch + NSTAGE - 1);
        if (!mbar_wait_bounded(s.full + st, (G / NSTAGE) & 1, q.err, q.wait_timeout_ns)) return false;
        const double* Bs = stages + st * STAGE_D;
        const double* As = diag ? Bs + h * UM * KC : Bs + UN * KC;  // diagonal tile: A rows are a slice of B
        // lane c4 takes the 16-byte chunks 2 c4 + ks of a row (a permutation of k shared by A and B): with the
        // 128-byte swizzle the eight lanes of an LDS.128 phase then hit eight different chunk columns
#if AGP_X_SIMPLE
#define AGP_ACTIVE_IF
#else
#define AGP_ACTIVE_IF if (active)
#endif
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            double2 a[4], b[4];
            AGP_ACTIVE_IF {
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) a[mb] = *reinterpret_cast<const double2*>(As + swz128(wm * 32 + mb * 8 + g, 2 * c4 + ks));
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) b[nb] = *reinterpret_cast<const double2*>(Bs + swz128(wn * 32 + nb * 8 + g, 2 * c4 + ks));
            }
            AGP_ACTIVE_IF {
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].x, b[nb].x);
            }
            if (ks == 1) {
                // Release of the stage.  The LDS above are generic-proxy reads, the next use of the stage is written by the
                // async proxy (TMA): every lane orders its own reads before the release with a cross-proxy fence, exactly as
                // CUTLASS does before consumer_release when a TMA-fed buffer is read with ordinary loads.  Without it the
                // TMA box of chunk ch + NSTAGE can land while a late LDS of chunk ch is still queued behind the co-resident
                // CTA's shared-memory traffic (round 1's "one wrong row / 8x8 block in 1 of 10^4 items").  The fence waits
                // for this warp's outstanding shared-memory loads: it sits behind the first half of the k-step's DMMAs,
                // which cannot issue before those loads have returned, so it never stalls.
#if !AGP_X_NO_RELEASE_FENCE
                fence_proxy_async();
#endif
                __syncwarp();
                if (lane == 0) mbar_arrive(s.empty + st);
            }
            AGP_ACTIVE_IF {
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].y, b[nb].y);
            }
        }
    }
    if (tid == 0) s.ctl[4] = G0 + nchunk;
    __syncthreads();  // every warp is done with every stage: the region may be rewritten

    // --- X = -acc:  a DIAG / PARTIAL tile goes back to L (lower part of a diagonal tile), a final panel to shared memory ----
    if (active) {
        double* Xs = s.region + XS_OFF;  // [UM][XS2]
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const int r = wm * 32 + mb * 8 + g, c = wn * 32 + nb * 8 + 2 * c4;
                const double2 x2 = make_double2(-acc[mb][nb][0], -acc[mb][nb][1]);
                if (!(mode & MODE_GLOBAL)) {
                    *reinterpret_cast<double2*>(Xs + r * XS2 + c) = x2;
                } else {
                    const int rd = diag ? h * UM + r : TB;  // row inside a diagonal tile: only c <= rd is kept
                    double* dst = ctile + (long long)r * ld + c;
                    if (c + 1 <= rd) *reinterpret_cast<double2*>(dst) = x2;
                    else if (c <= rd) dst[0] = x2.x;
                }
            }
    }
    return true;
}

// phase A of a DIAG / PANEL item: dependency waits, contraction, store.  Returns 0: failed (drain), 1: item complete,
// 2: a final panel, X is in shared memory and phase B follows
__device__ int update_contract(const BatchView& v, const SchedView& q, const TmaMaps& maps, int idx) {
    const Smem s = smem_view();
    const ItemFields f = decode_item(q, idx);
    const int tid = threadIdx.x;
    const int p = f.p, k = f.k, i = f.i;
    const int row0 = i * TB + f.h * UM, col0 = k * TB;
    const int ld = v.ld;

    // the operand tile rows k and i must be final over the contraction range (counter values chosen by
    // the queue builder); a continuation item also needs the partial tile of its predecessor
    if (f.need_k > 0 || f.need_i > 0 || f.extra_flag >= 0) {
        if (tid == 0) {
            bool ok = true;
            if (f.need_k > 0) ok = wait_ge(q.rowdone + p * q.nt_stride + k, f.need_k, q.err, q.wait_timeout_ns);
            if (ok && f.need_i > 0) ok = wait_ge(q.rowdone + p * q.nt_stride + i, f.need_i, q.err, q.wait_timeout_ns);
            if (ok && f.extra_flag >= 0) ok = wait_ge(q.head + f.extra_flag, f.extra_need, q.err, q.wait_timeout_ns);
            s.ctl[1] = ok ? 1 : 0;
        }
        __syncthreads();
        if (!s.ctl[1]) return 0;
    }
    stamp(q, idx, 1);

    // The Gram tile K(ts_i, ts_k) [+ noise I] was written into L by agp_gramfill_kernel.  Block
    // column 0 needs no contraction: the diagonal tile is already in place and a panel goes
    // straight to shared memory.  Otherwise the accumulators start from -K, so that after the
    // contraction  acc = -(K - sum_j L_ij L_kj^T).
    if (k == 0 && f.diag) {
        if (tid < UM) v.y[(long long)p * ld + row0 + tid] = (row0 + tid < v.n) ? v.xs[row0 + tid] : 0.0;
        signal_done(q.diagu + p * q.nt_stride + k);
        return 1;
    }
    double* ctile = v.L + (long long)p * v.mat_stride + (long long)row0 * ld + col0;
    if (!contract(maps, q, ctile, ld, ((f.j1 - f.j0) * TB) / KC, f.j0 * TB, p * ld + col0, p * ld + row0,
                  (f.diag ? MODE_DIAG : 0) | (f.h ? MODE_H : 0) | ((f.diag || f.partial) ? MODE_GLOBAL : 0)))
        return 0;
    stamp(q, idx, 2);
    if (f.diag) {
        signal_done(q.diagu + p * q.nt_stride + k);
        return 1;
    }
    if (f.partial) {
        signal_done(q.ppre + p * q.nt_stride + i);
        return 1;
    }
    return 2;
}

}  // namespace agp
