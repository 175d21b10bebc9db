// ITEM_DIAG of agp_chol_kernel: the lower triangle of a diagonal 128x128 tile,
//     K(ts_k, ts_k) + noise I - sum_{j0 <= j < j1} L_kj L_kj^T.
// Compiled as its own translation unit and called through the plain ABI, like POTF2 (agp_chol_common.cuh says why: the
// panel item's main loop keeps its register allocation whatever happens here).
//
// Rounds 1-2a ran a diagonal tile through the panel item's 32x32 warp tiles: two items of 64 rows, 3 and 7 of their 8
// warps busy (the tiles strictly above the diagonal skipped), each item as long as a full panel contraction.  In 8x8
// DMMA units the lower triangle is 136 units of work (six 32x32 blocks below the diagonal = 16 units each, four on it =
// 10 units each: their strictly upper units are never read), where the panel shape spends 16 units of time on each of 16
// warps.  Here every block below the diagonal is split into two 32x16 halves (8 units) and the work is dealt out over
// the same 2 x 8 warps: six warps of an item take one half block each (4 A fragments x 2 B fragments), two take one
// diagonal block each (4 fragments that serve as A and as B: 10 units), so an item lasts 10/16 of a panel contraction of
// the same depth.  The two items of a tile still exist (the queue, its counters and POTF2's wait are unchanged); they
// split the blocks, not the rows.  Both read the whole tile row k (B operand = A operand: 128 rows x 16 columns per
// pipeline chunk, 2-D TMA).  Measured on the way (n = 2048 x 64, per 128-deep product and item): panel-shaped item
// 11.6 us of CTA time (16 of 32 warp tiles idle); 36 blocks of 16x16 dealt out three to a warp: 16.2 us (runtime block
// lists: four LDS, their latency, eight DMMAs, a branch — no overlap between blocks).
#include "agp_chol_common.cuh"

namespace agp {

__device__ bool do_diag(const BatchView& v, const SchedView& q, const TmaMaps& maps, int idx) {
    const Smem s = smem_view();
    const ItemFields f = decode_item(q, idx);
    const int p = f.p, k = f.k, h = f.h, j0 = f.j0, j1 = f.j1;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int col0 = k * TB;
    const int ld = v.ld;
    double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    double* stages = s.region;

    if (f.need_k > 0 || f.need_i > 0 || f.extra_flag >= 0) {
        if (tid == 0) {
            bool ok = true;
            if (f.need_k > 0) ok = wait_ge(q.rowdone + p * q.nt_stride + k, f.need_k, q.err, q.wait_timeout_ns);
            if (ok && f.need_i > 0) ok = wait_ge(q.rowdone + p * q.nt_stride + f.i, f.need_i, q.err, q.wait_timeout_ns);
            if (ok && f.extra_flag >= 0) ok = wait_ge(q.head + f.extra_flag, f.extra_need, q.err, q.wait_timeout_ns);
            s.ctl[1] = ok ? 1 : 0;
        }
        __syncthreads();
        if (!s.ctl[1]) return false;
    }
    stamp(q, idx, 1);

    // Block column 0 needs no contraction: the Gram tile is the diagonal tile.  Its two items start the forward solve:
    // y_0 = xs over the 64 rows of their half.
    if (k == 0) {
        const int row0 = h * UM;
        if (tid < UM) v.y[(long long)p * ld + row0 + tid] = (row0 + tid < v.n) ? v.xs[row0 + tid] : 0.0;
        signal_done(q.diagu + p * q.nt_stride + k);
        return true;
    }

    const int g = lane >> 2, c4 = lane & 3;
    // this warp's share.  Warps 0..5: half hf of the block below the diagonal number 3 h + warp / 2 in the order
    // (1,0) (2,0) (2,1) | (3,0) (3,1) (3,2) (32x32 block row, column); warps 6, 7: diagonal block 2 h + warp - 6.
    const bool tri = warp >= 6;
    const int bidx = 3 * h + (warp >> 1);
    const int R = tri ? 2 * h + warp - 6 : (bidx >= 3 ? 3 : bidx >= 1 ? 2 : 1);
    const int C = tri ? R : (bidx >= 3 ? bidx - 3 : bidx >= 1 ? bidx - 1 : 0);
    const int arow = R * 32 + g;                     // + mb * 8: rows of the A fragments
    const int brow = C * 32 + (warp & 1) * 16 + g;   // + nb * 8: rows of tile row k that are this warp's COLUMNS (half blocks only)

    const int nchunk = ((j1 - j0) * TB) / KC;
    const int G0 = s.ctl[4];
    const int ccol = j0 * TB, trow = p * ld + col0;
    auto produce = [&](int c) {  // thread 0 only
        const int G = G0 + c, st = G % NSTAGE;
        if (G >= NSTAGE && !mbar_wait_bounded(s.empty + st, ((G / NSTAGE) - 1) & 1, q.err, q.wait_timeout_ns)) return;
        double* Bs = stages + st * STAGE_D;
        mbar_expect_tx(s.full + st, UN * KC * 8);
        tma_load_2d(Bs, &maps.b, ccol + c * KC, trow, s.full + st);
    };
    if (tid == 0) {
        fence_proxy_async_all();  // after the acquire of the dependency counters, before this item's first async-proxy reads of L
        for (int c = 0; c < NSTAGE - 1 && c < nchunk; ++c) produce(c);
    }
    // Accumulators start from minus the tile (Gram values, or the partial sums of an earlier item of this tile).
    // Half block: unit (mb, nb) -> acc[2 mb + nb], mb < 4, nb < 2.  Diagonal block: unit (mb, nb <= mb) -> acc[mb (mb + 1) / 2 + nb].
    double acc[10][2];
    if (tri) {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb <= mb; ++nb) {
                const double2 kv = __ldcg(reinterpret_cast<const double2*>(Lp + (long long)(col0 + arow + mb * 8) * ld + col0 + R * 32 + nb * 8 + 2 * c4));
                acc[mb * (mb + 1) / 2 + nb][0] = -kv.x;
                acc[mb * (mb + 1) / 2 + nb][1] = -kv.y;
            }
    } else {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
                const double2 kv = __ldcg(reinterpret_cast<const double2*>(Lp + (long long)(col0 + arow + mb * 8) * ld + col0 + brow - g + nb * 8 + 2 * c4));
                acc[2 * mb + nb][0] = -kv.x;
                acc[2 * mb + nb][1] = -kv.y;
            }
        acc[8][0] = acc[8][1] = acc[9][0] = acc[9][1] = 0.0;
    }

    for (int ch = 0; ch < nchunk; ++ch) {
        const int G = G0 + ch, st = G % NSTAGE;
        if (tid == 0 && ch + NSTAGE - 1 < nchunk) produce(ch + NSTAGE - 1);
        if (!mbar_wait_bounded(s.full + st, (G / NSTAGE) & 1, q.err, q.wait_timeout_ns)) return false;
        const double* Bs = stages + st * STAGE_D;
        // Per k-step: the fragment loads, every unit's even-k DMMA, the release of the stage after the second k-step's loads
        // (every lane orders its generic-proxy reads before the async-proxy write of the next box into the stage:
        // profiles/r02_race_experiments.txt; tools/sass_lint.py checks the built code), then every unit's odd-k DMMA — the
        // two DMMAs of a unit depend on each other and sit 8 to 10 DMMAs apart.
        if (tri) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                double2 a[4];  // rows of the diagonal block: A fragment of row unit mb = B fragment of column unit mb
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) a[mb] = *reinterpret_cast<const double2*>(Bs + swz128(arow + mb * 8, 2 * c4 + ks));
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb <= mb; ++nb) dmma884(acc[mb * (mb + 1) / 2 + nb][0], acc[mb * (mb + 1) / 2 + nb][1], a[mb].x, a[nb].x);
                if (ks == 1) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s.empty + st);
                }
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb <= mb; ++nb) dmma884(acc[mb * (mb + 1) / 2 + nb][0], acc[mb * (mb + 1) / 2 + nb][1], a[mb].y, a[nb].y);
            }
        } else {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                double2 a[4], b[2];
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) a[mb] = *reinterpret_cast<const double2*>(Bs + swz128(arow + mb * 8, 2 * c4 + ks));
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) b[nb] = *reinterpret_cast<const double2*>(Bs + swz128(brow + nb * 8, 2 * c4 + ks));
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb) dmma884(acc[2 * mb + nb][0], acc[2 * mb + nb][1], a[mb].x, b[nb].x);
                if (ks == 1) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s.empty + st);
                }
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb) dmma884(acc[2 * mb + nb][0], acc[2 * mb + nb][1], a[mb].y, b[nb].y);
            }
        }
    }
    __syncthreads();  // orders every thread's read of ctl[4] above before this write for tools that do not model mbarriers
    if (tid == 0) s.ctl[4] = G0 + nchunk;
    stamp(q, idx, 2);

    // the lower part goes back to L, for POTF2 or for the item that continues the contraction
    if (tri) {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb <= mb; ++nb) {
                const int r = mb * 8 + g, c = nb * 8 + 2 * c4;  // inside the diagonal block: only c <= r is kept
                double* dst = Lp + (long long)(col0 + R * 32 + r) * ld + col0 + R * 32 + c;
                const int u = mb * (mb + 1) / 2 + nb;
                if (c + 1 <= r) *reinterpret_cast<double2*>(dst) = make_double2(-acc[u][0], -acc[u][1]);
                else if (c <= r) dst[0] = -acc[u][0];
            }
    } else {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
                *reinterpret_cast<double2*>(Lp + (long long)(col0 + arow + mb * 8) * ld + col0 + brow - g + nb * 8 + 2 * c4) =
                    make_double2(-acc[2 * mb + nb][0], -acc[2 * mb + nb][1]);
    }
    signal_done(q.diagu + p * q.nt_stride + k);
    return true;
}

}  // namespace agp
