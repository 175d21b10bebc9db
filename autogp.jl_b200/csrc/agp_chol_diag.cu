// ITEM_DIAG of agp_chol_kernel: the lower triangle of a diagonal 128x128 tile,
//     K(ts_k, ts_k) + noise I - sum_{j0 <= j < j1} L_kj L_kj^T,
// dealt out as 16x16 blocks.  Compiled as its own translation unit and called through the plain ABI, like POTF2
// (agp_chol_common.cuh says why: the panel item's main loop keeps its register allocation whatever happens here).
//
// Rounds 1-2a ran a diagonal tile through the panel item's 32x32 warp tiles: two items of 64 rows, 3 and 7 of their 8
// warps busy (the tiles strictly above the diagonal skipped), each item as long as a full panel contraction.  Only 36 of
// the tile's 64 16x16 blocks are needed, and of the eight blocks ON the diagonal only three of their four 8x8 units:
// 136 units of work where the panel shape spends 16 units of time on each of 16 warps.  Here the 36 blocks are dealt
// out over the same 2 x 8 warps: twelve warps take two off-diagonal blocks (8 units), four take two diagonal blocks and
// one off-diagonal block (10 units), so an item lasts 10/16 of a panel contraction of the same depth.  The two items of
// a tile still exist (the queue, its counters and POTF2's wait are unchanged); they split the blocks, not the rows.
// Both read the whole tile row k (B operand = A operand: 128 rows x 16 columns per pipeline chunk, 2-D TMA).
#include "agp_chol_common.cuh"

namespace agp {

namespace {

// [item half][warp][slot] = block row << 4 | block column (16x16 blocks of the tile), 0xff = no block
__constant__ unsigned char kDiagBlocks[2][8][3] = {
    {{0x31, 0x32, 0xff}, {0x40, 0x41, 0xff}, {0x42, 0x43, 0xff}, {0x50, 0x51, 0xff}, {0x52, 0x53, 0xff}, {0x54, 0x60, 0xff},
     {0x00, 0x11, 0x10}, {0x22, 0x33, 0x20}},
    {{0x61, 0x62, 0xff}, {0x63, 0x64, 0xff}, {0x65, 0x70, 0xff}, {0x71, 0x72, 0xff}, {0x73, 0x74, 0xff}, {0x75, 0x76, 0xff},
     {0x44, 0x55, 0x21}, {0x66, 0x77, 0x30}},
};

}  // namespace

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
    int rb[3], cb[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int e = kDiagBlocks[h][warp][j];
        rb[j] = (e == 0xff) ? -1 : (e >> 4);
        cb[j] = e & 15;
    }

    const int nchunk = ((j1 - j0) * TB) / KC;
    const int G0 = s.ctl[4];
    const int ccol = j0 * TB, brow = p * ld + col0;
    auto produce = [&](int c) {  // thread 0 only
        const int G = G0 + c, st = G % NSTAGE;
        if (G >= NSTAGE && !mbar_wait_bounded(s.empty + st, ((G / NSTAGE) - 1) & 1, q.err, q.wait_timeout_ns)) return;
        double* Bs = stages + st * STAGE_D;
        mbar_expect_tx(s.full + st, UN * KC * 8);
        tma_load_2d(Bs, &maps.b, ccol + c * KC, brow, s.full + st);
    };
    if (tid == 0) {
        fence_proxy_async_all();  // after the acquire of the dependency counters, before this item's first async-proxy reads of L
        for (int c = 0; c < NSTAGE - 1 && c < nchunk; ++c) produce(c);
    }
    // accumulators start from minus the tile (Gram values, or the partial sums of an earlier item of this tile half)
    double acc[3][2][2][2];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (rb[j] >= 0) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int r = rb[j] * 16 + u * 8 + g, c = cb[j] * 16 + w * 8 + 2 * c4;
                    const double2 kv = __ldcg(reinterpret_cast<const double2*>(Lp + (long long)(col0 + r) * ld + col0 + c));
                    acc[j][u][w][0] = -kv.x;
                    acc[j][u][w][1] = -kv.y;
                }
        }

    for (int ch = 0; ch < nchunk; ++ch) {
        const int G = G0 + ch, st = G % NSTAGE;
        if (tid == 0 && ch + NSTAGE - 1 < nchunk) produce(ch + NSTAGE - 1);
        if (!mbar_wait_bounded(s.full + st, (G / NSTAGE) & 1, q.err, q.wait_timeout_ns)) return false;
        const double* Bs = stages + st * STAGE_D;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (rb[j] >= 0) {
                    double2 a[2], b[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        a[u] = *reinterpret_cast<const double2*>(Bs + swz128(rb[j] * 16 + u * 8 + g, 2 * c4 + ks));
                        b[u] = *reinterpret_cast<const double2*>(Bs + swz128(cb[j] * 16 + u * 8 + g, 2 * c4 + ks));
                    }
                    const bool dg = rb[j] == cb[j];  // a block on the diagonal: its upper-right 8x8 unit is never read
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
                            if (dg && u == 0 && w == 1) continue;
                            dmma884(acc[j][u][w][0], acc[j][u][w][1], a[u].x, b[w].x);
                            dmma884(acc[j][u][w][0], acc[j][u][w][1], a[u].y, b[w].y);
                        }
                }
            if (ks == 1) {
                // release of the stage: every lane orders its generic-proxy reads before the async-proxy write of the next
                // box into it (profiles/r02_race_experiments.txt; tools/sass_lint.py checks the built code)
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(s.empty + st);
            }
        }
    }
    if (tid == 0) s.ctl[4] = G0 + nchunk;
    stamp(q, idx, 2);

    // the lower part goes back to L, for POTF2 or for the item that continues the contraction
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (rb[j] >= 0) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int r = rb[j] * 16 + u * 8 + g, c = cb[j] * 16 + w * 8 + 2 * c4;
                    double* dst = Lp + (long long)(col0 + r) * ld + col0 + c;
                    if (c + 1 <= r) *reinterpret_cast<double2*>(dst) = make_double2(-acc[j][u][w][0], -acc[j][u][w][1]);
                    else if (c <= r) dst[0] = -acc[j][u][w][0];
                }
        }
    signal_done(q.diagu + p * q.nt_stride + k);
    return true;
}

}  // namespace agp
