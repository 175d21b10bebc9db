// ITEM_POTF2 of agp_chol_kernel (eight 16-column steps with look-ahead).  Compiled as its own translation unit.
#include "agp_chol_common.cuh"

namespace agp {

constexpr int BS = 33;  // potf2 32x32 block row stride (odd: lane-per-row walks conflict free)
constexpr int BLK = 32 * BS;
static_assert(10 * BLK <= REGION_D, "packed diagonal tile must fit in the region");

// ------------------------------------------------------------------------------------------
// ITEM_POTF2: blocked right-looking Cholesky of the diagonal tile, stored as packed 32x32 blocks, in eight steps of 16
// columns:
//   phase 1  warp 0 factors the 16x16 diagonal block in REGISTERS (lane = row, shuffles carry the pivot column, the
//            next pivot's rsqrt chain overlaps the current column's rank-1 update)
//   phase 2  one thread per sub-diagonal row (the observation vector rides along as row 128, so
//            z_k = L_kk^{-1} y_k needs no separate solve) substitutes against the block
//   phase 3  rank-16 update of the trailing part of the tile on DMMA; warp 0 updates only the NEXT diagonal block and goes
//            straight on to its phase 1 while warps 1..6 update the rest (look-ahead: the two longest phases overlap)
// Warp 7 inverts the 32x32 diagonal blocks for the panel solves behind a named barrier, off the critical path.
// Measured per item, clocks of thread 0 (n = 512, one particle, nothing else on the SM): four 32-column steps, units
// one by one: ~100 k; 16-column steps: phase 1 32 k + phase 2 8 k + phase 3 37 k; units four at a time: phase 3 28 k;
// look-ahead: see profiles/r02_potf2.txt.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int blk_off(int bi, int bj) { return (bi * (bi + 1) / 2 + bj) * BLK; }

// The named barriers of this item are reached by the factoring warps and by the inverting warp from different places of
// the code.  PTX allows that (bar.sync is per warp), compute-sanitizer's synccheck reports "divergent thread(s) in block"
// unless all of them execute the SAME barrier instruction: one out-of-line copy serves every call site.
__device__ __noinline__ void potf2_bar(int id, int count) { named_bar_sync(id, count); }

__device__ bool do_potf2(const BatchView& v, const SchedView& q, int idx) {
    const Smem s = smem_view();
    const int p = __ldg(&q.items[2 * idx].y), k = __ldg(&q.items[2 * idx].z), need_diag = __ldg(&q.items[2 * idx + 1].w);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ld = v.ld;
    const int o = k * TB;
    double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    double* yp = v.y + (long long)p * ld;
    double* Ab = s.region;
    double* ys = s.ys;
    double* Ri = s.Ri;

    if (tid == 0) {
        s.ctl[1] = wait_ge(q.diagu + p * q.nt_stride + k, need_diag, q.err, q.wait_timeout_ns) ? 1 : 0;
        s.ctl[2] = 0;
    }
    __syncthreads();
    if (!s.ctl[1]) return false;
    stamp(q, idx, 1);

    // the ten lower 32x32 blocks -> packed blocks: 5120 16-byte L2 loads (the tile was written by other CTAs of this
    // launch), 20 per thread in two batches of ten, all of a batch in flight (round 1: 64 scalar loads per thread in
    // eight dependent batches, 11 us of the item's 55)
#pragma unroll 1
    for (int u0 = 0; u0 < 20; u0 += 10) {
        double2 tmp[10];
#pragma unroll
        for (int uu = 0; uu < 10; ++uu) {
            const int e = (u0 + uu) * FT + tid;
            const int b = e >> 9, w = e & 511, r = w >> 4, c2 = w & 15;
            const int bi = (b >= 6) ? 3 : (b >= 3) ? 2 : (b >= 1) ? 1 : 0, bj = b - bi * (bi + 1) / 2;
            tmp[uu] = __ldcg(reinterpret_cast<const double2*>(Lp + (long long)(o + bi * 32 + r) * ld + o + bj * 32 + 2 * c2));
        }
#pragma unroll
        for (int uu = 0; uu < 10; ++uu) {
            const int e = (u0 + uu) * FT + tid;
            const int b = e >> 9, w = e & 511, r = w >> 4, c2 = w & 15;
            const bool dg = (b == 0) | (b == 2) | (b == 5) | (b == 9);  // a diagonal block: nothing above the diagonal
            double* dst = Ab + b * BLK + r * BS + 2 * c2;
            dst[0] = (dg && 2 * c2 > r) ? 0.0 : tmp[uu].x;
            dst[1] = (dg && 2 * c2 + 1 > r) ? 0.0 : tmp[uu].y;
        }
    }
    if (tid < TB) ys[tid] = __ldcg(yp + o + tid);
    __syncthreads();
    stamp(q, idx, 2);
    // diagnostics (tracing only): clocks thread 0 spends in the three phases, summed over the steps (its warp does phase 1
    // and phase 3 and waits through phase 2), packed into trace slot 4 as 21 bits each in units of 16 clocks
    long long tph1 = 0, tph2 = 0, tph3 = 0, tc = clock64();

    const bool want_dinv = true;  // also for the last block column: a later agp_lml_run_append solves new tile rows against it
    constexpr int NW = FT / 32;             // 8 warps
    constexpr int WORKERS = (NW - 1) * 32;  // warps 0..6 factor; warp 7 inverts diagonal blocks
    // Eight steps of 16 columns (round 1-2a: four of 32).  The in-register factorisation of the diagonal block is the
    // serial part: its cost per column grows with the block width (31 - j shuffles + FMAs behind every pivot), so halving
    // the width moves half of that work into the rank-16 DMMA update and halves the row substitution too.
    constexpr int W = 16;
    if (warp == NW - 1) {
#pragma unroll 1
        for (int sb = 0; sb < TB / W; ++sb) {
            potf2_bar(1, FT);  // columns [16 sb, 16 sb + 16) of the diagonal are final
            if (!(sb & 1)) continue;
            const int jb = sb >> 1, j0 = jb * 32;
            const double* Dg = Ab + blk_off(jb, jb);
            if (want_dinv) {
                // inverse of the 32x32 diagonal block (for the panel solves), lane = column of the inverse
                double x[32];
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    double sacc = 0.0;
#pragma unroll
                    for (int m = 0; m < r; ++m) sacc = fma(Dg[r * BS + m], x[m], sacc);  // L(r,m), broadcast
                    const double rhs = (r == lane) ? 1.0 : 0.0;
                    x[r] = (r < lane) ? 0.0 : (rhs - sacc) * Ri[j0 + r];
                }
                double* out = v.dinv + (((long long)p * q.nt_stride + k) * 4 + jb) * 1024;
#pragma unroll
                for (int r = 0; r < 32; ++r) out[r * 32 + lane] = x[r];
            }
        }
    } else {
#pragma unroll 1
        for (int sb = 0; sb < TB / W; ++sb) {
            const int jb = sb >> 1, off = (sb & 1) * W, c0 = sb * W;
            double* Dg = Ab + blk_off(jb, jb) + off * BS + off;  // the 16x16 diagonal block of this step
            // ---- phase 1: diagonal block in registers (warp 0, lane = row; lanes 16.. carry identity rows) ----
            if (warp == 0) {
                double a[W];
#pragma unroll
                for (int c = 0; c < W; ++c) a[c] = (lane < W) ? Dg[lane * BS + c] : ((c == lane - W) ? 1.0 : 0.0);
                int bad = 0;
                // software-pipelined over the columns: the next pivot (shuffle -> test -> rsqrt) is computed while the
                // remaining updates of the current column are issued
                double d = __shfl_sync(0xffffffffu, a[0], 0);
                if (!(d > 0.0)) {  // also catches NaN; LAPACK dpotrf: info = j (1-based)
                    bad = o + c0 + 1;
                    d = 1.0;
                }
                double inv = rsqrt(d);
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const double l = (lane == j) ? d * inv : a[j] * inv;
                    a[j] = l;
                    if (lane == 0) Ri[c0 + j] = inv;
                    if (j + 1 < W) {
                        const double l1 = __shfl_sync(0xffffffffu, l, j + 1);
                        a[j + 1] = fma(-l, l1, a[j + 1]);
                        d = __shfl_sync(0xffffffffu, a[j + 1], j + 1);
                        if (!(d > 0.0)) {
                            if (bad == 0) bad = o + c0 + j + 2;
                            d = 1.0;
                        }
                        inv = rsqrt(d);
                    }
#pragma unroll
                    for (int c = j + 2; c < W; ++c) {
                        const double lc = __shfl_sync(0xffffffffu, l, c);
                        a[c] = fma(-l, lc, a[c]);
                    }
                }
                if (lane < W) {
#pragma unroll
                    for (int c = 0; c < W; ++c)
                        if (c <= lane) Dg[lane * BS + c] = a[c];
                }
                if (lane == 0 && bad != 0 && s.ctl[2] == 0) s.ctl[2] = bad;
            }
            potf2_bar(1, FT);
            if (q.trace != nullptr && tid == 0) { const long long t = clock64(); tph1 += t - tc; tc = t; }
            // ---- phase 2: rows below the block, one thread per row (threads 32..); row 128 = the observation vector ----
            const int R = TB + 1 - (c0 + W);
            if (tid >= 32 && tid - 32 < R) {
                const int i = c0 + W + (tid - 32);
                double* rowp = (i < TB) ? Ab + blk_off(i >> 5, jb) + (i & 31) * BS + off : ys + c0;
                double a[W];
#pragma unroll
                for (int c = 0; c < W; ++c) a[c] = rowp[c];
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const double l = a[j] * Ri[c0 + j];
                    a[j] = l;
#pragma unroll
                    for (int c = j + 1; c < W; ++c) a[c] = fma(-l, Dg[c * BS + j], a[c]);  // broadcast
                }
#pragma unroll
                for (int c = 0; c < W; ++c) rowp[c] = a[c];
            }
            potf2_bar(2, WORKERS);
            if (q.trace != nullptr && tid == 0) { const long long t = clock64(); tph2 += t - tc; tc = t; }
            // ---- phase 3: trailing update  A[i][c] -= sum_{m<16} L[i][c0+m] L[c][c0+m]  (DMMA) ----
            const int T = TB - (c0 + W);  // trailing rows/cols inside the tile
            if (T > 0) {
                // 8x8 units of the trailing lower triangle, unit row bi has bi + 1 units.  LOOK-AHEAD: unit rows 0 and 1 are
                // the next step's diagonal block; warp 0 updates those three units and goes straight on to factor that block
                // (phase 1 of the next step) while warps 1..6 update the rest — the two longest phases of a step overlap.
                // The other unit rows are dealt out in pairs (2 + pr, nb8 - 1 - pr): nb8 + 3 units per pair, nb8 is even, one
                // pair per warp and round; within a unit row the A fragments are loaded once and the units go four at a
                // time, so that eight DMMA chains and their shared-memory round trips are in flight together (the first
                // version walked the units one by one behind a square-root index: 36 k of the item's 77 k clocks).
                const int nb8 = T >> 3;
                const int g = lane >> 2, c4 = lane & 3;
                auto unit_row = [&](int bi) {
                    const int ri = c0 + W + bi * 8 + g;  // row of the A fragment / of C
                    const double* ap = Ab + blk_off(ri >> 5, jb) + (ri & 31) * BS + off + c4;
                    const double a0 = ap[0], a1 = ap[4], a2 = ap[8], a3 = ap[12];
#pragma unroll 1
                    for (int bc0 = 0; bc0 <= bi; bc0 += 4) {
                        double e[4][2], d[4][2];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            e[u][0] = e[u][1] = d[u][0] = d[u][1] = 0.0;
                            if (bc0 + u <= bi) {
                                const int rc = c0 + W + (bc0 + u) * 8 + g;  // row of the B fragment
                                const double* bp = Ab + blk_off(rc >> 5, jb) + (rc & 31) * BS + off + c4;
                                const double b0 = bp[0], b1 = bp[4], b2 = bp[8], b3 = bp[12];
                                dmma884(e[u][0], e[u][1], a0, b0);
                                dmma884(d[u][0], d[u][1], a1, b1);
                                dmma884(e[u][0], e[u][1], a2, b2);
                                dmma884(d[u][0], d[u][1], a3, b3);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (bc0 + u <= bi) {
                                const int cc = c0 + W + (bc0 + u) * 8 + 2 * c4;  // column of C
                                double* cp = Ab + blk_off(ri >> 5, cc >> 5) + (ri & 31) * BS + (cc & 31);
                                cp[0] -= e[u][0] + d[u][0];
                                cp[1] -= e[u][1] + d[u][1];
                            }
                    }
                };
                if (warp == 0) {
                    unit_row(0);
                    unit_row(1);
                    __syncwarp();  // phase 1 of the next step reads these units row by row
                } else {
                    for (int pr = warp - 1; pr < ((nb8 - 2) >> 1); pr += NW - 2) {
                        unit_row(2 + pr);
                        unit_row(nb8 - 1 - pr);
                    }
                }
                // observation row: y[c] -= sum_m z_step[m] L[c][c0+m]
                if (warp == NW - 2) {
                    const double* zp = ys + c0;
                    for (int cc = lane; cc < T; cc += 32) {
                        const int rc = c0 + W + cc;
                        const double* lp = Ab + blk_off(rc >> 5, jb) + (rc & 31) * BS + off;
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int m = 0; m < W; m += 2) {
                            s0 = fma(zp[m], lp[m], s0);
                            s1 = fma(zp[m + 1], lp[m + 1], s1);
                        }
                        ys[rc] -= s0 + s1;
                    }
                }
            }
            // no barrier here: the next step's first barrier joins warp 0 (ahead in phase 1) and the updating warps
            if (q.trace != nullptr && tid == 0) { const long long t = clock64(); tph3 += t - tc; tc = t; }
        }
        if (q.trace != nullptr && tid == 0)
            q.trace[(long long)idx * 8 + 4] = (((tph1 >> 4) & 0x1fffff) << 42) | (((tph2 >> 4) & 0x1fffff) << 21) | ((tph3 >> 4) & 0x1fffff);
    }
    __syncthreads();
    stamp(q, idx, 3);

    // write L_kk: the ten lower 32x32 blocks as 16-byte stores, 20 per thread (the strictly upper part of the diagonal
    // blocks zeroed, so the tile is a clean factor; the six blocks above the diagonal hold the Gram fill's zeros and no item
    // ever writes them).  This store sits on every particle's critical path: POTF2 -> panel solve -> next diagonal tile.
#pragma unroll 4
    for (int u = 0; u < 20; ++u) {
        const int e = u * FT + tid;
        const int b = e >> 9, w = e & 511, r = w >> 4, c2 = w & 15;
        const int bi = (b >= 6) ? 3 : (b >= 3) ? 2 : (b >= 1) ? 1 : 0, bj = b - bi * (bi + 1) / 2;
        const double* src = Ab + b * BLK + r * BS + 2 * c2;
        const bool dg = bi == bj;
        const double2 val = make_double2((dg && 2 * c2 > r) ? 0.0 : src[0], (dg && 2 * c2 + 1 > r) ? 0.0 : src[1]);
        *reinterpret_cast<double2*>(Lp + (long long)(o + bi * 32 + r) * ld + o + bj * 32 + 2 * c2) = val;
    }
    // z_k, sum z^2, sum log L_jj
    if (tid < TB) {
        double zj = ys[tid];
        v.z[(long long)p * ld + o + tid] = zj;
        double part_zz = zj * zj;
        double part_ld = log(Ab[blk_off(tid >> 5, tid >> 5) + (tid & 31) * BS + (tid & 31)]);
        part_ld = warp_sum(part_ld);
        part_zz = warp_sum(part_zz);
        if (lane == 0) {
            s.red[warp * 2] = part_ld;
            s.red[warp * 2 + 1] = part_zz;
        }
    }
    __syncthreads();
    if (tid == 0) {
        double sl = ((s.red[0] + s.red[2]) + s.red[4]) + s.red[6];
        double sz = ((s.red[1] + s.red[3]) + s.red[5]) + s.red[7];
        // running sums per block column (a later call may continue the factorisation from any column)
        double* cum = v.cum + ((long long)p * q.nt_stride + k) * 2;
        double tot_l = (k == 0 ? 0.0 : __ldcg(cum - 2)) + sl;
        double tot_z = (k == 0 ? 0.0 : __ldcg(cum - 1)) + sz;
        cum[0] = tot_l;
        cum[1] = tot_z;
        int info = (k == 0) ? 0 : __ldcg(v.info + p);
        if (info == 0 && s.ctl[2] != 0) info = s.ctl[2];
        v.info[p] = info;
        if (k == v.nt - 1) {
            // -(n log 2pi + logdet)/2 - z'z/2, logdet = 2 sum log L_ii
            const double log2pi = 1.8378770664093453;
            double lml = -0.5 * ((double)v.n * log2pi + 2.0 * tot_l) - 0.5 * tot_z;
            v.lml[p] = (info == 0) ? lml : __longlong_as_double(0x7ff8000000000000LL);
        }
    }
    signal_done(q.fdone + p);
    return true;
}


}  // namespace agp
