// ITEM_POTF2 of agp_chol_kernel, round-1 form (four 32-column block steps).  Compiled as its own translation unit.
#include "agp_chol_common.cuh"

namespace agp {

constexpr int BS = 33;  // potf2 32x32 block row stride (odd: lane-per-row walks conflict free)
constexpr int BLK = 32 * BS;
static_assert(10 * BLK <= REGION_D, "packed diagonal tile must fit in the region");

// ------------------------------------------------------------------------------------------
// ITEM_POTF2: blocked right-looking Cholesky of the diagonal tile, stored as packed 32x32 blocks
//   phase 1  warp 0 factors the 32x32 diagonal block in REGISTERS (lane = row, shuffles carry the
//            pivot column)
//   phase 2  one thread per sub-diagonal row (the observation vector rides along as row 128, so
//            z_k = L_kk^{-1} y_k needs no separate solve) substitutes against the block
//   phase 3  rank-32 update of the trailing part of the tile on DMMA
// Warp 7 inverts the diagonal blocks for the panel solves behind a named barrier, off the
// critical path.
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

    const bool want_dinv = true;  // also for the last block column: a later agp_lml_run_append solves new tile rows against it
    constexpr int NW = FT / 32;             // 8 warps
    constexpr int WORKERS = (NW - 1) * 32;  // warps 0..6 factor; warp 7 inverts diagonal blocks
    if (warp == NW - 1) {
#pragma unroll 1
        for (int jb = 0; jb < 4; ++jb) {
            const int j0 = jb * 32;
            const double* Dg = Ab + blk_off(jb, jb);
            potf2_bar(1, FT);  // diagonal block jb is final
            if (want_dinv) {
                // inverse of the diagonal block, lane = column of the inverse
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
        for (int jb = 0; jb < 4; ++jb) {
            const int j0 = jb * 32;
            double* Dg = Ab + blk_off(jb, jb);
            // ---- phase 1: diagonal block in registers (warp 0) ---------------------------
            if (warp == 0) {
                double a[32];
                const double* rowp = Dg + lane * BS;
#pragma unroll
                for (int c = 0; c < 32; ++c) a[c] = rowp[c];
                int bad = 0;
                // Software-pipelined over the columns: as soon as column j is scaled, column j + 1 receives its update, and
                // the NEXT pivot (shuffle -> test -> rsqrt: ~130 of the 353 clocks a column took when this chain ran after
                // the whole rank-1 update) is computed while the remaining 30 - j updates of column j are issued.  Same
                // operations on the same operands in the same order per entry: bitwise the unpipelined loop.
                double d = __shfl_sync(0xffffffffu, a[0], 0);
                if (!(d > 0.0)) {  // also catches NaN; LAPACK dpotrf: info = j (1-based)
                    bad = o + j0 + 1;
                    d = 1.0;
                }
                double inv = rsqrt(d);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const double l = (lane == j) ? d * inv : a[j] * inv;
                    a[j] = l;
                    if (lane == 0) Ri[j0 + j] = inv;
                    if (j + 1 < 32) {
                        const double l1 = __shfl_sync(0xffffffffu, l, j + 1);
                        a[j + 1] = fma(-l, l1, a[j + 1]);
                        d = __shfl_sync(0xffffffffu, a[j + 1], j + 1);
                        if (!(d > 0.0)) {
                            if (bad == 0) bad = o + j0 + j + 2;
                            d = 1.0;
                        }
                        inv = rsqrt(d);
                    }
#pragma unroll
                    for (int c = j + 2; c < 32; ++c) {
                        const double lc = __shfl_sync(0xffffffffu, l, c);
                        a[c] = fma(-l, lc, a[c]);
                    }
                }
                double* roww = Dg + lane * BS;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (c <= lane) roww[c] = a[c];
                if (lane == 0 && bad != 0 && s.ctl[2] == 0) s.ctl[2] = bad;
            }
            potf2_bar(1, FT);
            // ---- phase 2: rows below the block, one thread per row (threads 32..) -----------
            const int R = TB + 1 - (j0 + 32);  // rows j0+32 .. 128 (row 128 = y)
            if (tid >= 32 && tid - 32 < R) {
                const int i = j0 + 32 + (tid - 32);
                double* rowp = (i < TB) ? Ab + blk_off(i >> 5, jb) + (i & 31) * BS : ys + j0;
                double a[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) a[c] = rowp[c];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const double l = a[j] * Ri[j0 + j];
                    a[j] = l;
#pragma unroll
                    for (int c = j + 1; c < 32; ++c) a[c] = fma(-l, Dg[c * BS + j], a[c]);  // broadcast
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) rowp[c] = a[c];
            }
            potf2_bar(2, WORKERS);
            // ---- phase 3: trailing update  A[i][c] -= sum_m L[i][m] L[c][m]  (DMMA) ---------
            const int T = TB - (j0 + 32);  // trailing rows/cols inside the tile
            if (T > 0) {
                const int nb8 = T >> 3;
                const int nblk = nb8 * (nb8 + 1) / 2;
                const int g = lane >> 2, c4 = lane & 3;
                for (int blk = warp; blk < nblk; blk += NW - 1) {
                    int bi = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
                    while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
                    while (bi * (bi + 1) / 2 > blk) --bi;
                    const int bc = blk - bi * (bi + 1) / 2;
                    const int ri = j0 + 32 + bi * 8 + g;  // row of the A fragment / of C
                    const int rc = j0 + 32 + bc * 8 + g;  // row of the B fragment
                    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                    const double* ap = Ab + blk_off(ri >> 5, jb) + (ri & 31) * BS + c4;
                    const double* bp = Ab + blk_off(rc >> 5, jb) + (rc & 31) * BS + c4;
#pragma unroll
                    for (int kk = 0; kk < 32; kk += 8) {
                        dmma884(c0, c1, ap[kk], bp[kk]);
                        dmma884(d0, d1, ap[kk + 4], bp[kk + 4]);
                    }
                    const int cc = j0 + 32 + bc * 8 + 2 * c4;  // column of C
                    double* cp = Ab + blk_off(ri >> 5, cc >> 5) + (ri & 31) * BS + (cc & 31);
                    cp[0] -= c0 + d0;
                    cp[1] -= c1 + d1;
                }
                // observation row: y[c] -= sum_m z_panel[m] L[c][m]
                if (warp == NW - 2) {
                    const double* zp = ys + j0;
                    for (int cc = lane; cc < T; cc += 32) {
                        const int rc = j0 + 32 + cc;
                        const double* lp = Ab + blk_off(rc >> 5, jb) + (rc & 31) * BS;
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int m = 0; m < 32; m += 2) {
                            s0 = fma(zp[m], lp[m], s0);
                            s1 = fma(zp[m + 1], lp[m + 1], s1);
                        }
                        ys[rc] -= s0 + s1;
                    }
                }
            }
            potf2_bar(3, WORKERS);
        }
    }
    __syncthreads();

    // write L_kk (lower, row-major; strictly-upper zeroed so the tile is a clean factor)
    for (int idx = tid; idx < TB * TB; idx += FT) {
        int r = idx >> 7, c = idx & (TB - 1);
        Lp[(long long)(o + r) * ld + o + c] = (c <= r) ? Ab[blk_off(r >> 5, c >> 5) + (r & 31) * BS + (c & 31)] : 0.0;
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
