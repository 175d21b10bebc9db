// ITEM_POTF2 of agp_chol_kernel.  Compiled as its own translation unit (agp_chol_common.cuh says why).
#include "agp_chol_common.cuh"

namespace agp {

// ------------------------------------------------------------------------------------------
// ITEM_POTF2: Cholesky of the 128x128 diagonal tile AND its inverse W = L_kk^{-1}, in sixteen micro-panels of 8 columns.
//
//   P1  thread r (< 128) owns tile row r.  Every active thread (r >= c0) loads the 8x8 diagonal block and factors it
//       redundantly in registers (8 pivots: rsqrt, scale, rank-1 update; no shuffles, no broadcast step on the
//       pivot chain), and in lock-step with the pivots substitutes its own row's 8 panel values against it.  The
//       observation vector rides along: every thread also carries the 8 entries of y through the same substitution
//       (z = L^{-1} y) and then updates its own y entry.  Rows of the diagonal block write the factor's rows and the
//       8x8 inverse W_D.  The factor goes to global memory here, 64 bytes per row and panel.
//   P2  rank-8 updates on FP64 tensor cores, one 8x8 block per DMMA pair, work dealt out by block row rb:
//         trailing factor part   A[rb][cb] -= L[rb][panel] L[cb][panel]^T                    jp < cb <= rb
//         inverse part           V[rb][cb] -= (L[rb][panel] W_D) V[panel][cb]                cb < jp
//       where V holds, in the positions (r, c), c < r, that the finished columns of the factor left behind, the
//       running sums  -sum_m L[r][m] W[m][c]  of the inverse (W = L^{-1}: L W = I solved right-looking by rows).
//       After a CTA barrier the panel's own positions receive their starting value -(L[rb][panel] W_D), and the
//       panel's rows of the inverse are finished: W[panel][cb] = W_D V[panel][cb].
// When the last panel is done the tile holds W (its diagonal 8x8 blocks sit in Wd), which goes to global memory
// as a dense 128x128 lower-triangular matrix (zeros above the diagonal): the B operand of the panel items' triangular
// product  L_ik = (K_ik - sum_j L_ij L_kj^T) W^T.
// ------------------------------------------------------------------------------------------
constexpr int PB = 36;         // packed 32x32 block row stride: 4 mod 16 -> conflict-free DMMA fragment loads, 16-byte rows
constexpr int PBLK = 32 * PB;  // doubles per packed block; blocks (bi, bj), bj <= bi, at (bi (bi + 1) / 2 + bj) * PBLK
constexpr int NPANEL = TB / 8;
static_assert(10 * PBLK + NPANEL * 64 + 72 <= REGION_D, "packed diagonal tile + the 8x8 diagonal inverses + one 8x8 factor must fit in the region");

// 1 / sqrt(d) for a positive, normal d: hardware seed (rsqrt.approx.ftz.f64, ~2^-22) + one third-order step, no special
// cases and no out-of-line slow path on the pivot chain (libdevice's rsqrt: seed + two Newton steps + a call for the
// subnormal range).  Relative error < 2^-52.5; the pivot L_jj = d * (1 / sqrt(d)) is within 2 ulp of sqrt(d).
__device__ __forceinline__ double rsqrt_pos(double d) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
    const double t = d * y0;
    const double e = fma(-t, y0, 1.0);      // 1 - d y0^2
    const double pp = fma(0.375, e, 0.5);   // 1/2 + 3/8 e
    const double qq = y0 * e;
    return fma(qq, pp, y0);                 // y0 (1 + e/2 + 3 e^2/8)
}

__device__ __forceinline__ int toff(int r, int c) {
    const int bi = r >> 5, bj = c >> 5;
    return (bi * (bi + 1) / 2 + bj) * PBLK + (r & 31) * PB + (c & 31);
}

__device__ bool do_potf2(const BatchView& v, const SchedView& q, int idx) {
    const Smem s = smem_view();
    const int p = __ldg(&q.items[2 * idx].y), k = __ldg(&q.items[2 * idx].z), need_diag = __ldg(&q.items[2 * idx + 1].w);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, c4 = lane & 3;
    const int ld = v.ld;
    const int o = k * TB;
    double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    double* yp = v.y + (long long)p * ld;
    double* Tt = s.region;               // packed lower tile
    double* Wd = s.region + 10 * PBLK;   // [NPANEL][8][8] inverses of the 8x8 diagonal blocks
    double* ys = s.ys;                   // running y_k, then untouched
    double* zo = s.zs;                   // z_k
    double* Lg = s.Ri;                   // diagonal of L_kk (for the log det)
    double* Ld = Wd + NPANEL * 64;       // [8][8] the current diagonal block's factor

    if (tid == 0) {
        s.ctl[1] = wait_ge(q.diagu + p * q.nt_stride + k, need_diag, q.err, q.wait_timeout_ns) ? 1 : 0;
        s.ctl[2] = 0;
    }
    __syncthreads();
    if (!s.ctl[1]) return false;
    stamp(q, idx, 1);

    // lower 32x32 blocks -> packed tile: 5120 16-byte loads, 20 per thread, all in flight (L2: written by other CTAs)
    {
        double2 tmp[20];
#pragma unroll
        for (int u = 0; u < 20; ++u) {
            const int e = u * FT + tid;
            const int b = e >> 9, w = e & 511, r = w >> 4, c2 = w & 15;
            const int bi = (b >= 6) ? 3 : (b >= 3) ? 2 : (b >= 1) ? 1 : 0, bj = b - bi * (bi + 1) / 2;
            tmp[u] = __ldcg(reinterpret_cast<const double2*>(Lp + (long long)(o + bi * 32 + r) * ld + o + bj * 32 + 2 * c2));
        }
#pragma unroll
        for (int u = 0; u < 20; ++u) {
            const int e = u * FT + tid;
            const int b = e >> 9, w = e & 511, r = w >> 4, c2 = w & 15;
            *reinterpret_cast<double2*>(Tt + b * PBLK + r * PB + 2 * c2) = tmp[u];
        }
    }
    if (tid < TB) ys[tid] = __ldcg(yp + o + tid);
    __syncthreads();
    stamp(q, idx, 2);

#if AGP_X_POTF2_CLK
    long long clk_p1 = 0, clk_p2 = 0, clk_t = clock64();
#endif
#pragma unroll 1
    for (int jp = 0; jp < (AGP_X_SKIP_POTF2 ? 0 : NPANEL); ++jp) {
        const int c0 = jp * 8;
        // ---- P1 ------------------------------------------------------------------------------
        // rows of the tile below the block carry their 8 panel values; row di of the block carries the unit vector e_di
        // (its substitution result is column di of W_D = L_D^{-1}); thread 128 carries the observation entries y (result: z)
        double a[8];
        const bool p1_row = tid < TB && tid >= c0, p1_y = tid == TB;
        const bool below = tid >= c0 + 8 && tid < TB;
        double* arow = Tt + toff(below ? tid : c0, c0);
        if (p1_row || p1_y) {
            const int di = tid - c0;  // row inside the diagonal block (when 0 <= di < 8)
            const double* dblk = Tt + toff(c0, c0);
            double d[36];  // lower 8x8, (i, j) at i (i + 1) / 2 + j
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) d[i * (i + 1) / 2 + j] = dblk[i * PB + j];
            if (below) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const double2 t = *reinterpret_cast<const double2*>(arow + j);
                    a[j] = t.x;
                    a[j + 1] = t.y;
                }
            } else if (p1_y) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = ys[c0 + j];
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = (j == di) ? 1.0 : 0.0;
            }
            int bad = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                double dj = d[j * (j + 1) / 2 + j];
                if (!(dj > 0.0)) {  // also catches NaN; LAPACK dpotrf: info = column (1-based)
                    if (bad == 0) bad = o + c0 + j + 1;
                    dj = 1.0;
                }
                const double iv = rsqrt_pos(dj);
                const double ljj = dj * iv;
#pragma unroll
                for (int i = j + 1; i < 8; ++i) d[i * (i + 1) / 2 + j] *= iv;
#pragma unroll
                for (int i = j + 1; i < 8; ++i)
#pragma unroll
                    for (int m = j + 1; m <= i; ++m) d[i * (i + 1) / 2 + m] = fma(-d[i * (i + 1) / 2 + j], d[m * (m + 1) / 2 + j], d[i * (i + 1) / 2 + m]);
                const double xj = a[j] * iv;
                a[j] = xj;
#pragma unroll
                for (int m = j + 1; m < 8; ++m) a[m] = fma(-xj, d[m * (m + 1) / 2 + j], a[m]);
                if (p1_y) {  // publishes the finished column of L_D (for the factor's rows in global memory) and L_jj
                    Ld[j * 8 + j] = ljj;
#pragma unroll
                    for (int i = j + 1; i < 8; ++i) Ld[i * 8 + j] = d[i * (i + 1) / 2 + j];
#pragma unroll
                    for (int i = 0; i < j; ++i) Ld[i * 8 + j] = 0.0;
                    Lg[c0 + j] = ljj;
                }
            }
            if (below) {
                double* grow = Lp + (long long)(o + tid) * ld + o + c0;
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const double2 t = make_double2(a[j], a[j + 1]);
                    *reinterpret_cast<double2*>(arow + j) = t;
                    *reinterpret_cast<double2*>(grow + j) = t;
                }
            } else if (p1_y) {
#pragma unroll
                for (int j = 0; j < 8; ++j) zo[c0 + j] = a[j];
                if (bad != 0 && s.ctl[2] == 0) s.ctl[2] = bad;
            } else {
                double* wcol = Wd + jp * 64 + di;  // column di of W_D (zeros above the diagonal come out of the substitution)
#pragma unroll
                for (int i = 0; i < 8; ++i) wcol[i * 8] = a[i];
            }
        }
        __syncthreads();
#if AGP_X_POTF2_CLK
        { const long long t = clock64(); clk_p1 += t - clk_t; clk_t = t; }
#endif
        // ---- P2: rank-8 updates, block row rb per warp (two when more than eight are left) -----------
        if (below) {  // y_r -= L[r][panel] z_panel
            double dot = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) dot = fma(zo[c0 + j], a[j], dot);
            ys[tid] -= dot;
        }
        if (warp == FT / 32 - 1) {  // the factor's diagonal-block rows -> global memory (8 rows x 64 bytes)
            const int r = lane >> 2, c = (lane & 3) * 2;
            *reinterpret_cast<double2*>(Lp + (long long)(o + c0 + r) * ld + o + c0 + c) = *reinterpret_cast<const double2*>(Ld + r * 8 + c);
        }
        double qn[2][2];
        int rbs[2];
        const double* wd = Wd + jp * 64;
        const int bip = c0 >> 5, cc = c0 & 31, jpb = jp >> 2, jpu = jp & 3;
        // four 8x8 blocks of one 32-column group at a time (their shared-memory round trips and DMMA latencies overlap);
        // every address is a group base + a compile-time multiple of a stride
        auto grp = [&](double* cp, const double* bp, int su, int sk, double x0, double x1, int umask) {
            double2 c[4];
            double b0[4], b1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((umask >> u) & 1) {  // warp-uniform
                    b0[u] = bp[u * su];
                    b1[u] = bp[u * su + sk];
                    c[u] = *reinterpret_cast<const double2*>(cp + u * 8);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((umask >> u) & 1) {
                    dmma884(c[u].x, c[u].y, x0, b0[u]);
                    dmma884(c[u].x, c[u].y, x1, b1[u]);
                    *reinterpret_cast<double2*>(cp + u * 8) = c[u];
                }
        };
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int rb = (t == 0) ? NPANEL - 1 - warp : jp + 1 + warp;
            const bool valid = (t == 0) ? (rb > jp) : (rb < NPANEL - 8);
            rbs[t] = valid ? rb : -1;
            qn[t][0] = qn[t][1] = 0.0;
            if (valid) {
                const int bir = rb >> 2;
                const double* ap = Tt + toff(rb * 8 + g, c0 + c4);
                const double a0 = ap[0], a1 = ap[4];
                // Q = L[rb][panel] W_D  (B[n][k] = W_D[k][n])
                double q0 = 0.0, q1 = 0.0;
                dmma884(q0, q1, a0, wd[c4 * 8 + g]);
                dmma884(q0, q1, a1, wd[(c4 + 4) * 8 + g]);
                const double nq0 = -q0, nq1 = -q1;
                qn[t][0] = nq0;
                qn[t][1] = nq1;
                double* crow = Tt + (bir * (bir + 1) / 2) * PBLK + ((rb * 8 + g) & 31) * PB + 2 * c4;  // + bj PBLK + u 8
                // inverse part, column blocks cb < jp:  V[rb][cb] -= Q V[panel][cb].  The accumulator fragment of Q serves as
                // the A operand with k taken in the order the fragment holds it (lane c4: k = 2 c4, then 2 c4 + 1); B follows.
                const double* binv = Tt + (bip * (bip + 1) / 2) * PBLK + (cc + 2 * c4) * PB + g;  // + bj PBLK + u 8; next k: + PB
                for (int bj = 0; bj < jpb; ++bj) grp(crow + bj * PBLK, binv + bj * PBLK, 8, PB, nq0, nq1, 0xf);
                if (jpu > 0) grp(crow + jpb * PBLK, binv + jpb * PBLK, 8, PB, nq0, nq1, (1 << jpu) - 1);
                // factor part, column blocks jp < cb <= rb:  A[rb][cb] -= L[rb][panel] L[cb][panel]^T
                // (cb == jp is the panel itself: it starts the inverse sums from -Q after the barrier)
                const double na0 = -a0, na1 = -a1;
                const double* bfac = Tt + bip * PBLK + g * PB + cc + c4;  // + (bj (bj + 1) / 2) PBLK + u 8 PB; next k: + 4
                {
                    const int last = (bir == jpb) ? (rb & 3) : 3;  // the panel's own 32-column group
                    const int m = ((2 << last) - 1) & ~((2 << jpu) - 1);
                    if (m) grp(crow + jpb * PBLK, bfac + (jpb * (jpb + 1) / 2) * PBLK, 8 * PB, 4, na0, na1, m);
                }
                for (int bj = jpb + 1; bj <= bir; ++bj)
                    grp(crow + bj * PBLK, bfac + (bj * (bj + 1) / 2) * PBLK, 8 * PB, 4, na0, na1, bj < bir ? 0xf : (2 << (rb & 3)) - 1);
            }
        }
        __syncthreads();
        // the panel's own positions start the inverse sums; the panel's rows of the inverse are finished
#pragma unroll
        for (int t = 0; t < 2; ++t)
            if (rbs[t] >= 0) *reinterpret_cast<double2*>(Tt + toff(rbs[t] * 8 + g, c0 + 2 * c4)) = make_double2(qn[t][0], qn[t][1]);
        for (int cb = warp; cb < jp; cb += FT / 32) {
            const double* bp = Tt + toff(c0 + c4, cb * 8 + g);
            double c0v = 0.0, c1v = 0.0;
            dmma884(c0v, c1v, wd[g * 8 + c4], bp[0]);
            dmma884(c0v, c1v, wd[g * 8 + c4 + 4], bp[4 * PB]);
            *reinterpret_cast<double2*>(Tt + toff(c0 + g, cb * 8 + 2 * c4)) = make_double2(c0v, c1v);
        }
#if AGP_X_POTF2_CLK
        { const long long t = clock64(); clk_p2 += t - clk_t; clk_t = t; }
#endif
    }
#if AGP_X_POTF2_CLK
    if (q.trace != nullptr && tid == 0) q.trace[(long long)idx * 8 + 4] = (clk_p1 << 32) | (clk_p2 & 0xffffffffll);
#endif
    __syncthreads();
    stamp(q, idx, 3);

    // W -> global, dense lower triangular
    {
        double* wout = v.dinv + ((long long)p * q.nt_stride + k) * (TB * TB);
#pragma unroll 4
        for (int e = tid; e < TB * TB / 2; e += FT) {
            const int r = e >> 6, c = (e & 63) * 2;
            double2 val = make_double2(0.0, 0.0);
            if (c <= r) {
                if ((c >> 3) == (r >> 3)) {
                    const double* wdp = Wd + (r >> 3) * 64 + (r & 7) * 8 + (c & 7);
                    val = make_double2(wdp[0], wdp[1]);
                } else {
                    val = *reinterpret_cast<const double2*>(Tt + toff(r, c));
                }
            }
            *reinterpret_cast<double2*>(wout + r * TB + c) = val;
        }
    }
    // z_k, sum z^2, sum log L_jj
    if (tid < TB) {
        const double zj = zo[tid];
        v.z[(long long)p * ld + o + tid] = zj;
        double part_zz = zj * zj;
        double part_ld = log(Lg[tid]);
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
