// sm_100a kernels of the fused GP log-marginal-likelihood path.
//
//   update : tile (i,k) of block column k  <-  K(ts_i, ts_k) [+ noise I]  -  sum_{j<k} L_ij L_kj^T
//            The Gram tile is generated on the fly from the kernel-tree program (K is never
//            written to HBM on this path); the contraction runs on FP64 tensor cores
//            (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind) fed by a
//            3-stage cp.async pipeline; ts slices arrive by 1-D TMA bulk copies.
//   potf2  : 128x128 diagonal tile Cholesky in shared memory, augmented with the observation
//            row so z_k = L_kk^{-1} y_k falls out of the same sweep; accumulates log det and
//            z'z; LAPACK-style info; inverts the four 32x32 diagonal blocks for trsm.
//   trsm   : L_ik = C_ik L_kk^{-T} by blocked substitution on DMMA, then y_i -= L_ik z_k
//            (the forward solve rides along the factorisation sweep).
//   gram   : stand-alone K(ts,ts) + noise I, column-major, both triangles (HBM-write bound).
//
// Reference semantics: src/GP.jl:137-503, 666-684; src/Model.jl:134-136; Distributions'
// MvNormal logpdf = -(n log 2pi + logdet)/2 - |U^{-T} x|^2 / 2 with K = U'U (upper Cholesky).
// Our row-major lower factor L is bit-for-bit the column-major upper factor U = L'.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "agp_eval.cuh"
#include "agp_kernels.cuh"
#include "agp_ptx.cuh"

namespace agp {

// ------------------------------------------------------------------------------------------
// update kernel: one CTA = 128 rows x 64 columns of tile (i,k); two CTAs are resident per SM so
// that one CTA's Gram/store epilogue (FP64 ALU + LSU) overlaps the other's DMMA main loop.
// ------------------------------------------------------------------------------------------
constexpr int TBN = 64;            // CTA tile columns
constexpr int KC = 16;             // K-chunk per pipeline stage (doubles) = one 128-byte row
constexpr int NSTAGE = 4;
constexpr int UPD_THREADS = 256;   // 8 warps: 4 (m) x 2 (n), warp tile 32x32
constexpr int CS_STRIDE = TBN + 8; // accumulator staging row stride (doubles): 72 = 8 mod 16
constexpr int PROG_SMEM = 64;      // instructions cached in shared memory
constexpr int STAGE_DOUBLES = (TB + TBN) * KC;
constexpr int UPD_SMEM_STAGES = NSTAGE * STAGE_DOUBLES * 8;                                // 98304
constexpr int UPD_SMEM_BYTES = UPD_SMEM_STAGES + (TB + TBN) * 8 + PROG_SMEM * 32 + 64;     // + ts_r, ts_c, program, mbarrier

static_assert(TB * CS_STRIDE * 8 <= UPD_SMEM_STAGES, "accumulator staging must fit in the pipeline buffers");
static_assert(2 * (UPD_SMEM_BYTES + 1024) <= 227 * 1024, "two CTAs per SM");

// element (row, chunk) of a [rows][KC] operand tile; the eight 16-byte chunks of a row are
// swizzled so that the LDS.128 fragment loads of two adjacent rows hit disjoint bank halves
__device__ __forceinline__ int swz(int row, int chunk) { return row * KC + ((chunk ^ ((row & 1) << 2)) << 1); }

__global__ void __launch_bounds__(UPD_THREADS, 2) agp_update_kernel(BatchView v, int k) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* ts_r = reinterpret_cast<double*>(smem_raw + UPD_SMEM_STAGES);
    double* ts_c = ts_r + TB;
    AgpInstr* prog_s = reinterpret_cast<AgpInstr*>(ts_c + TBN);
    uint64_t* bar = reinterpret_cast<uint64_t*>(prog_s + PROG_SMEM);

    const int tid = threadIdx.x;
    const int p = v.p0 + blockIdx.y;
    const int it = k + (blockIdx.x >> 1);  // tile row
    const int half = blockIdx.x & 1;       // column half of the tile
    const bool diag = (it == k);
    const int row0 = it * TB, col0 = k * TB + half * TBN;
    double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    const int ld = v.ld;

    // --- stage ts slices with TMA bulk copies; program into shared memory -----------------
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (TB + TBN) * 8);
        tma_bulk_g2s(ts_r, v.ts + row0, TB * 8, bar);
        tma_bulk_g2s(ts_c, v.ts + col0, TBN * 8, bar);
    }
    const int poff = v.prog_off[p];
    const int pm = v.prog_off[p + 1] - poff;
    const AgpInstr* prog = v.prog + poff;
    if (pm <= PROG_SMEM) {
        // 32-byte instructions = 4 x 8-byte words
        const double* src = reinterpret_cast<const double*>(prog);
        double* dst = reinterpret_cast<double*>(prog_s);
        for (int q = tid; q < pm * 4; q += UPD_THREADS) dst[q] = src[q];
        prog = prog_s;
    }
    if (k == 0 && half == 0) {
        // first touch of this batch: reset the forward-solve vector and the accumulators
        double* yp = v.y + (long long)p * ld;
        for (int r = tid; r < TB; r += UPD_THREADS) yp[row0 + r] = (row0 + r < v.n) ? v.xs[row0 + r] : 0.0;
        if (it == 0 && tid == 0) {
            v.logdet_half[p] = 0.0;
            v.zz[p] = 0.0;
            v.info[p] = 0;
        }
    }

    // --- contraction: acc = sum_{j<k} L_ij L_kj^T over K = k*TB --------------------------
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, c4 = lane & 3;
    // warp tiles strictly above the diagonal of a diagonal tile are never read: skip their math.
    // (warp w sits on scheduler w%4, so the active warps of a half-empty CTA still spread over
    // all four schedulers and the co-resident CTA picks up the freed tensor-pipe time)
    const bool active = !diag || (wm * 32 + 31 >= half * TBN + wn * 32);
    double acc[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;

    const int nchunk = (k * TB) / KC;
    const double* __restrict__ Ag = Lp + (long long)row0 * ld;
    const double* __restrict__ Bg = Lp + (long long)col0 * ld;

    auto load_stage = [&](int s, int chunk) {
        double* As = stages + s * STAGE_DOUBLES;
        double* Bs = As + TB * KC;
        const int kk0 = chunk * KC;
#pragma unroll
        for (int e = 0; e < (TB * KC / 2) / UPD_THREADS; ++e) {  // 4
            int q = tid + e * UPD_THREADS;
            int row = q >> 3, ch = q & 7;
            cp_async16(As + swz(row, ch), Ag + (long long)row * ld + kk0 + ch * 2);
        }
        if (!diag) {
#pragma unroll
            for (int e = 0; e < (TBN * KC / 2) / UPD_THREADS; ++e) {  // 2
                int q = tid + e * UPD_THREADS;
                int row = q >> 3, ch = q & 7;
                cp_async16(Bs + swz(row, ch), Bg + (long long)row * ld + kk0 + ch * 2);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (s < nchunk) load_stage(s, s);
        cp_async_commit();
    }
    for (int ch = 0; ch < nchunk; ++ch) {
        cp_async_wait<NSTAGE - 2>();
        __syncthreads();
        {
            int nxt = ch + NSTAGE - 1;
            if (nxt < nchunk) load_stage(nxt % NSTAGE, nxt);
            cp_async_commit();
        }
        if (active) {
            const double* As = stages + (ch % NSTAGE) * STAGE_DOUBLES;
            const double* Bs = diag ? As + half * TBN * KC : As + TB * KC;  // diagonal tile: B rows are a slice of A
#pragma unroll
            for (int ks = 0; ks < KC / 8; ++ks) {
                double2 a[4], b[4];
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) a[mb] = *reinterpret_cast<const double2*>(As + swz(wm * 32 + mb * 8 + g, ks * 4 + c4));
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) b[nb] = *reinterpret_cast<const double2*>(Bs + swz(wn * 32 + nb * 8 + g, ks * 4 + c4));
                // two independent passes over the 16 accumulators: consecutive DMMAs never
                // depend on each other (dependency distance = 16 instructions)
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].x, b[nb].x);
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].y, b[nb].y);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // --- epilogue: stage accumulators, then out = K(ts_r, ts_c) - acc, coalesced ----------
    double* Cs = stages;
    if (k > 0) {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                int r = wm * 32 + mb * 8 + g, c = wn * 32 + nb * 8 + 2 * c4;
                *reinterpret_cast<double2*>(Cs + r * CS_STRIDE + c) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
            }
    }
    mbar_wait(bar, 0);
    __syncthreads();

    const int need = v.prog_need[p];
    const double noise = v.noise[p];
    const int n = v.n;
    // thread -> column c, rows rbase + 4*e (e = 0..31), evaluated four entries at a time
    const int c = tid & (TBN - 1), rbase = tid >> 6;
    const int gc = col0 + c;
    const int cdiag = half * TBN + c;  // column index inside the 128x128 tile
    const double tcol = ts_c[c];
#pragma unroll 1
    for (int e4 = 0; e4 < 8; ++e4) {
        const int rlast = rbase + 4 * (4 * e4 + 3);
        if (diag && cdiag > rlast) continue;  // strictly-upper part of a diagonal tile is never read
        double t1[4], t2[4], val[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            t1[j] = tcol;  // upper-triangle element (gc, gr): row index gc <= gr
            t2[j] = ts_r[rbase + 4 * (4 * e4 + j)];
        }
        eval_entries<4>(prog, pm, need, t1, t2, 0, val);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = rbase + 4 * (4 * e4 + j);
            const int gr = row0 + r;
            if (diag && cdiag > r) continue;
            double out;
            if (gr < n) {  // gc <= gr < n
                out = val[j];
                if (gr == gc) out = out + noise;  // + noise*I, src/GP.jl:667
            } else {
                out = (gr == gc) ? 1.0 : 0.0;  // padding: identity block, contributes log 1 = 0
            }
            if (k > 0) out = out - Cs[r * CS_STRIDE + c];
            Lp[(long long)gr * ld + gc] = out;
        }
    }
}

// ------------------------------------------------------------------------------------------
// potf2 kernel: one CTA per particle, diagonal tile k (+ observation row)
//
// Blocked right-looking Cholesky of the 128x128 tile in shared memory, 32-wide panels:
//   phase 1  warp 0 factors the 32x32 diagonal block in REGISTERS (lane = row, shuffles carry the
//            pivot column) — the only inherently serial chain: 32 x (shfl, rsqrt, mul, fma)
//   phase 2  one thread per sub-diagonal row (the observation vector y rides along as row 128,
//            so z_k = L_kk^{-1} y_k needs no separate solve) substitutes against the block;
//            meanwhile warp 1 inverts the diagonal block for the trsm kernel
//   phase 3  rank-32 update of the trailing part of the tile on DMMA
// ------------------------------------------------------------------------------------------
constexpr int PF_THREADS = 512;
constexpr int SA = TB + 1;   // 129: odd stride, lane-per-row walks are conflict free; row TB = y
constexpr int LPS = 36;      // panel staging stride (4 mod 16 doubles): DMMA fragment loads conflict free
constexpr int PF_SMEM_BYTES = ((TB + 1) * SA + (TB - 32 + 1) * LPS + TB + 16) * 8;

__global__ void __launch_bounds__(PF_THREADS, 1) agp_potf2_kernel(BatchView v, int k) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);  // [TB+1][SA] lower triangle + y row
    double* Lpn = As + (TB + 1) * SA;                   // [97][LPS] current panel, rows below the diagonal block
    double* Ri = Lpn + (TB - 32 + 1) * LPS;            // [TB] 1 / L_jj
    double* red = Ri + TB;                             // reduction scratch
    __shared__ int bad_s;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int p = v.p0 + blockIdx.x;
    const int ld = v.ld;
    const int o = k * TB;
    double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    double* yp = v.y + (long long)p * ld;

    if (tid == 0) bad_s = 0;
    // lower triangle + observation row, all loads in flight at once (8-byte cp.async: the odd
    // row stride rules out 16-byte copies)
    for (int idx = tid; idx < TB * TB; idx += PF_THREADS) {
        int r = idx >> 7, c = idx & (TB - 1);
        if (c <= r) cp_async8(As + r * SA + c, Lp + (long long)(o + r) * ld + o + c);
        else As[r * SA + c] = 0.0;
    }
    for (int c = tid; c < TB; c += PF_THREADS) cp_async8(As + TB * SA + c, yp + o + c);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const bool want_dinv = (k < v.nt - 1);
    constexpr int NW = PF_THREADS / 32;      // 16 warps
    constexpr int WORKERS = (NW - 1) * 32;   // warps 0..14 factor; warp 15 inverts diagonal blocks
    // Barriers: id 1 = "diagonal block jb is final" (all 16 warps); ids 2, 3 = phase boundaries of
    // the 15 factor warps.  The inverse warp only joins barrier 1, so inverting block jb overlaps
    // phases 2, 3 of panel jb and phase 1 of panel jb+1 instead of sitting on the critical path.
    if (warp == NW - 1) {
#pragma unroll 1
        for (int jb = 0; jb < 4; ++jb) {
            const int j0 = jb * 32;
            named_bar_sync(1, PF_THREADS);
            if (want_dinv) {
                // inverse of the diagonal block, lane = column of the inverse
                double x[32];
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    double sacc = 0.0;
#pragma unroll
                    for (int m = 0; m < r; ++m) sacc = fma(As[(j0 + r) * SA + j0 + m], x[m], sacc);  // L(r,m), broadcast
                    const double rhs = (r == lane) ? 1.0 : 0.0;
                    x[r] = (r < lane) ? 0.0 : (rhs - sacc) * Ri[j0 + r];
                }
                double* out = v.dinv + ((long long)p * 4 + jb) * 1024;
#pragma unroll
                for (int r = 0; r < 32; ++r) out[r * 32 + lane] = x[r];
            }
        }
    } else {
#pragma unroll 1
        for (int jb = 0; jb < 4; ++jb) {
            const int j0 = jb * 32;
            // ---- phase 1: diagonal block in registers (warp 0) ---------------------------
            if (warp == 0) {
                double a[32];
                const double* rowp = As + (j0 + lane) * SA + j0;
#pragma unroll
                for (int c = 0; c < 32; ++c) a[c] = rowp[c];
                int bad = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    double d = __shfl_sync(0xffffffffu, a[j], j);
                    if (!(d > 0.0)) {  // also catches NaN; LAPACK dpotrf: info = j (1-based)
                        if (bad == 0) bad = o + j0 + j + 1;
                        d = 1.0;
                    }
                    const double inv = rsqrt(d);
                    const double l = (lane == j) ? d * inv : a[j] * inv;
                    a[j] = l;
                    if (lane == 0) Ri[j0 + j] = inv;
#pragma unroll
                    for (int c = j + 1; c < 32; ++c) {
                        const double lc = __shfl_sync(0xffffffffu, l, c);
                        a[c] = fma(-l, lc, a[c]);
                    }
                }
                double* roww = As + (j0 + lane) * SA + j0;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (c <= lane) roww[c] = a[c];
                if (lane == 0 && bad != 0 && bad_s == 0) bad_s = bad;
            }
            named_bar_sync(1, PF_THREADS);
            // ---- phase 2: rows below the block, one thread per row (threads 32..) -----------
            const int R = TB + 1 - (j0 + 32);  // rows j0+32 .. 128 (row 128 = y)
            if (tid >= 32 && tid - 32 < R) {
                const int t = tid - 32;
                const int i = j0 + 32 + t;
                double a[32];
                double* rowp = As + i * SA + j0;
#pragma unroll
                for (int c = 0; c < 32; ++c) a[c] = rowp[c];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const double l = a[j] * Ri[j0 + j];
                    a[j] = l;
#pragma unroll
                    for (int c = j + 1; c < 32; ++c) a[c] = fma(-l, As[(j0 + c) * SA + j0 + j], a[c]);  // broadcast
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    rowp[c] = a[c];
                    Lpn[t * LPS + c] = a[c];
                }
            }
            named_bar_sync(2, WORKERS);
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
                    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                    const double* ap = Lpn + (bi * 8 + g) * LPS + c4;
                    const double* bp = Lpn + (bc * 8 + g) * LPS + c4;
#pragma unroll
                    for (int kk = 0; kk < 32; kk += 8) {
                        dmma884(c0, c1, ap[kk], bp[kk]);
                        dmma884(d0, d1, ap[kk + 4], bp[kk + 4]);
                    }
                    double* cp = As + (j0 + 32 + bi * 8 + g) * SA + j0 + 32 + bc * 8 + 2 * c4;
                    cp[0] -= c0 + d0;
                    cp[1] -= c1 + d1;
                }
                // observation row (t = T): y[c] -= sum_m z_panel[m] L[c][m]
                if (warp == NW - 2) {
                    const double* zp = Lpn + T * LPS;
                    for (int cc = lane; cc < T; cc += 32) {
                        const double* lp = Lpn + cc * LPS;
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int m = 0; m < 32; m += 2) {
                            s0 = fma(zp[m], lp[m], s0);
                            s1 = fma(zp[m + 1], lp[m + 1], s1);
                        }
                        As[TB * SA + j0 + 32 + cc] -= s0 + s1;
                    }
                }
            }
            named_bar_sync(3, WORKERS);
        }
    }
    __syncthreads();

    // write L_kk (lower, row-major; strictly-upper zeroed so the tile is a clean factor)
    for (int idx = tid; idx < TB * TB; idx += PF_THREADS) {
        int r = idx >> 7, c = idx & (TB - 1);
        Lp[(long long)(o + r) * ld + o + c] = (c <= r) ? As[r * SA + c] : 0.0;
    }
    // z_k, sum z^2, sum log L_jj
    double part_ld = 0.0, part_zz = 0.0;
    if (tid < TB) {
        double zj = As[TB * SA + tid];
        v.z[(long long)p * ld + o + tid] = zj;
        part_zz = zj * zj;
        part_ld = log(As[tid * SA + tid]);
        part_ld = warp_sum(part_ld);
        part_zz = warp_sum(part_zz);
        if (lane == 0) {
            red[warp * 2] = part_ld;
            red[warp * 2 + 1] = part_zz;
        }
    }
    __syncthreads();
    if (tid == 0) {
        double sl = ((red[0] + red[2]) + red[4]) + red[6];
        double sz = ((red[1] + red[3]) + red[5]) + red[7];
        double tot_l = v.logdet_half[p] + sl;
        double tot_z = v.zz[p] + sz;
        v.logdet_half[p] = tot_l;
        v.zz[p] = tot_z;
        int info = v.info[p];
        if (info == 0 && bad_s != 0) {
            info = bad_s;
            v.info[p] = info;
        }
        if (k == v.nt - 1) {
            // -(n log 2pi + logdet)/2 - z'z/2, logdet = 2 sum log L_ii
            const double log2pi = 1.8378770664093453;
            double lml = -0.5 * ((double)v.n * log2pi + 2.0 * tot_l) - 0.5 * tot_z;
            v.lml[p] = (info == 0) ? lml : __longlong_as_double(0x7ff8000000000000LL);
        }
    }
}

// ------------------------------------------------------------------------------------------
// trsm kernel: 64 rows of tile (i,k) per CTA.  Blocked substitution over the four 32-column
// blocks of L_kk:  X_jb = (C_jb - sum_{m<jb} X_m L[jb,m]^T) inv(L[jb,jb])^T, all on DMMA.
// Each warp owns 8 rows for the whole sweep, so the block-to-block dependency is warp-local.
// Operands arrive in four cp.async groups (one per column block) so the first block's math
// starts while the rest of L_kk is still in flight.
// ------------------------------------------------------------------------------------------
constexpr int TR_THREADS = 256;
constexpr int TR_ROWS = 64;
constexpr int XS = 136;  // stride = 8 mod 16 doubles: LDS.128 / STS.128 of two adjacent rows hit disjoint bank halves
constexpr int TR_SMEM_BYTES = (TR_ROWS * XS + TB * XS + TB) * 8;

__global__ void __launch_bounds__(TR_THREADS, 1) agp_trsm_kernel(BatchView v, int k) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Xs = reinterpret_cast<double*>(smem_raw);  // [64][XS]   C tile rows -> X
    double* Ls = Xs + TR_ROWS * XS;                    // [128][XS]  L_kk (diag 32x32 blocks replaced by their inverses)
    double* zs = Ls + TB * XS;                         // [128]

    const int tid = threadIdx.x;
    const int p = v.p0 + blockIdx.y;
    const int it = k + 1 + (blockIdx.x >> 1);
    const int r0 = it * TB + (blockIdx.x & 1) * TR_ROWS;
    const int o = k * TB;
    const int ld = v.ld;
    double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    const double* dinv = v.dinv + (long long)p * 4096;

    // group jb: columns [32 jb, 32 jb + 32) of the C rows, and row panel jb of L_kk
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
        for (int q = tid; q < TR_ROWS * 16; q += TR_THREADS) {  // 64 rows x 16 chunks
            int r = q >> 4, ch = jb * 16 + (q & 15);
            cp_async16(Xs + r * XS + ch * 2, Lp + (long long)(r0 + r) * ld + o + ch * 2);
        }
        const int nch = (jb + 1) * 16;  // chunks per row of the panel (lower blocks + diagonal block)
        for (int q = tid; q < 32 * nch; q += TR_THREADS) {
            int r = jb * 32 + q / nch, ch = q % nch;
            if (ch < jb * 16) cp_async16(Ls + r * XS + ch * 2, Lp + (long long)(o + r) * ld + o + ch * 2);
            else cp_async16(Ls + r * XS + ch * 2, dinv + jb * 1024 + (r & 31) * 32 + (ch & 15) * 2);
        }
        cp_async_commit();
    }
    if (tid < TB) zs[tid] = v.z[(long long)p * ld + o + tid];

    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, c4 = lane & 3;
    double* xrow = Xs + (warp * 8 + g) * XS;  // this lane's row (fragment row g of the warp's 8 rows)
    // forward-solve vector entry of row warp*8 + lane (lanes 0..7): fetched now, consumed at the end
    double* yp = v.y + (long long)p * ld;
    const double y_old = (lane < 8) ? yp[r0 + warp * 8 + lane] : 0.0;

#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
        if (jb == 0) cp_async_wait<3>();
        else if (jb == 1) cp_async_wait<2>();
        else if (jb == 2) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncthreads();
        // two accumulator sets (even / odd k of each LDS.128 pair): 8 independent DMMA chains
        double acc0[4][2], acc1[4][2];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) acc0[nb][0] = acc0[nb][1] = acc1[nb][0] = acc1[nb][1] = 0.0;
        // S = sum_{m<jb} X_m L[jb,m]^T     (k runs over columns [0, 32 jb); thread c4 takes k = kk+2c4, kk+2c4+1)
#pragma unroll 2
        for (int kk = 0; kk < jb * 32; kk += 8) {
            const double2 a = *reinterpret_cast<const double2*>(xrow + kk + 2 * c4);
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const double2 b = *reinterpret_cast<const double2*>(Ls + (jb * 32 + nb * 8 + g) * XS + kk + 2 * c4);
                dmma884(acc0[nb][0], acc0[nb][1], a.x, b.x);
                dmma884(acc1[nb][0], acc1[nb][1], a.y, b.y);
            }
        }
        // T = C_jb - S  (own rows only: warp-local dependency)
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
        // X_jb = T inv(L_jb,jb)^T
#pragma unroll
        for (int kk = 0; kk < 32; kk += 8) {
            const double2 a = *reinterpret_cast<const double2*>(xrow + jb * 32 + kk + 2 * c4);
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const double2 b = *reinterpret_cast<const double2*>(Ls + (jb * 32 + nb * 8 + g) * XS + jb * 32 + kk + 2 * c4);
                dmma884(acc0[nb][0], acc0[nb][1], a.x, b.x);
                dmma884(acc1[nb][0], acc1[nb][1], a.y, b.y);
            }
        }
        __syncwarp();
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
            *reinterpret_cast<double2*>(xrow + jb * 32 + nb * 8 + 2 * c4) =
                make_double2(acc0[nb][0] + acc1[nb][0], acc0[nb][1] + acc1[nb][1]);
        __syncwarp();
    }

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
            Lp[(long long)(r0 + r) * ld + o + c] = x;
            sacc = fma(x, zs[c], sacc);
        }
        sacc = warp_sum(sacc);
        if (lane == rr) dot_mine = sacc;
    }
    if (lane < 8) yp[r0 + warp * 8 + lane] = y_old - dot_mine;
}

// ------------------------------------------------------------------------------------------
// stand-alone Gram kernel: column-major K, both triangles; 64x64 tile pairs mirrored via smem
// ------------------------------------------------------------------------------------------
constexpr int GT = 64;
constexpr int GR_THREADS = 256;

__global__ void __launch_bounds__(GR_THREADS) agp_gram_kernel(const AgpInstr* __restrict__ prog_g, int m, int need, const double* __restrict__ ts,
                                                             int n, double noise, int form, double* __restrict__ K) {
    __shared__ double tile[GT][GT + 1];
    __shared__ double tsi[GT], tsj[GT];
    __shared__ AgpInstr prog_s[PROG_SMEM];
    // linear block id -> upper-triangular tile pair (bi <= bj)
    const int ntile = (n + GT - 1) / GT;
    int bj = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((long long)(bj + 1) * (bj + 2) / 2 <= (long long)blockIdx.x) ++bj;
    while ((long long)bj * (bj + 1) / 2 > (long long)blockIdx.x) --bj;
    const int bi = blockIdx.x - bj * (bj + 1) / 2;
    (void)ntile;
    const int tid = threadIdx.x;
    const int i0 = bi * GT, j0 = bj * GT;
    if (tid < GT) {
        tsi[tid] = (i0 + tid < n) ? ts[i0 + tid] : 0.0;
        tsj[tid] = (j0 + tid < n) ? ts[j0 + tid] : 0.0;
    }
    const AgpInstr* prog = prog_g;
    if (m <= PROG_SMEM) {
        const double* src = reinterpret_cast<const double*>(prog_g);
        double* dst = reinterpret_cast<double*>(prog_s);
        for (int q = tid; q < m * 4; q += GR_THREADS) dst[q] = src[q];
        prog = prog_s;
    }
    __syncthreads();
    const int il = tid & (GT - 1);
    const int gi = i0 + il;
#pragma unroll 1
    for (int e4 = 0; e4 < GT / 16; ++e4) {
        double t1[4], t2[4], val[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jl = (tid >> 6) + (e4 * 4 + j) * 4;
            // Symmetric(K): entry (i,j) takes the upper-triangle element (min,max)
            const bool up = gi <= j0 + jl;
            t1[j] = up ? tsi[il] : tsj[jl];
            t2[j] = up ? tsj[jl] : tsi[il];
        }
        eval_entries<4>(prog, m, need, t1, t2, form, val);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jl = (tid >> 6) + (e4 * 4 + j) * 4;
            const int gj = j0 + jl;
            double v = 0.0;
            if (gi < n && gj < n) {
                v = val[j];
                if (gi == gj) v = v + noise;
                K[(long long)gj * n + gi] = v;  // column j, rows contiguous
            }
            tile[jl][il] = v;
        }
    }
    if (bi != bj) {
        __syncthreads();
        // mirrored block: rows j0.., columns i0..  (element (gj, gi) = value(gi, gj))
        const int jl2 = tid & (GT - 1);
#pragma unroll 1
        for (int e = 0; e < GT / 4; ++e) {
            const int il2 = (tid >> 6) + e * 4;
            const int gi = i0 + il2, gj = j0 + jl2;
            if (gi < n && gj < n) K[(long long)gi * n + gj] = tile[jl2][il2];
        }
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
cudaError_t configure_kernels() {
    cudaError_t e;
    e = cudaFuncSetAttribute(agp_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UPD_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(agp_potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(agp_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_SMEM_BYTES);
    return e;
}

void launch_update(const BatchView& v, int P, int k, cudaStream_t s) {
    dim3 grid(2 * (v.nt - k), P);  // two 128x64 half tiles per tile
    agp_update_kernel<<<grid, UPD_THREADS, UPD_SMEM_BYTES, s>>>(v, k);
}

void launch_potf2(const BatchView& v, int P, int k, cudaStream_t s) {
    agp_potf2_kernel<<<P, PF_THREADS, PF_SMEM_BYTES, s>>>(v, k);
}

void launch_trsm(const BatchView& v, int P, int k, cudaStream_t s) {
    if (v.nt - k - 1 <= 0) return;
    dim3 grid(2 * (v.nt - k - 1), P);
    agp_trsm_kernel<<<grid, TR_THREADS, TR_SMEM_BYTES, s>>>(v, k);
}

void launch_gram(const AgpInstr* prog, int m, int need, const double* ts, int n, double noise, int form, double* K, cudaStream_t s) {
    if (n <= 0) return;
    int nt = (n + GT - 1) / GT;
    int blocks = nt * (nt + 1) / 2;
    agp_gram_kernel<<<blocks, GR_THREADS, 0, s>>>(prog, m, need, ts, n, noise, form, K);
}

}  // namespace agp
