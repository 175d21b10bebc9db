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
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "agp_eval.cuh"
#include "agp_kernels.cuh"
#include "agp_ptx.cuh"

// Switches that reproduce the round-1 stage race (profiles/r02_race_experiments.txt, tools/race_variants.sh); all 0 = product.
#ifndef AGP_X_SIMPLE
#define AGP_X_SIMPLE 0            // main loop without the `if (active)` blocks: ptxas then places the last LDS of a stage right before the release
#endif
#ifndef AGP_X_NO_RELEASE_FENCE
#define AGP_X_NO_RELEASE_FENCE 0  // drop the cross-proxy fence between a warp's reads of a stage and the stage's release (THE BUG: 1 bad run in 4)
#endif

namespace agp {

namespace {

constexpr int FT = 256;   // threads per CTA: 8 warps
constexpr int UM = 64;    // item rows
constexpr int UN = TB;    // item columns (one block column)
constexpr int KC = 16;    // K-chunk per pipeline stage (doubles) = one 128-byte row
constexpr int NSTAGE = 4;
constexpr int STAGE_D = (UM + UN) * KC;  // doubles per stage
constexpr int XS = 136;                  // X / L_kk-panel row stride: 8 mod 16 doubles -> conflict-free LDS.128
constexpr int BS = 33;                   // potf2 32x32 block row stride (odd: lane-per-row walks conflict free)
constexpr int BLK = 32 * BS;
constexpr int REGION_D = UM * XS + 32 * XS;  // X rows + one 32-row panel of L_kk
constexpr int PROG_SMEM = 64;
// tail of the shared-memory image (doubles): zs[TB] ys[TB] Ri[TB] red[16]
constexpr int TAIL_D = 3 * TB + 16;
constexpr int FUSED_SMEM = (REGION_D + TAIL_D) * 8 + 64 + 64;  // + ctl[16] ints + full[4], empty[4] mbarriers

static_assert(NSTAGE * STAGE_D <= REGION_D, "pipeline stages must fit in the region");
static_assert(10 * BLK <= REGION_D, "packed diagonal tile must fit in the region");
static_assert(UM * XS + 4 * 32 * 32 <= REGION_D, "X rows + the solve's ring of four 32x32 operand blocks");
static_assert(2 * (FUSED_SMEM + 1024) <= 228 * 1024, "two CTAs per SM");

// double offset of 16-byte chunk `chunk` (0..7) of row `row` in a [rows][128 B] tile written by TMA with SWIZZLE_128B
__device__ __forceinline__ int swz128(int row, int chunk) { return row * KC + ((chunk ^ (row & 7)) << 1); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}

// thread 0 only: wait until *flag >= need.  A wait that exceeds the limit (2 s by default) raises the scheduler error
// flag so every CTA drains instead of hanging the device.
__device__ bool wait_ge(const int* flag, int need, int* err, unsigned long long limit_ns) {
    if (ld_acquire_gpu(flag) >= need) return true;
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    for (;;) {
        __nanosleep(64);
        if (ld_acquire_gpu(flag) >= need) return true;
        if ((++spins & 255u) == 0) {
            if (ld_relaxed_gpu(err) != 0) return false;
            if (globaltimer_ns() - t0 > limit_ns) {
                atomicExch(err, 1);
                return false;
            }
        }
    }
}

// mbarrier wait that cannot hang the device: gives up (and raises the scheduler error flag) after the same limit
// as the dependency waits
__device__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int* err, unsigned long long limit_ns) {
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
        if ((spins & 1023u) == 1023u) {
            if (ld_relaxed_gpu(err) != 0) return false;
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > limit_ns) {
                atomicExch(err, 1);
                return false;
            }
        }
    }
}

// diagnostics: thread 0 stamps phase boundaries of item `idx` when tracing is on
__device__ __forceinline__ void stamp(const SchedView& q, int idx, int slot) {
    if (q.trace != nullptr && threadIdx.x == 0) q.trace[(long long)idx * 8 + slot] = (long long)globaltimer_ns();
}

// all threads: release this item's global writes, then bump the counter
__device__ __forceinline__ void signal_done(int* counter) {
    fence_proxy_async();  // this item's shared-memory traffic is ordered before the next item's TMA copies into the same buffers
    fence_proxy_async_global();  // this thread's generic-proxy stores to L are ordered before the TMA (async-proxy) reads of the CTAs the counter releases
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1);
    }
}

// Shared-memory image of one CTA.  Every device function rebuilds this view from the extern
// array itself (never through a pointer argument) so the compiler keeps the shared address
// space and emits LDS/STS instead of generic loads.
struct Smem {
    double* region;
    double* zs;
    double* ys;
    double* Ri;
    double* red;
    int* ctl;  // [0] item index, [1] wait result, [2] potf2 info, [4] pipeline chunks issued so far by this CTA (mbarrier phases)
    uint64_t* full;   // [NSTAGE] stage filled (TMA transaction bytes)
    uint64_t* empty;  // [NSTAGE] stage read by all 8 warps
};

__device__ __forceinline__ Smem smem_view() {
    extern __shared__ __align__(1024) unsigned char smem_raw[];  // TMA destinations with SWIZZLE_128B need 1 KB alignment
    Smem s;
    s.region = reinterpret_cast<double*>(smem_raw);
    s.zs = s.region + REGION_D;
    s.ys = s.zs + TB;
    s.Ri = s.ys + TB;
    s.red = s.Ri + TB;
    s.ctl = reinterpret_cast<int*>(s.red + 16);
    s.full = reinterpret_cast<uint64_t*>(s.ctl + 16);
    s.empty = s.full + NSTAGE;
    return s;
}

// ------------------------------------------------------------------------------------------
// ITEM_DIAG / ITEM_PANEL
// ------------------------------------------------------------------------------------------
// Contraction range [j0, j1) in block columns.  A finishing item (j1 == k) hands the diagonal tile to
// potf2 or takes a panel through the triangular solve.  A PARTIAL item only stores: the tile in L
// receives K - sum_{j<j1}; either a later item with j0 = j1 picks it up from there (the accumulators
// always start from minus the tile), which takes the long early part of the contraction of the
// next diagonal tile and of the panel below it off the per-particle critical path.
__device__ __noinline__ bool do_update(const BatchView& v, const SchedView& q, const TmaMaps& maps, int idx, int p, int k, int i, int h, bool diag, bool partial,
                                       bool yinit, int j0, int j1, int need_k, int need_i, int extra_flag, int extra_need) {
    const Smem s = smem_view();
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
    // warp tiles strictly above the diagonal of a diagonal tile are never read
    const bool active = !diag || (wn * 32 <= h * UM + wm * 32 + 31);
    // The Gram tile K(ts_i, ts_k) [+ noise I] was written into L by agp_gramfill_kernel.  Block
    // column 0 needs no contraction: the diagonal tile is already in place and a panel goes
    // straight to shared memory.  Otherwise the accumulators start from -K, so that after the
    // contraction  acc = -(K - sum_j L_ij L_kj^T).
    if (k == 0 && diag) {
        if (tid < UM) v.y[(long long)p * ld + row0 + tid] = (row0 + tid < v.n) ? v.xs[row0 + tid] : 0.0;
        signal_done(q.diagu + p * q.nt_stride + k);
        return true;
    }
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
            const double2 kv = __ldcg(reinterpret_cast<const double2*>(Lp + (long long)(row0 + r) * ld + col0 + c));
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
            if (ks == 1) {
                // Release of the stage.  The LDS above are generic-proxy reads, the next use of the stage is written by the
                // async proxy (TMA): every lane orders its own reads before the release with a cross-proxy fence, exactly as
                // CUTLASS does before consumer_release when a TMA-fed buffer is read with ordinary loads.  Without it the
                // TMA box of chunk ch + NSTAGE can land while a late LDS of chunk ch is still queued behind the co-resident
                // CTA's shared-memory traffic (round 1's "one wrong row / 8x8 block in 1 of 10^4 items").
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
                    for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].x, b[nb].x);
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].y, b[nb].y);
            }
        }
    }
    if (tid == 0) s.ctl[4] = G0 + nchunk;
    __syncthreads();
    stamp(q, idx, 2);

    // --- X = -acc:  the diagonal tile goes back to L (lower part), a panel stays in shared memory ----
    double* Xs = s.region;            // [UM][XS]
    double* Ls = s.region + UM * XS;  // [32][XS]
    if (active) {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const int r = wm * 32 + mb * 8 + g, c = wn * 32 + nb * 8 + 2 * c4;
                const double2 x2 = make_double2(-acc[mb][nb][0], -acc[mb][nb][1]);
                if (!diag && !partial) {
                    *reinterpret_cast<double2*>(Xs + r * XS + c) = x2;
                } else {
                    const int rd = diag ? h * UM + r : TB;  // row inside a diagonal tile: only c <= rd is kept
                    double* dst = Lp + (long long)(row0 + r) * ld + col0 + c;
                    if (c + 1 <= rd) *reinterpret_cast<double2*>(dst) = x2;
                    else if (c <= rd) dst[0] = x2.x;
                }
            }
    }
    if (diag) {
        signal_done(q.diagu + p * q.nt_stride + k);
        return true;
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

__device__ __noinline__ bool do_potf2(const BatchView& v, const SchedView& q, int idx, int p, int k, int need_diag) {
    const Smem s = smem_view();
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

    // lower triangle -> packed blocks (L2 loads: the tile was written by other CTAs of this launch)
#pragma unroll 1
    for (int base = 0; base < TB * TB; base += FT * 8) {
        double tmp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int idx = base + u * FT + tid;
            int r = idx >> 7, c = idx & (TB - 1);
            tmp[u] = (c <= r) ? __ldcg(Lp + (long long)(o + r) * ld + o + c) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int idx = base + u * FT + tid;
            int r = idx >> 7, c = idx & (TB - 1);
            if ((c >> 5) <= (r >> 5)) Ab[blk_off(r >> 5, c >> 5) + (r & 31) * BS + (c & 31)] = tmp[u];
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
            named_bar_sync(1, FT);  // diagonal block jb is final
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
                double* roww = Dg + lane * BS;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (c <= lane) roww[c] = a[c];
                if (lane == 0 && bad != 0 && s.ctl[2] == 0) s.ctl[2] = bad;
            }
            named_bar_sync(1, FT);
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
            named_bar_sync(3, WORKERS);
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

}  // namespace

__global__ void __launch_bounds__(FT, 2) agp_chol_kernel(BatchView v, SchedView q, const __grid_constant__ TmaMaps maps) {
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
        const int4 it = __ldg(q.items + 2 * idx), dep = __ldg(q.items + 2 * idx + 1);
        const int type = it.x & 0xff, h = (it.x >> 8) & 1;
        bool ok;
        stamp(q, idx, 0);
        if (type == ITEM_POTF2) ok = do_potf2(v, q, idx, it.y, it.z, dep.w);
        else ok = do_update(v, q, maps, idx, it.y, it.z, it.w, h, type == ITEM_DIAG, (it.x & ITEM_PARTIAL) != 0, (it.x & ITEM_YINIT) != 0,
                            dep.x & 0xffff, dep.x >> 16, dep.y & 0xffff, dep.y >> 16, dep.z, dep.w);
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

// ------------------------------------------------------------------------------------------
// Gram fill: L tile (i,k), i >= k  <-  K(ts_i, ts_k) [+ noise I]; identity in the padding rows.
// grid (2 * nt(nt+1)/2, P): one CTA = 64 rows x 128 columns; thread -> column, 32 rows, E
// entries per interpreter pass; rows are written as coalesced 1 KB segments.
// ------------------------------------------------------------------------------------------
#ifndef AGP_GF_E
#define AGP_GF_E 4
#endif
#ifndef AGP_GRAD_E
#define AGP_GRAD_E 2
#endif
#ifndef AGP_GF_MINB
#define AGP_GF_MINB 2
#endif
constexpr int GF_E = AGP_GF_E, GF_MINB = AGP_GF_MINB;

// LONGPROG = false: every program of the batch fits the shared-memory cache (the interpreter then reads
// instructions with LDS instead of generic loads); the host picks the variant per batch.
template <int E, int MINB, bool LONGPROG>
__global__ void __launch_bounds__(FT, MINB) agp_gramfill_kernel(BatchView v, int tile_id0) {
    __shared__ __align__(128) double ts_r[UM];
    __shared__ __align__(128) double ts_c[UN];
    __shared__ AgpInstr prog_s[PROG_SMEM];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int p = blockIdx.y;
    // linear id -> lower-triangular tile (i >= k) and row half
    const int t = tile_id0 + (blockIdx.x >> 1), h = blockIdx.x & 1;
    int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    const int k = t - i * (i + 1) / 2;
    const bool diag = (i == k);
    const int row0 = i * TB + h * UM, col0 = k * TB;
    const int ld = v.ld;
    double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (UM + UN) * 8);
        tma_bulk_g2s(ts_r, v.ts + row0, UM * 8, &bar);
        tma_bulk_g2s(ts_c, v.ts + col0, UN * 8, &bar);
    }
    const int poff = v.prog_off[p];
    const int pm = v.prog_off[p + 1] - poff;
    if (!LONGPROG) {
        const double* src = reinterpret_cast<const double*>(v.prog + poff);
        double* dst = reinterpret_cast<double*>(prog_s);
        for (int w = tid; w < pm * AGP_INSTR_DOUBLES; w += FT) dst[w] = src[w];
    }
    mbar_wait(&bar, 0);
    __syncthreads();

    const int need = v.prog_need[p];
    const double noise = v.noise[p];
    const int n = v.n;
    const int c = tid & (UN - 1), rbase = tid >> 7;
    const int gc = col0 + c;
    const double tcol = ts_c[c];
    const bool plain = !diag && row0 + UM <= n && col0 + UN <= n;
    const bool no_kernel_rows = v.aug_identity && row0 >= v.nt * TB;
#pragma unroll 1
    for (int eb = 0; eb < 32 / E; ++eb) {
        double t1[E], t2[E], val[E];
        const int rlast = rbase + 2 * (E * eb + E - 1);
        // nothing to evaluate: the strictly-upper part of a diagonal tile (zeros), and every appended row of an
        // identity-augmented batch ([I 0]: agp_lml_grad_batch — three quarters of that matrix)
        const bool skip = (diag && c > rlast + h * UM) || no_kernel_rows;
        if (!skip) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                t1[j] = tcol;  // upper-triangle element (gc, gr): gc <= gr
                t2[j] = ts_r[rbase + 2 * (E * eb + j)];
            }
            if (LONGPROG) eval_entries<E>(v.prog + poff, pm, need, t1, t2, 0, val);
            else eval_entries<E>(prog_s, pm, need, t1, t2, 0, val);
        }
        if (plain) {
            // interior tile (all rows and columns are observations, no diagonal): nothing to decide per entry
            double* dst = Lp + (long long)(row0 + rbase + 2 * E * eb) * ld + gc;
#pragma unroll
            for (int j = 0; j < E; ++j) dst[(long long)(2 * j) * ld] = val[j];
            continue;
        }
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int r = rbase + 2 * (E * eb + j);
            const int gr = row0 + r;
            // rows/columns: [0, n) observations | [n, nt*TB) padding | [nt*TB, nt*TB + n_pred) appended
            // prediction points | padding.  Padding rows are independent unit-variance dummies
            // (identity block: log 1 = 0 in the log det, 0 in the quadratic form).
            double out = 0.0;
            if (!(diag && c > r + h * UM)) {
                const int lt = v.nt * TB;
                if (gr < n) {  // gc <= gr < n
                    out = val[j];
                    if (gr == gc) out = out + noise;  // + noise*I, src/GP.jl:667
                } else if (v.aug_identity && gr >= lt) {
                    out = (gr - lt == gc) ? 1.0 : 0.0;  // [I 0]: the appended rows solve to L^{-T}, -K^{-1}, -alpha
                } else if (gr >= lt && gr < lt + v.n_pred && (gc < n || gc >= lt)) {
                    out = val[j];  // K(t, t*) and K(t*, t*), no noise (src/GP.jl:743-747)
                } else {
                    out = (gr == gc) ? 1.0 : 0.0;
                }
            }
            Lp[(long long)gr * ld + gc] = out;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Predictive distribution out of the augmented factorisation (src/GP.jl:731-758): with the m
// prediction points appended as extra rows, the panel solves leave L_21 = K_21 L_11^{-T}, the
// forward solve leaves y_2 = 0 - L_21 z = -K_21 K_11^{-1} xs, and the partial items leave the Schur
// complement K_22 - L_21 L_21^T in the trailing lower tiles.
//   mean[p][a]   = -y[nt*TB + a]
//   cov[p][a][b] = S[max(a,b)][min(a,b)] + noise_pred[p] * (a == b)      (symmetric by construction)
// ------------------------------------------------------------------------------------------
__global__ void agp_predict_extract_kernel(BatchView v, const double* __restrict__ noise_pred, double* __restrict__ mean_out,
                                           double* __restrict__ cov_out) {
    const int p = blockIdx.y;
    const int m = v.n_pred, o = v.nt * TB, ld = v.ld;
    const double* Lp = v.L + (long long)p * v.mat_stride;
    const double np_ = noise_pred[p];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)m * m; e += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(e / m), b = (int)(e % m);
        const int hi = a > b ? a : b, lo = a > b ? b : a;
        double s = Lp[(long long)(o + hi) * ld + o + lo];
        if (a == b) s = s + np_;
        cov_out[(long long)p * m * m + e] = s;
    }
    if (blockIdx.x == 0)
        for (int a = threadIdx.x; a < m; a += blockDim.x) mean_out[(long long)p * m + a] = -v.y[(long long)p * ld + o + a];
}

void launch_predict_extract(const BatchView& v, int P, const double* noise_pred, double* mean_out, double* cov_out, cudaStream_t s) {
    if (P <= 0 || v.n_pred <= 0) return;
    long long e = (long long)v.n_pred * v.n_pred;
    int bx = (int)((e + 255) / 256);
    if (bx > 1024) bx = 1024;
    agp_predict_extract_kernel<<<dim3(bx, P), 256, 0, s>>>(v, noise_pred, mean_out, cov_out);
}

// ------------------------------------------------------------------------------------------
// Gradient of the log marginal likelihood (Gen mvnormal logpdf_grad + ReverseDiff
// through eval_cov in the reference):   dLML/dtheta = 1/2 sum_ik A_ik dK_ik/dtheta,
// A = alpha alpha^T - K^{-1}.  After the identity-augmented factorisation the trailing lower tiles
// hold -K^{-1} and the appended forward-solve entries hold -alpha, so
//   A_ik = y[lt+i] y[lt+k] + L[lt+i][lt+k].
// One CTA = 64 rows x 128 columns of a lower tile; every thread walks its entries through the
// reverse-mode interpreter and keeps per-parameter sums; the CTA reduces them in a fixed order and
// a second kernel adds the per-CTA partials in a fixed order (bitwise reproducible, no atomics).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FT) agp_grad_kernel(BatchView v, const int* __restrict__ param_off, double* __restrict__ partial) {
    __shared__ AgpInstr prog_s[PROG_SMEM];
    __shared__ unsigned char opa_s[PROG_SMEM], opb_s[PROG_SMEM];
    __shared__ double red[FT / 32][AGP_GRAD_MAX_PARAMS + 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = blockIdx.y;
    const int t = blockIdx.x >> 1, h = blockIdx.x & 1;
    int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    const int k = t - i * (i + 1) / 2;
    const int row0 = i * TB + h * UM, col0 = k * TB;
    const int ld = v.ld, n = v.n, lt = v.nt * TB;
    const double* __restrict__ Lp = v.L + (long long)p * v.mat_stride;
    const double* __restrict__ nal = v.y + (long long)p * ld + lt;  // -alpha
    const int poff = v.prog_off[p];
    const int pm = v.prog_off[p + 1] - poff;
    {
        const double* src = reinterpret_cast<const double*>(v.prog + poff);
        double* dst = reinterpret_cast<double*>(prog_s);
        for (int w = tid; w < pm * AGP_INSTR_DOUBLES; w += FT) dst[w] = src[w];  // host guarantees pm <= PROG_SMEM
    }
    __syncthreads();
    if (tid == 0) grad_operands(prog_s, pm, opa_s, opb_s);
    __syncthreads();
    const int np = param_off[p + 1] - param_off[p];  // host guarantees np <= AGP_GRAD_MAX_PARAMS
    double g[AGP_GRAD_MAX_PARAMS + 1];
    for (int j = 0; j <= np; ++j) g[j] = 0.0;
    const int c = tid & (UN - 1), rbase = tid >> 7;
    const int gc = col0 + c;
    if (gc < n) {
        const double tcol = v.ts[gc];
        const double nac = nal[gc];
        constexpr int GE_ = AGP_GRAD_E;  // entries per interpreter pass
#pragma unroll 1
        for (int e0 = 0; e0 < 32; e0 += GE_) {
            double t1[GE_], t2[GE_], wgt[GE_];
            bool any = false;
#pragma unroll
            for (int u = 0; u < GE_; ++u) {
                const int gr = row0 + rbase + 2 * (e0 + u);
                const bool ok = gr < n && gc <= gr;
                t1[u] = tcol;
                t2[u] = ok ? v.ts[gr] : tcol;  // an entry outside the lower triangle runs as a (k(t,t), weight 0) dummy
                double A = 0.0;
                if (ok) {
                    A = nal[gr] * nac + Lp[(long long)(lt + gr) * ld + lt + gc];
                    if (gr == gc) g[np] += A;  // dK/dnoise = I
                }
                wgt[u] = ok ? ((gr == gc) ? A : 2.0 * A) : 0.0;  // off-diagonal entries count twice (symmetry)
                any = any || ok;
            }
            if (!any) continue;
            eval_entries_grad<GE_>(prog_s, pm, opa_s, opb_s, t1, t2, wgt, [&](int j, double d) { g[j] += d; });
        }
    }
    for (int j = 0; j <= np; ++j) {
        const double sres = warp_sum(g[j]);
        if (lane == 0) red[warp][j] = sres;
    }
    __syncthreads();
    if (tid <= np) {
        double sres = 0.0;
#pragma unroll
        for (int w = 0; w < FT / 32; ++w) sres += red[w][tid];
        partial[((long long)p * gridDim.x + blockIdx.x) * (AGP_GRAD_MAX_PARAMS + 1) + tid] = sres;
    }
}

__global__ void agp_grad_reduce_kernel(const double* __restrict__ partial, int blocks, const int* __restrict__ param_off, double* __restrict__ grad_out,
                                       double* __restrict__ gnoise_out) {
    const int p = blockIdx.x, j = threadIdx.x;
    const int np = param_off[p + 1] - param_off[p];
    if (j > np) return;
    double sres = 0.0;
    for (int b = 0; b < blocks; ++b) sres += partial[((long long)p * blocks + b) * (AGP_GRAD_MAX_PARAMS + 1) + j];
    if (j < np) grad_out[param_off[p] + j] = 0.5 * sres;
    else gnoise_out[p] = 0.5 * sres;
}

int grad_blocks_per_particle(const BatchView& v) { return v.nt * (v.nt + 1); }

void launch_grad(const BatchView& v, int P, const int* param_off, double* partial, double* grad_out, double* gnoise_out, cudaStream_t s) {
    if (P <= 0 || v.nt <= 0) return;
    const int blocks = grad_blocks_per_particle(v);
    agp_grad_kernel<<<dim3(blocks, P), FT, 0, s>>>(v, param_off, partial);
    agp_grad_reduce_kernel<<<P, AGP_GRAD_MAX_PARAMS + 1, 0, s>>>(partial, blocks, param_off, grad_out, gnoise_out);
}

cudaError_t configure_fused() {
    return cudaFuncSetAttribute(agp_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
}

void launch_gramfill(const BatchView& v, int P, int row_tile0, cudaStream_t s) {
    if (P <= 0 || v.nt_total <= row_tile0) return;
    const int tile_id0 = row_tile0 * (row_tile0 + 1) / 2;  // lower tiles are numbered row by row
    const int tiles = v.nt_total * (v.nt_total + 1) / 2 - tile_id0;
    dim3 grid(2 * tiles, P);  // 2 row halves per tile
    // four entries per interpreter pass in lock-step (agp_math.cuh), two CTAs per SM
    if (v.max_prog_len <= PROG_SMEM) agp_gramfill_kernel<GF_E, GF_MINB, false><<<grid, FT, 0, s>>>(v, tile_id0);
    else agp_gramfill_kernel<GF_E, GF_MINB, true><<<grid, FT, 0, s>>>(v, tile_id0);
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
