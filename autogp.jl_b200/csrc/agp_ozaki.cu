// Error-free int8 split of the long contractions of the blocked Cholesky onto tcgen05.mma kind::i8 (sm_100a).
// Interface and the arithmetic of the scheme: agp_ozaki.cuh.  Three kernels:
//
//   agp_ozaki_rowscale_kernel   per row r of every particle: e_r = ceil(log2 sqrt(K_rr)) from the Gram diagonal (appended rows of
//                               the gradient calls: from the noise)
//   agp_ozaki_slice_kernel      finished tiles of L -> seven int8 digit planes
//   agp_ozaki_update2_kernel    T_ik -= sum_j L_ij L_kj^T over the block columns left of a super-column, for every lower tile of it:
//                               warp-specialised (TMA producer, one MMA-issuing warp, eight epilogue warps), 128-column int32
//                               accumulators in TMEM, two passes over the weight groups, per CTA (<1>) or per CTA pair (<2>,
//                               tcgen05.mma.cta_group::2) — described at the kernel
//
// Reference semantics: the products are part of dpotrf as called by PDMats for `mvnormal` (src/Model.jl:136).
#include "agp_ozaki.cuh"

#include <math.h>

#include "agp_ozaki_digits.cuh"
#include "agp_ptx.cuh"

namespace agp {

namespace {

// K-major operand tile, rows of 128 bytes, SWIZZLE_128B: 8-row groups of 1024 bytes (stride byte offset), version 1
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ unsigned long long oz_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}
// mbarrier wait that cannot hang the device (same policy as the persistent kernel's waits)
__device__ __forceinline__ bool oz_wait(uint64_t* bar, uint32_t parity, int* err, unsigned long long limit_ns) {
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
            const unsigned long long now = oz_timer();
            if (t0 == 0) t0 = now;
            else if (now - t0 > limit_ns) {
                atomicExch(err, 2);
                return false;
            }
        }
    }
}

#ifndef OZ_STATS
#define OZ_STATS 0
#endif
#if OZ_STATS
__device__ long long oz_stats[8];  // MMA thread: total, wait B, wait A, wait TMEM free; epilogue warp 2: total busy; items
#define OZ_T0() const long long t_s_ = clock64()
#define OZ_ACC(slot) oz_acc_[slot] += clock64() - t_s_
#else
#define OZ_T0()
#define OZ_ACC(slot)
#endif

// ---- the update kernel: 128-column accumulators, two passes over the weight groups, optional CTA pairs -------------
// An M128 instruction re-reads its 4 KB A operand from shared memory whatever N is: N = 64 costs 49.5 clocks instead of 32
// (6 KB at 128 B/clk, profiles/r02_i8_shape_probe.txt; the first generation of this kernel — N = 64, one accumulator per
// weight group, one pass — was bound by that at 65 % of the int8 rate).  Here an accumulator is 128 columns wide, so FOUR
// weight groups fit into TMEM and a tile takes two passes over the contraction: groups 0..3 (10 digit-plane products per
// 128-deep chunk: planes 0..3 of either operand), then groups 4..6 (18 products, all seven planes).  With G = 2 two CTAs of a
// cluster (an SM pair) work on the tiles (i, k) and (i + 1, k): tcgen05.mma.cta_group::2 (M = 256) reads each CTA's own A rows
// and HALF of the 128 B rows from either CTA's shared memory — 6 KB per 64 clocks and SM instead of 8 — issued by the leader
// CTA for both; the operands arrive by cta_group::2 TMA loads that signal the leader's barriers, stages are released in both
// CTAs by multicast commits.
template <int G>
struct Oz2 {
    static constexpr int NA = (G == 2) ? 8 : 6;    // A ring: one digit plane of 128 rows x 128 bytes per slot
    static constexpr int NB = (G == 2) ? 10 : 7;   // B ring: one digit plane of 128 / G rows per slot
    static constexpr uint32_t A_BYTES = 128 * 128;
    static constexpr uint32_t B_BYTES = 128 * 128 / G;
    static constexpr uint32_t OFF_B = NA * A_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_B + NB * B_BYTES;
    static constexpr int SMEM = (int)OFF_BAR + 512;
    static constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | (((128u * G) >> 4) << 24);
};
static_assert(Oz2<1>::SMEM <= 227 * 1024 && Oz2<2>::SMEM <= 227 * 1024, "one CTA per SM");

template <int G>
__device__ __forceinline__ void oz2_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    if constexpr (G == 1) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(da), "l"(db), "r"(Oz2<1>::IDESC), "r"(accumulate), "r"(0u)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(da), "l"(db), "r"(Oz2<2>::IDESC), "r"(accumulate), "r"(0u)
            : "memory");
    }
}
template <int G>
__device__ __forceinline__ void oz2_commit(uint64_t* bar) {
    if constexpr (G == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
    } else {
        // arrives on the barrier at this offset in BOTH CTAs of the pair
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
    }
}
// TMA tensor load whose completion bytes go to the LEADER CTA's barrier (cta_group::2); leader_bar is a shared::cluster address
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tensor_map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(tensor_map), "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// one lane of a converged warp (the MMA warp runs its control flow on all 32 lanes so that descriptors and barrier
// addresses stay warp-uniform — uniform registers, no per-thread election loop around every tcgen05 instruction)
__device__ __forceinline__ uint32_t oz_elect_one() {
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n"
        ".reg .b32 %%rx;\n"
        ".reg .pred %%px;\n"
        "     elect.sync %%rx|%%px, %2;\n"
        "@%%px mov.s32 %1, 1;\n"
        "     mov.s32 %0, %%rx;\n"
        "}\n"
        : "+r"(laneid), "+r"(pred)
        : "r"(0xFFFFFFFF));
    return pred;
}

// units of one launch (a unit = one tile, or with G = 2 the tiles (i0, k) and (i0 + 1, k) of a CTA pair): particle-major,
// then tile rows (pairs aligned at c0), the block columns of the super-column innermost — the units that run at the same
// time then read the same A rows, and the B rows of the few block columns stay in L2.  A tile above the diagonal
// (i0 < k <= i0 + 1) or below the matrix is computed and not stored.
__device__ __forceinline__ int oz2_fc(int r, int nt) { return r >= nt ? r - nt : 0; }
template <int G>
__device__ __forceinline__ int oz2_kcount(int i0, const OzakiParams& prm) {
    int kc = i0 + G - prm.k_lo;  // tile rows k_lo .. i0 + G - 1 have a tile on or below the diagonal in this row group
    kc = kc < prm.k_hi - prm.k_lo ? kc : prm.k_hi - prm.k_lo;
    return kc > 0 ? kc : 0;
}
template <int G>
__device__ __forceinline__ int oz2_units_per_particle(const OzakiParams& prm) {
    int n = 0;
    for (int i0 = prm.r_lo; i0 < prm.r_hi; i0 += G) n += oz2_kcount<G>(i0, prm);
    return n;
}
struct Oz2Unit {
    int p, i, k, clo;
    bool valid;
};
template <int G>
__device__ __forceinline__ Oz2Unit oz2_decode(int unit, int per_p, const OzakiParams& prm, int rank) {
    Oz2Unit it;
    it.p = unit / per_p;
    int r = unit - it.p * per_p;
    int i0 = prm.r_lo;
    for (;;) {
        const int kc = oz2_kcount<G>(i0, prm);
        if (r < kc) break;
        r -= kc;
        i0 += G;
    }
    it.k = prm.k_lo + r;
    it.i = i0 + rank;
    it.valid = it.i < prm.r_hi && it.i >= it.k;
    if (it.i >= prm.r_hi) it.i = prm.r_hi - 1;  // stay inside the matrix; nothing is stored
    const int fi = oz2_fc(i0, prm.nt), fk = oz2_fc(it.k, prm.nt);
    it.clo = fi > fk ? fi : fk;  // the same for both CTAs of a pair: the leader issues for both
    return it;
}

// the MMA instructions of one 128-deep chunk of one pass; PASS 0: weight groups 0..3, PASS 1: groups 4..6
template <int G, int PASS>
__device__ __forceinline__ bool oz2_chunk(uint64_t* a_full, uint64_t* a_empty, uint64_t* b_full, uint64_t* b_empty, uint64_t da0, uint64_t db0,
                                          uint32_t tmem, uint32_t first, int& a_slot, int& a_par, int& b_slot, int& b_par, const OzakiParams& prm
#if OZ_STATS
                                          , long long* oz_acc_
#endif
                                          ) {
    constexpr int NA = Oz2<G>::NA, NB = Oz2<G>::NB;
    constexpr int PMAX = PASS ? OZ_SLICES - 1 : 3, GMIN = PASS ? 4 : 0;
    static_assert(OZ_SLICES == 7, "pass structure: groups 0..3 | 4..6");
    // ring positions of this chunk's B planes: plane q sits in slot (b_slot + q) mod NB
    uint64_t dbq[PMAX + 1];
    int bs[PMAX + 1];
#pragma unroll
    for (int q = 0; q <= PMAX; ++q) {
        int s = b_slot + q;
        if (s >= NB) s -= NB;
        bs[q] = s;
        dbq[q] = db0 + (uint64_t)(s * (int)(Oz2<G>::B_BYTES >> 4));
    }
#pragma unroll
    for (int p = PMAX; p >= 0; --p) {
        const int qn = PMAX - p;  // the plane of B that is new at this step
        {
            const int par = (b_slot + qn >= NB) ? (b_par ^ 1) : b_par;
            OZ_T0();
            if (!oz_wait(b_full + bs[qn], par, prm.err, prm.wait_timeout_ns)) return false;
            OZ_ACC(1);
        }
        {
            OZ_T0();
            if (!oz_wait(a_full + a_slot, a_par, prm.err, prm.wait_timeout_ns)) return false;
            OZ_ACC(2);
        }
        tc_fence_after();
        const uint64_t da = da0 + (uint64_t)(a_slot * (int)(Oz2<G>::A_BYTES >> 4));
        if (oz_elect_one()) {
#pragma unroll
            for (int q = (GMIN - p > 0 ? GMIN - p : 0); q <= PMAX - p; ++q) {
                const uint32_t d = tmem + (uint32_t)((p + q) & 3) * 128u;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) oz2_mma<G>(d, da + 2 * k4, dbq[q] + 2 * k4, (q == 0 && k4 == 0) ? first : 1u);
            }
            oz2_commit<G>(a_empty + a_slot);
            // planes of B whose last product was in this step
            if (PASS == 1 && p >= 1 && p <= 4) oz2_commit<G>(b_empty + bs[4 - p]);
            if (p == 0) {
#pragma unroll
                for (int q = (PASS ? 4 : 0); q <= PMAX; ++q) oz2_commit<G>(b_empty + bs[q]);
            }
        }
        __syncwarp();
        if (++a_slot == NA) a_slot = 0, a_par ^= 1;
    }
    b_slot += PMAX + 1;
#pragma unroll
    for (int w = 0; w < 2; ++w)  // PMAX + 1 may exceed the ring: up to two trips
        if (b_slot >= NB) b_slot -= NB, b_par ^= 1;
    return true;
}

}  // namespace

constexpr int OZ2_THREADS = 320;  // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue
template <int G>
__global__ void __launch_bounds__(OZ2_THREADS, 1) agp_ozaki_update2_kernel(const __grid_constant__ OzakiParams prm, const __grid_constant__ OzakiMaps maps) {
    using Lay = Oz2<G>;
    constexpr int NA = Lay::NA, NB = Lay::NB;
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    unsigned char* As = oz_smem;
    unsigned char* Bs = oz_smem + Lay::OFF_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(oz_smem + Lay::OFF_BAR);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + NA;
    uint64_t* b_full = a_empty + NA;
    uint64_t* b_empty = b_full + NB;
    uint64_t* acc_full = b_empty + NB;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P = prm.P, ld = prm.ld;
    uint32_t rank = 0;
    if constexpr (G == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
    const int per_p = oz2_units_per_particle<G>(prm);
    const int n_units = per_p * P;
    const int unit0 = blockIdx.x / G, unit_stride = gridDim.x / G;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&maps.a) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&maps.b) : "memory");
        for (int s = 0; s < NA; ++s) {
            mbar_init(a_full + s, 1);
            mbar_init(a_empty + s, 1);
        }
        for (int s = 0; s < NB; ++s) {
            mbar_init(b_full + s, 1);
            mbar_init(b_empty + s, 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8 * G);
        mbar_fence_init();
    }
    if (warp == 0) {
        if constexpr (G == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
        }
    }
    tc_fence_before();
    if constexpr (G == 2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- producer (either CTA loads its own A rows and its share of the B rows) ------------------------
            int a_slot = 0, a_n = 0, b_slot = 0, b_n = 0;  // ring position, completed trips round the ring
            bool ok = true;
            for (int unit = unit0; unit < n_units && ok; unit += unit_stride) {
                const Oz2Unit it = oz2_decode<G>(unit, per_p, prm, (int)rank);
                const int arow = it.p * ld + it.i * 128, brow = it.p * ld + it.k * 128 + (int)rank * (128 / G);
                for (int pass = 0; pass < 2 && ok; ++pass) {
                    const int pmax = pass ? OZ_SLICES - 1 : 3;
                    for (int c = it.clo; c < prm.chi && ok; ++c) {
                        for (int p = pmax; p >= 0; --p) {
                            const int q = pmax - p;
                            if (b_n >= 1) ok = oz_wait(b_empty + b_slot, (b_n - 1) & 1, prm.err, prm.wait_timeout_ns);
                            if (!ok) break;
                            if constexpr (G == 1) {
                                mbar_expect_tx(b_full + b_slot, Lay::B_BYTES);
                                tma_load_2d(Bs + b_slot * Lay::B_BYTES, &maps.a, c * 128, q * P * ld + brow, b_full + b_slot);
                            } else {
                                if (rank == 0) mbar_expect_tx(b_full + b_slot, 2 * Lay::B_BYTES);
                                tma_load_2d_pair(Bs + b_slot * Lay::B_BYTES, &maps.b, c * 128, q * P * ld + brow, map_to_cta(smem_u32(b_full + b_slot), 0));
                            }
                            if (++b_slot == NB) b_slot = 0, ++b_n;
                            if (a_n >= 1) ok = oz_wait(a_empty + a_slot, (a_n - 1) & 1, prm.err, prm.wait_timeout_ns);
                            if (!ok) break;
                            if constexpr (G == 1) {
                                mbar_expect_tx(a_full + a_slot, Lay::A_BYTES);
                                tma_load_2d(As + a_slot * Lay::A_BYTES, &maps.a, c * 128, p * P * ld + arow, a_full + a_slot);
                            } else {
                                if (rank == 0) mbar_expect_tx(a_full + a_slot, 2 * Lay::A_BYTES);
                                tma_load_2d_pair(As + a_slot * Lay::A_BYTES, &maps.a, c * 128, p * P * ld + arow, map_to_cta(smem_u32(a_full + a_slot), 0));
                            }
                            if (++a_slot == NA) a_slot = 0, ++a_n;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ---- MMA issuer (the leader CTA issues for the pair; all 32 lanes run the control flow, one issues) ------------------------------------------------
            int a_slot = 0, a_par = 0, b_slot = 0, b_par = 0, use = 0;
            bool ok = true;
            const uint64_t da0 = oz_desc(smem_u32(As)), db0 = oz_desc(smem_u32(Bs));
#if OZ_STATS
            long long oz_acc_[4] = {0, 0, 0, 0};
            const long long t_all_ = clock64();
#define OZ_EXTRA , oz_acc_
#else
#define OZ_EXTRA
#endif
            for (int unit = unit0; unit < n_units && ok; unit += unit_stride) {
                const int clo = oz2_decode<G>(unit, per_p, prm, 0).clo;
                for (int pass = 0; pass < 2 && ok; ++pass, ++use) {
                    {
                        OZ_T0();
                        if (use >= 1) ok = oz_wait(acc_empty, (use - 1) & 1, prm.err, prm.wait_timeout_ns);  // the previous sums have left TMEM
                        OZ_ACC(3);
                    }
                    if (!ok) break;
                    tc_fence_after();
                    for (int c = clo; c < prm.chi && ok; ++c) {
                        const uint32_t first = (c == clo) ? 0u : 1u;
                        ok = pass ? oz2_chunk<G, 1>(a_full, a_empty, b_full, b_empty, da0, db0, tmem, first, a_slot, a_par, b_slot, b_par, prm OZ_EXTRA)
                                  : oz2_chunk<G, 0>(a_full, a_empty, b_full, b_empty, da0, db0, tmem, first, a_slot, a_par, b_slot, b_par, prm OZ_EXTRA);
                    }
                    if (ok && oz_elect_one()) oz2_commit<G>(acc_full);
                    __syncwarp();
                }
            }
#if OZ_STATS
            if (lane == 0) {
                atomicAdd((unsigned long long*)&oz_stats[0], (unsigned long long)(clock64() - t_all_));
                for (int e = 1; e < 4; ++e) atomicAdd((unsigned long long*)&oz_stats[e], (unsigned long long)oz_acc_[e]);
                atomicAdd((unsigned long long*)&oz_stats[5], (unsigned long long)(use / 2));
            }
#endif
        }
    } else {
        // ---- epilogue: eight warps, thread = 64 columns of one row of this CTA's 128 x 128 tile -------------------
        // While the tensor pipe waits (after either pass): the four integer sums of every entry leave TMEM and are combined
        // exactly in int64; pass 0 keeps them as one FP64 number per entry in registers, pass 1 folds its own onto them
        // (one fma: t0 + t1 2^-24, the only rounding before the final subtraction), then the accumulators are handed back.
        // Next to the following unit's products: scale by the two rows' powers of two and subtract from the tile row in L
        // (read-modify-write; the row was prefetched into L2 when the unit started).  Reading the sums out of TMEM is the
        // floor of the first part: 256 KB per pass at 64 B/clk.
        const int ew = warp & 3;          // the TMEM lanes [32 ew, 32 ew + 32) are the ones this warp may read
        const int hh = (warp - 2) >> 2;   // column half
        const int row = ew * 32 + lane;
        const uint32_t acc_empty_leader = (G == 2) ? map_to_cta(smem_u32(acc_empty), 0) : 0u;
        int use = 0;
        bool ok = true;
        for (int unit = unit0; unit < n_units && ok; unit += unit_stride) {
            const Oz2Unit it = oz2_decode<G>(unit, per_p, prm, (int)rank);
            const double sr = __ldg(prm.rscale + 2 * ((long long)it.p * ld + it.i * 128 + row));
            const double* sc = prm.rscale + 2 * ((long long)it.p * ld + it.k * 128 + hh * 64);
            double* Trow = prm.L + (long long)it.p * prm.mat_stride + (long long)(it.i * 128 + row) * ld + it.k * 128 + hh * 64;
            // only what is stored back is read: the strict upper triangle of a diagonal tile keeps its Gram values (it is never
            // written by the Gram fill), and a pair's idle CTA has nothing to do
            const int cmax = !it.valid ? -1 : (it.i == it.k) ? row - hh * 64 : 63;
            if (cmax >= 0) {
#pragma unroll
                for (int e = 0; e < 4; ++e) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(Trow + 16 * e));
            }
            double d[64];
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                ok = ok && oz_wait(acc_full, (use + pass) & 1, prm.err, prm.wait_timeout_ns);
                if (!ok) break;
                tc_fence_after();
#if OZ_STATS
                const long long t_epi_ = clock64();
#endif
#pragma unroll
                for (int cb = 0; cb < 8; ++cb) {
                    uint32_t a[4][8];
                    const uint32_t taddr = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(hh * 64 + cb * 8);
#pragma unroll
                    for (int g = 0; g < (pass ? 3 : 4); ++g) tmem_ld8(taddr + (uint32_t)g * 128u, a[g]);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // exact in int64 (|sum| < 2^31 per group); group g carries the weight 2^(-14 - 8 g)
                        if (pass == 0) {
                            const long long t = ((long long)(int)a[0][j] << 24) + ((long long)(int)a[1][j] << 16) + ((long long)(int)a[2][j] << 8) + (long long)(int)a[3][j];
                            d[cb * 8 + j] = (double)t;                                    // groups 0..3, in units of 2^-38
                        } else {
                            const long long t = ((long long)(int)a[0][j] << 16) + ((long long)(int)a[1][j] << 8) + (long long)(int)a[2][j];
                            d[cb * 8 + j] = fma((double)t, 0x1p-24, d[cb * 8 + j]);       // groups 4..6, in units of 2^-62 = 2^-38 2^-24
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (G == 1) mbar_arrive(acc_empty);
                    else mbar_arrive_cluster(acc_empty_leader);
                }
#if OZ_STATS
                if (tid == 64 && rank == 0) atomicAdd((unsigned long long*)&oz_stats[4], (unsigned long long)(clock64() - t_epi_));
#endif
            }
            use += 2;
            if (!ok) break;
            const double w = 0x1p-38 * sr;  // exact: powers of two
            if (cmax >= 63) {
                double2 cur[4], nxt[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) cur[e] = __ldcg(reinterpret_cast<const double2*>(Trow) + e);
#pragma unroll
                for (int cb = 0; cb < 8; ++cb) {
                    if (cb < 7) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) nxt[e] = __ldcg(reinterpret_cast<const double2*>(Trow + cb * 8 + 8) + e);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const double s0 = w * __ldg(sc + 2 * (cb * 8 + 2 * e)), s1 = w * __ldg(sc + 2 * (cb * 8 + 2 * e + 1));
                        cur[e].x = fma(-d[cb * 8 + 2 * e], s0, cur[e].x);
                        cur[e].y = fma(-d[cb * 8 + 2 * e + 1], s1, cur[e].y);
                        reinterpret_cast<double2*>(Trow + cb * 8)[e] = cur[e];
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) cur[e] = nxt[e];
                }
            } else if (cmax >= 0) {
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c <= cmax) Trow[c] = fma(-d[c], w * __ldg(sc + 2 * c), __ldcg(Trow + c));
            }
        }
    }
    tc_fence_before();
    if constexpr (G == 2) cluster_sync_all();
    else __syncthreads();
    if (warp == 0) {
        if constexpr (G == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u));
    }
}

// ---- row scales -----------------------------------------------------------------------------------
__global__ void agp_ozaki_rowscale_kernel(const double* __restrict__ L, long long mat_stride, int ld, int P, double* __restrict__ rscale,
                                          int ld_obs, const double* __restrict__ noise) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (long long)P * ld) return;
    const int p = (int)(w / ld), r = (int)(w - (long long)p * ld);
    int e = 0;
    if (r >= ld_obs) {
        // appended row of an identity-augmented batch: it ends as a row of L^{-T}, |entries| <= sqrt((K^{-1})_rr) <= 1 / sqrt(noise);
        // one extra bit for kernel matrices that are PSD only up to rounding; padding columns hold 1
        const double nz = noise[p];
        if (nz > 0.0 && nz < 1e300) {
            int ex;
            frexp(1.0 / nz, &ex);
            e = (ex + 1) >> 1;
        }
        e = (e > 0 ? e : 0) + 1;
    } else {
        const double kd = L[(long long)p * mat_stride + (long long)r * ld + r];
        if (kd > 0.0 && kd < 1e300) {
            int ex;
            frexp(kd, &ex);        // kd = m 2^ex, 1/2 <= m < 1: sqrt(kd) < 2^(ex / 2)
            e = (ex + 1) >> 1;     // ceil(ex / 2)
            // seven balanced base-256 digits reach 0.996 of the scale: keep |L_ij| <= sqrt(K_rr) below 0.99 of it
            if (kd > 0.98 * ldexp(1.0, 2 * e)) e += 1;
        }
    }
    rscale[2 * w] = ldexp(1.0, e);
    rscale[2 * w + 1] = ldexp(1.0, 55 - e);
}

// ---- digit planes ---------------------------------------------------------------------------------
// One CTA = one 128 x 128 tile of L; a warp walks rows, lane = 4 consecutive columns: every plane receives 128
// contiguous bytes per row (digits: agp_ozaki_digits.cuh).
__global__ void __launch_bounds__(256) agp_ozaki_slice_kernel(const double* __restrict__ L, long long mat_stride, int ld, int P,
                                                              const double* __restrict__ rscale, int8_t* __restrict__ S, int c0, int ncol, int r0) {
    const int p = blockIdx.z, i = r0 + blockIdx.y, j = c0 + blockIdx.x;
    (void)ncol;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long plane = (long long)P * ld * ld;
    for (int rr = warp; rr < 128; rr += 8) {
        const int r = i * 128 + rr;
        const double f = __ldg(rscale + 2 * ((long long)p * ld + r) + 1);
        const double* src = L + (long long)p * mat_stride + (long long)r * ld + j * 128 + lane * 4;
        const double2 x01 = __ldcs(reinterpret_cast<const double2*>(src));
        const double2 x23 = __ldcs(reinterpret_cast<const double2*>(src) + 1);
        oz_store_digits4(x01.x, x01.y, x23.x, x23.y, f, S + ((long long)p * ld + r) * ld + j * 128 + lane * 4, plane);
    }
}

// ---- host side --------------------------------------------------------------------------------------
void launch_ozaki_rowscale(const double* L, long long mat_stride, int ld, int P, double* rscale, cudaStream_t s, int ld_obs, const double* noise) {
    const long long n = (long long)P * ld;
    if (n <= 0) return;
    agp_ozaki_rowscale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(L, mat_stride, ld, P, rscale, noise ? ld_obs : (1 << 30), noise);
}

void launch_ozaki_slice(const double* L, long long mat_stride, int ld, int nt, int P, const double* rscale, int8_t* S, int c0, int c1, int r0,
                        cudaStream_t s) {
    if (c1 <= c0 || r0 >= nt || P <= 0) return;
    dim3 grid(c1 - c0, nt - r0, P);
    agp_ozaki_slice_kernel<<<grid, 256, 0, s>>>(L, mat_stride, ld, P, rscale, S, c0, c1 - c0, r0);
}

void launch_ozaki_update(const OzakiParams& prm_in, const OzakiMaps& maps, int ctas, cudaStream_t s, int variant) {
    OzakiParams prm = prm_in;
    if (prm.r_hi <= prm.r_lo) {  // plain form: lower tiles of block columns [c0, c1) over [0, c0)
        prm.r_lo = prm.c0, prm.r_hi = prm.nt, prm.k_lo = prm.c0, prm.k_hi = prm.c1, prm.chi = prm.c0;
    }
    if (prm.r_hi <= prm.r_lo || prm.k_hi <= prm.k_lo || prm.chi <= 0 || prm.P <= 0) return;
    if (variant == 2) {
        agp_ozaki_update2_kernel<1><<<ctas, OZ2_THREADS, Oz2<1>::SMEM, s>>>(prm, maps);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(ctas & ~1), 1, 1);
        cfg.blockDim = dim3(OZ2_THREADS, 1, 1);
        cfg.dynamicSmemBytes = Oz2<2>::SMEM;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, agp_ozaki_update2_kernel<2>, prm, maps);
    }
}

cudaError_t configure_ozaki() {
    cudaError_t e = cudaFuncSetAttribute(agp_ozaki_update2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz2<1>::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(agp_ozaki_update2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz2<2>::SMEM);
}

bool make_ozaki_maps(int8_t* S, int ld, int P, OzakiMaps* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return false;
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)OZ_SLICES * (cuuint64_t)P * (cuuint64_t)ld};
    const cuuint64_t strides[1] = {(cuuint64_t)ld};
    const cuuint32_t box_a[2] = {128, 128}, box_b[2] = {128, 64}, estr[2] = {1, 1};
    const CUresult r1 = encode(&out->a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, S, dims, strides, box_a, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const CUresult r2 = encode(&out->b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, S, dims, strides, box_b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS;
}

}  // namespace agp
