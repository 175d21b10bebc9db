// Phase B of the final PANEL work items of agp_chol_kernel: the triangular solve as a tensor-core product with the
// inverted diagonal tile, store, folded forward solve.  Compiled as its own translation unit (agp_chol_common.cuh says why).
#include "agp_chol_common.cuh"

namespace agp {

// L_ik = X W^T with W = L_kk^{-1} (lower triangular, written by POTF2(k) as a dense 128x128 tile): the panel's
// triangular solve as ONE more tensor-core contraction, 64 x 128 x 128 with the zero half of W skipped.  X (A operand)
// stays in shared memory, W (B operand) streams through two 16 KB stages by TMA, eight chunks of 16 columns, on the
// same full / empty barriers as the main loop (the chunk counter keeps counting).  Warp (wm, wn) owns the 32 x 32 tile
// of columns wn: it needs k < 32 (wn + 1) only, i.e. the chunks 0 .. 2 wn + 1.  The result leaves the accumulators for L
// directly; the forward solve y_i -= L_ik z_k is folded in (fixed summation order).
__device__ __forceinline__ bool solve_store(const TmaMaps& maps, const SchedView& q, double* __restrict__ ctile, int ld, int wrow, const double* __restrict__ zk,
                                         double* __restrict__ yrow, const double* __restrict__ yinit_src, int yinit_valid) {
    const Smem s = smem_view();
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, c4 = lane & 3;
    double* Wst = s.region;
    const double* Xs = s.region + XS_OFF;
    constexpr int NCH = TB / KC;  // 8
    const int G1 = s.ctl[4];
    auto produce = [&](int c) {  // thread 0 only
        const int G = G1 + c, bi = G % NSTAGE;
        if (G >= NSTAGE && !mbar_wait_bounded(s.empty + bi, ((G / NSTAGE) - 1) & 1, q.err, q.wait_timeout_ns)) return;
        if (c >= 2 && !mbar_wait_bounded(s.empty + ((G - 2) % NSTAGE), ((G - 2) / NSTAGE) & 1, q.err, q.wait_timeout_ns)) return;  // the buffer's last reader
        mbar_expect_tx(s.full + bi, WST_D * 8);
        tma_load_2d(Wst + (c & 1) * WST_D, &maps.w, c * KC, wrow, s.full + bi);
    };
    if (tid == 0) {
        fence_proxy_async_all();  // acquire of fdone (POTF2's generic-proxy stores of W) -> async-proxy reads of W
        produce(0);
        produce(1);
    }
    if (tid < TB) s.zs[tid] = __ldcg(zk + tid);
    double acc[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    const int my_chunks = 2 * wn + 2;
    for (int ch = 0; ch < NCH; ++ch) {
        const int G = G1 + ch, bi = G % NSTAGE;
        if (!mbar_wait_bounded(s.full + bi, (G / NSTAGE) & 1, q.err, q.wait_timeout_ns)) return false;
        const double* Bs = Wst + (ch & 1) * WST_D;
        const double* As = Xs + ch * KC;
        const bool mine = ch < my_chunks;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            double2 a[4], b[4];
            if (mine) {
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) a[mb] = *reinterpret_cast<const double2*>(As + (wm * 32 + mb * 8 + g) * XS2 + 2 * (2 * c4 + ks));
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) b[nb] = *reinterpret_cast<const double2*>(Bs + swz128(wn * 32 + nb * 8 + g, 2 * c4 + ks));
            }
            if (mine) {
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].x, b[nb].x);
            }
            if (ks == 1) {
                fence_proxy_async();  // see contract(): the stage is read with LDS and rewritten by TMA
                __syncwarp();
                if (lane == 0) mbar_arrive(s.empty + bi);
                if (tid == 0 && ch + 2 < NCH) produce(ch + 2);
            }
            if (mine) {
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].y, b[nb].y);
            }
        }
    }
    if (tid == 0) s.ctl[4] = G1 + NCH;
    // store L_ik from the fragments and fold the forward solve: y_i -= L_ik z_k
    double* part = s.region;  // [UM][4] partial dot products per row and column block (the W stages are free: see the barrier below)
    __syncthreads();          // z_k in shared memory; every warp is done with both W stages
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        const int r = wm * 32 + mb * 8 + g;
        double dsum = 0.0;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const int c = wn * 32 + nb * 8 + 2 * c4;
            *reinterpret_cast<double2*>(ctile + (long long)r * ld + c) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
            dsum = fma(acc[mb][nb][0], s.zs[c], dsum);
            dsum = fma(acc[mb][nb][1], s.zs[c + 1], dsum);
        }
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
        if (c4 == 0) part[r * 4 + wn] = dsum;
    }
    __syncthreads();
    if (tid < UM) {
        const double y_old = yinit_src ? (tid < yinit_valid ? yinit_src[tid] : 0.0) : __ldcg(yrow + tid);  // first panel of the tile row: y starts from xs (0 in the padding)
        yrow[tid] = y_old - (((part[tid * 4] + part[tid * 4 + 1]) + part[tid * 4 + 2]) + part[tid * 4 + 3]);
    }
    return true;
}

// phase B of a final PANEL item: the triangular product with W = L_kk^{-1}, store, forward solve
__device__ int update_solve(const BatchView& v, const SchedView& q, const TmaMaps& maps, int idx) {
    const Smem s = smem_view();
    const ItemFields f = decode_item(q, idx);
    const int tid = threadIdx.x;
    const int p = f.p, k = f.k, i = f.i;
    const int row0 = i * TB + f.h * UM, col0 = k * TB;
    const int ld = v.ld;
    stamp(q, idx, 3);
    if (tid == 0) s.ctl[1] = wait_ge(q.fdone + p, k + 1, q.err, q.wait_timeout_ns) ? 1 : 0;
    __syncthreads();  // also publishes X
    if (!s.ctl[1]) return 0;
    stamp(q, idx, 4);
    double* ctile = v.L + (long long)p * v.mat_stride + (long long)row0 * ld + col0;
#if AGP_X_SKIP_SOLVE
    signal_done(q.rowdone + p * q.nt_stride + i);
    return 1;
#endif
    if (!solve_store(maps, q, ctile, ld, (p * q.nt_stride + k) * TB, v.z + (long long)p * ld + col0, v.y + (long long)p * ld + row0,
                     f.yinit ? v.xs + row0 : nullptr, v.n - row0))
        return 0;
    signal_done(q.rowdone + p * q.nt_stride + i);
    return 1;
}

}  // namespace agp
