// ITEM_SLICE of agp_chol_kernel (hybrid schedule, agp_ozaki.cu): the int8 digit planes of 64 rows of a finished panel
// tile, cut by whatever CTA is free while the segment's POTF2 / panel chain goes on.  As a launch of its own between the
// segment and the int8 update the same work is pure waiting time for everybody (0.38 ms of a 6.7 ms step at n = 2048 x 64);
// as queue items it is HBM traffic next to CTAs that mostly spin on dependencies.  Compiled as its own translation unit
// and called through the plain ABI (agp_chol_common.cuh says why).
#include "agp_chol_common.cuh"
#include "agp_ozaki_digits.cuh"

namespace agp {

__device__ bool do_slice(const BatchView& v, const SchedView& q, int idx) {
    const Smem s = smem_view();
    const ItemFields f = decode_item(q, idx);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (f.need_i > 0) {
        // both halves of the tile are final once tile row i has finished its panels up to block column k
        if (tid == 0) s.ctl[1] = wait_ge(q.rowdone + f.p * q.nt_stride + f.i, f.need_i, q.err, q.wait_timeout_ns) ? 1 : 0;
        __syncthreads();
        if (!s.ctl[1]) return false;
    }
    stamp(q, idx, 1);
    const int ld = v.ld;
    const long long plane = v.oz_plane;
    const double* __restrict__ Lp = v.L + (long long)f.p * v.mat_stride;
#pragma unroll 2
    for (int rr = warp; rr < UM; rr += FT / 32) {
        const int r = f.i * TB + f.h * UM + rr;
        const double fac = __ldg(v.oz_rscale + 2 * ((long long)f.p * ld + r) + 1);
        const double* src = Lp + (long long)r * ld + f.k * TB + lane * 4;
        // written by other CTAs of this launch: bypass L1
        const double2 x01 = __ldcg(reinterpret_cast<const double2*>(src));
        const double2 x23 = __ldcg(reinterpret_cast<const double2*>(src) + 1);
        oz_store_digits4(x01.x, x01.y, x23.x, x23.y, fac, v.oz_S + ((long long)f.p * ld + r) * ld + f.k * TB + lane * 4, plane);
    }
    __syncthreads();  // the next item reuses ctl[]
    return true;
}

}  // namespace agp
