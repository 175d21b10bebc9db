// Digit planes of four consecutive FP64 entries of a row (shared by agp_ozaki_slice_kernel and the SLICE items of the
// persistent kernel, agp_chol_slice.cu).  An entry is rounded once to 55 bits below its row scale, v = rint(x 2^(55 - e)),
// and v = sum_p a_p 256^(6 - p) is peeled into balanced base-256 digits from the bottom (integer arithmetic: exact):
// x = 2^e sum_p a_p 2^(-7 - 8 p), a_p in [-128, 127].  Seven such digits reach +-127 (256^7 - 1) / 255 = 0.99608 2^55: the row
// scales are chosen with |x| <= 0.99 2^e (agp_ozaki_rowscale_kernel), the clamp below only guards against garbage rows of
// a failed factorisation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace agp {

constexpr int OZ_DIGITS = 7;

// f = 2^(55 - e); dst = plane 0 address of the first entry; plane = bytes between two planes
__device__ __forceinline__ void oz_store_digits4(double x0, double x1, double x2, double x3, double f, int8_t* dst, long long plane) {
    long long v[4] = {__double2ll_rn(x0 * f), __double2ll_rn(x1 * f), __double2ll_rn(x2 * f), __double2ll_rn(x3 * f)};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const long long lim = 35887507618889599ll;  // 127 (256^7 - 1) / 255
        v[e] = v[e] < -lim ? -lim : (v[e] > lim ? lim : v[e]);
    }
#pragma unroll
    for (int q = OZ_DIGITS - 1; q >= 0; --q) {
        uint32_t word = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            long long d;
            if (q > 0) {
                d = ((v[e] + 128) & 255) - 128;  // balanced digit in [-128, 127]
                v[e] = (v[e] - d) >> 8;          // exact
            } else {
                d = v[e];                        // the leading digit: within [-127, 127]
            }
            word |= ((uint32_t)(d & 0xff)) << (8 * e);
        }
        *reinterpret_cast<uint32_t*>(dst + (long long)q * plane) = word;
    }
}

}  // namespace agp
