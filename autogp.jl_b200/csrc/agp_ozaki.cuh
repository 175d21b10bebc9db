// Error-free int8 split of the factorisation's long contractions onto the 5th-generation tensor cores
// (tcgen05.mma kind::i8, int32 accumulators in TMEM) — interface of agp_ozaki.cu.
//
// tcgen05.mma has no f64 kind, so the FP64 tensor path of sm_100a is DMMA (mma.sync m8n8k4) at 64 FMA/clk/SM.  The
// int8 kind runs at 8192 MAC/clk/SM.  An FP64 product  sum_j L_ij L_kj^T  is rebuilt from exact integer products:
// every row of L is scaled by a power of two known BEFORE the factorisation (|L_ij| <= sqrt(K_ii) <= 2^e_i), the scaled
// entry is rounded once to 55 bits below that scale and cut into seven balanced base-256 digits
// x = 2^e sum_p a_p 2^(-7-8p)  (a_p in [-128, 127]: exactly the range of a signed byte), and the 28 digit-plane products
// with p + q <= 6 are accumulated EXACTLY in int32, one accumulator per weight g = p + q, over the WHOLE contraction
// depth (|sum| <= 7 depth 2^14 < 2^31 up to depth 18 724).  The integer sums are recombined once per output tile in FP64.
// Error of the scheme on the benchmark factor (tests/test_ozaki_scheme.py): 4.0e-15 of sqrt(K_ii K_kk) for a 1920-deep
// contraction — plain FP64 accumulation: 2.8e-15.  (Round 2 started with eight 7-bit digits, 36 products, 2.6e-14:
// tools/ozaki_accuracy.py; full bytes cost nothing in the int32 sums and save a plane and eight products.)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace agp {

constexpr int OZ_SLICES = 7;

struct OzakiMaps {
    CUtensorMap a, b;  // digit planes viewed as one [8 P ld][ld] byte matrix: boxes of 128 bytes x 128 rows (A) / x 64 rows (B), 128-byte swizzle
};

struct OzakiParams {
    double* L;             // [P][ld][ld] the factor / running Schur complement (row-major lower)
    long long mat_stride;  // ld * ld
    int ld, nt, P;
    const double* rscale;  // [P][ld][2]: row scale 2^e, digit factor 2^(55 - e)
    int c0, c1;            // tiles (i, k), c0 <= k < c1, k <= i < nt:  T_ik -= sum_{j < c0} L_ij L_kj^T
    int* err;              // raised when a barrier wait times out (never in a correct run)
    unsigned long long wait_timeout_ns;
    // generalised form (second-generation kernels): tile rows r_lo <= i < r_hi against tile rows k_lo <= k < k_hi, k <= i,
    // contraction over block columns [max(fc(i), fc(k)), chi), fc(r) = r - nt for the appended rows r >= nt of an
    // identity-augmented batch (their tiles left of block column r - nt are structurally zero), else 0
    // (all zero: the plain form above)
    int r_lo, r_hi, k_lo, k_hi, chi;
};

// rscale[p][r] <- power-of-two bound of sqrt(K_rr), read from the diagonal of L right after the Gram fill
// rows >= ld_obs (appended [I 0] rows of an identity-augmented batch, which end as rows of L^{-T}): bound 1 / sqrt(noise),
// since sum_j (L^{-1})_jr^2 = (K^{-1})_rr <= 1 / lambda_min(K) <= 1 / noise for K = (PSD kernel matrix) + noise I
void launch_ozaki_rowscale(const double* L, long long mat_stride, int ld, int P, double* rscale, cudaStream_t s, int ld_obs = 1 << 30,
                           const double* noise = nullptr);
// digit planes of the tiles (i, j), c0 <= j < c1, r0 <= i < nt (all 128 rows, 128 columns each)
void launch_ozaki_slice(const double* L, long long mat_stride, int ld, int nt, int P, const double* rscale, int8_t* S, int c0, int c1, int r0,
                        cudaStream_t s);
// the contraction over block columns [0, c0) of every lower tile of block columns [c0, c1), subtracted in place
// variant 2: one CTA per unit; 3: CTA pairs (cta_group::2)
void launch_ozaki_update(const OzakiParams& prm, const OzakiMaps& maps, int ctas, cudaStream_t s, int variant = 3);
bool make_ozaki_maps(int8_t* S, int ld, int P, OzakiMaps* out);
cudaError_t configure_ozaki();

}  // namespace agp
