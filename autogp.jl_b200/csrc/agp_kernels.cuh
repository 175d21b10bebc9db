// Kernel launch interface shared by agp_kernels.cu and agp_api.cu.
#pragma once
#include <cuda.h>  // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_runtime.h>
#include <stdint.h>

#include "agp_program.h"

namespace agp {

constexpr int TB = 128;  // tile edge = Cholesky block-column width

// Device view of one resident batch of particles (all pointers are device pointers).
struct BatchView {
    double* L;               // [P][ld][ld] row-major; lower-triangle tiles hold K, then L (== column-major U)
    long long mat_stride;    // ld*ld
    int ld;                  // row stride (allocated npad)
    int n;                   // active number of observations (data prefix)
    int nt;                  // active tiles = ceil(n / TB)
    const double* ts;        // [ld]   time points, zero padded
    const double* xs;        // [ld]   observations, zero padded
    double* y;               // [P][ld] forward-substitution work vector
    double* z;               // [P][ld] z = L^{-1} xs
    const AgpInstr* prog;    // concatenated device programs
    const int* prog_off;     // [P+1]
    const int* prog_need;    // [P] register-stack depth
    int max_prog_len;        // longest program of the batch (instructions)
    const double* noise;     // [P]
    double* lml;             // [P]
    int* info;               // [P]
    double* dinv;            // [P][ld/128][4][32][32] inverses of the diagonal 32x32 blocks of every L_kk
    // rows beyond the factored block (prediction points appended at a tile
    // boundary, see agp_predict_batch) and per-block-column running sums (so a factorisation can be
    // continued from any block column, see agp_lml_run_append)
    int nt_total;            // tile rows in the matrix: nt + ceil(n_pred / TB)
    int n_pred;              // appended time points (0 for a plain LML batch)
    double* cum;             // [P][ld/TB][2] running (sum log L_ii, sum z_i^2) after each block column
    int aug_identity;        // appended rows carry [I 0] instead of kernel values: their Schur complement is -K^{-1}
                             // and their forward-solve entries are -alpha (agp_lml_grad_batch)
    // hybrid schedule (agp_ozaki.cu), for the SLICE items of the persistent kernel
    signed char* oz_S;       // digit planes [8][P][ld][ld]
    const double* oz_rscale; // [P][ld][2] row scale, digit factor
    long long oz_plane;      // bytes between two planes = P ld ld
};

// ---- persistent dataflow scheduler (agp_fused.cu) ------------------------------------------
// Work items: two int4 each,
//   {type | h << 8 | ITEM_PARTIAL?, particle, block column k, tile row i}   ITEM_PARTIAL: store-only item,
//                          the tile in L receives K - sum_{j0<=j<j1} (no solve, no factorisation follows from it)
//                          ITEM_YINIT: this panel is the first of its tile row (starts the forward-solve entry from xs)
//   {j0 | j1 << 16, need_k | need_i << 16, extra_flag, extra_need}: contraction range [j0, j1) in block
//   columns; the values rowdone[p][k] / rowdone[p][i] must have reached (finished panel items of the
//   two operand tile rows); for a continuation item, the index (relative to SchedView::head) and
//   value of the counter its predecessor bumps; POTF2 carries in extra_need how many DIAG items
//   finish its tile.
//   ITEM_GRAM: {ITEM_GRAM | h << 8, particle, block column k, tile row i}, {0, 0, flag, 0}: the Gram unit of that tile
//   half; bumps counter `flag` (relative to SchedView::head), which the first PANEL item of the tile half carries as
//   its extra_flag with extra_need = 1 (the two units of a diagonal tile bump one flag, its first DIAG items need 2).
//   ITEM_SLICE: {ITEM_SLICE | h << 8, particle, block column k, tile row i}, {0, need_i << 16, -1, 0}: the int8 digit planes of that
//   finished tile half (hybrid schedule); waits for rowdone[p][i] >= need_i, nobody in the launch waits for it.
enum { ITEM_DIAG = 0, ITEM_POTF2 = 1, ITEM_PANEL = 2, ITEM_GRAM = 3, ITEM_SLICE = 4, ITEM_PARTIAL = 1 << 9, ITEM_YINIT = 1 << 10 };

struct SchedView {
    const int4* items;  // in-order queue (2 x int4 per item): every item's producers sit earlier in the list
    int n_items;
    int* head;          // queue head (atomic ticket)
    int* err;           // set when a dependency wait timed out (never in a valid schedule)
    int* rowdone;       // [P][nt_stride] finished 64-row panel items of tile row i (2 per block column)
    int* diagu;         // [P][nt_stride] finished DIAG items of diagonal tile k
    int* ppre;          // [P][nt_stride] finished PARTIAL panel items of tile row i (look-ahead)
    int* fdone;         // [P] factored diagonal tiles
    int nt_stride;
    unsigned long long wait_timeout_ns;  // dependency wait limit before the error flag is raised
    long long* trace;   // optional [n_items][8] globaltimer stamps (diagnostics; nullptr = off)
};

// Gram fill: lower tiles of tile rows [row_tile0, nt_total) of every particle  <-  K(ts,ts) + noise*I
// (runs before launch_chol)
void launch_gramfill(const BatchView& v, int P, int row_tile0, cudaStream_t s);
// dLML/dparams, dLML/dnoise out of the identity-augmented factorisation (agp_lml_grad_batch):
// partial[P][blocks][AGP_GRAD_MAX_PARAMS + 1] per-CTA sums, then grad_out / gnoise_out
constexpr int AGP_GRAD_MAX_PARAMS = 64;
// returns the number of launches; big: a kernel of the batch exceeds the hot variant's limits (agp_fused.cu)
int launch_grad(const BatchView& v, int P, const int* param_off, double* partial, double* grad_out, double* gnoise_out, int max_params, bool big,
                cudaStream_t s);
int grad_blocks_per_particle(const BatchView& v);
// dLML/dnoise alone out of factorisation + trtri (agp_lml_grad_noise_batch): partial[P][blocks] per-CTA sums
void launch_noise_grad(const BatchView& v, int P, double* partial, double* gnoise_out, cudaStream_t s);
int noise_grad_blocks_per_particle(const BatchView& v);
// [I 0] rows and the zero trailing block of an identity-augmented batch (tile rows nt .. nt_total)
void launch_augfill(const BatchView& v, int P, cudaStream_t s);
void launch_predict_extract_marginals(const BatchView& v, int P, const double* noise_pred, double* mean_out, double* var_out, cudaStream_t s);
// Summand programs of agp_predict_sum_batch: component c of particle p = instructions [off[p M + c], off[p M + c + 1])
struct ComponentView {
    const AgpInstr* prog;
    const int* off;   // [P * M + 1]
    const int* need;  // [P * M] operand-stack depth
    int M;            // summands per particle
    int m_each;       // prediction points
};
// rewrites the appended rows of a batch uploaded with the sum kernel and (M + 1) copies of the prediction points
// (runs between launch_gramfill and launch_chol)
void launch_component_fill(const BatchView& v, int P, const ComponentView& cv, cudaStream_t s);
// Predictive mean / covariance out of an augmented factorisation (agp_predict_batch)
void launch_predict_extract(const BatchView& v, int P, const double* noise_pred, double* mean_out, double* cov_out, cudaStream_t s);
// 2-D TMA descriptors over L viewed as one [P * ld][ld] FP64 matrix: boxes of 16 columns (128 bytes, hardware
// 128-byte swizzle) x 64 rows (A operand: the item's rows) and x 128 rows (B operand: tile row k)
struct TmaMaps {
    CUtensorMap a, b;
};
// returns false when the driver refuses a descriptor (reported by the caller)
bool make_tma_maps(double* L, int ld, long long rows, TmaMaps* out);
// One launch = the whole batch: Cholesky + solve + logdet for every particle.
void launch_chol(const BatchView& v, const SchedView& q, const TmaMaps& maps, int ctas, cudaStream_t s);
cudaError_t configure_fused();
#ifndef launch_chol  // the second instantiation (Makefile: RDC_SOLO) for one CTA per SM; inside its own translation units the names above ARE these
void launch_chol_solo(const BatchView& v, const SchedView& q, const TmaMaps& maps, int ctas, cudaStream_t s);
cudaError_t configure_fused_solo();
#endif

// Stand-alone Gram matrix (drop-in for compute_cov_matrix[_vectorized]): column-major, both triangles
void launch_gram(const AgpInstr* prog, int m, int need, const double* ts, int n, double noise, int form,
                 double* K, cudaStream_t s);

}  // namespace agp
