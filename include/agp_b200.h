/*
 * agp_b200.h — C-ABI of the B200-native GP log-marginal-likelihood engine.
 *
 * Drop-in boundary for the ONE hot path of probsys/AutoGP.jl (reference @ 2ad372d):
 *
 *     noise      = transform_param(:noise, z) + JITTER                 src/Model.jl:134
 *     cov_matrix = GP.compute_cov_matrix_vectorized(node, noise, ts)   src/Model.jl:135 -> src/GP.jl:666-668
 *     xs ~ mvnormal(zeros(n), cov_matrix)                              src/Model.jl:136 (Gen -> Distributions -> PDMats -> dpotrf)
 *
 * The reference has no FFI layer; these are the entry points a Julia `ccall` (or Python
 * ctypes) binding for that path would bind.  Plain pointers and sizes
 * only.  Every function is re-entrant per handle; the library keeps no global mutable state.
 * Return value: 0 = success, <0 = error (text via agp_last_error); never throws/aborts.
 *
 * Kernel programs ("wire format"): a kernel tree is sent in the reference's own postfix
 * order, i.e. the order of `GP.unroll(node)` (src/GP.jl:111-113), as
 *   ops[m]        int32  node-type code  (GPConfig codes, src/GP.jl:1101-1108, + 9 = WhiteNoise)
 *   param_off[m]  int32  index of the node's first parameter in params[] (ignored for Plus/Times)
 *   params[]      f64    leaf/ChangePoint parameters in Julia `fieldnames` order:
 *       Constant(value) | Linear(intercept,bias,amplitude) | SquaredExponential(lengthscale,amplitude)
 *       GammaExponential(lengthscale,gamma,amplitude) | Periodic(lengthscale,period,amplitude)
 *       WhiteNoise(value) | ChangePoint(location,scale)          (src/GP.jl:131-133,157-159,185-192,
 *                                                                  228-234,269-277,315-322,466-473)
 */
#ifndef AGP_B200_H
#define AGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct agp_handle agp_handle;

/* node-type codes == GP.GPConfig defaults (src/GP.jl:1101-1108); WhiteNoise has no code there */
enum {
    AGP_OP_CONSTANT = 1,
    AGP_OP_LINEAR = 2,
    AGP_OP_SQUARED_EXPONENTIAL = 3,
    AGP_OP_GAMMA_EXPONENTIAL = 4,
    AGP_OP_PERIODIC = 5,
    AGP_OP_PLUS = 6,
    AGP_OP_TIMES = 7,
    AGP_OP_CHANGEPOINT = 8,
    AGP_OP_WHITE_NOISE = 9
};

/* error codes */
enum {
    AGP_OK = 0,
    AGP_ERR_ARG = -1,      /* bad argument (null pointer, negative size, ...) */
    AGP_ERR_PROGRAM = -2,  /* malformed kernel program (bad opcode, stack under/overflow, gamma range) */
    AGP_ERR_CUDA = -3,     /* CUDA runtime error */
    AGP_ERR_NOMEM = -4,    /* device workspace allocation failed */
    AGP_ERR_STATE = -5     /* call sequence error (run before upload, ...) */
};

/* Gram evaluation form */
enum {
    AGP_FORM_VECTORIZED = 0, /* eval_cov(node, ts::Vector)  src/GP.jl:137-503, used by Model.model */
    AGP_FORM_SCALAR = 1      /* eval_cov(node, t1, t2)      src/GP.jl:135-491, used by compute_cov_matrix */
};

/* ---- lifetime ------------------------------------------------------------------------ */

/* Create an engine bound to CUDA device `device`; owns one stream + its workspaces.
 * Replaces nothing in the reference (it has no device state); one handle per Julia thread
 * or one shared handle for the batched lock-step path. */
int agp_create(int device, agp_handle** out);
void agp_destroy(agp_handle* h);
/* Size the handle's large device buffers once for the biggest call it will see: `max_n` observations (plus
 * `max_pred` prediction points), `max_batch` particles, `with_gradient` != 0 when agp_lml_grad_batch /
 * agp_lml_grad_noise_batch will be called at that size (their identity-augmented factor has twice the rows).
 * Optional: every call grows what it needs, but a data-annealing run (src/inference_smc_anneal_data.jl:206-217) grows the
 * series round by round, and each growth is a cudaFree + cudaMalloc of gigabytes (20 - 200 ms) in the middle of the loop.
 * This is the `agp_create(device, max_n, max_batch)` of SURVEY.md §8b as a separate call.  Never shrinks. */
int agp_reserve(agp_handle* h, int32_t max_n, int32_t max_pred, int32_t max_batch, int32_t with_gradient);
/* Text of the last error on this handle (valid until the next call on it). */
const char* agp_last_error(const agp_handle* h);
/* "major.minor.patch+sm_100a" */
const char* agp_version(void);

/* ---- site 1: Gram matrix ------------------------------------------------------------- */

/* K_out (host, n*n doubles, column-major, BOTH triangles) = eval_cov(program, ts) + noise*I.
 * Replaces GP.compute_cov_matrix_vectorized (src/GP.jl:666-668) with form = VECTORIZED,
 * GP.compute_cov_matrix (src/GP.jl:674-684) with form = SCALAR, and eval_cov(node, ts)
 * (src/GP.jl:61) with noise = 0.  `ts` is not modified; no pointer is retained. */
int agp_gram(agp_handle* h, const int32_t* ops, const int32_t* param_off, int32_t m,
             const double* params, int32_t n_params, const double* ts, int32_t n, double noise,
             int32_t form, double* K_out);

/* Same, but K_out is a DEVICE pointer (n*n doubles) and the call is asynchronous on the
 * handle's stream: what bench.py times for the HBM-bound Gram stage. */
int agp_gram_device(agp_handle* h, const int32_t* ops, const int32_t* param_off, int32_t m,
                    const double* params, int32_t n_params, const double* ts, int32_t n,
                    double noise, int32_t form, double* K_out_dev);

/* ---- site 2: the log marginal likelihood, batched over particles ----------------------- */

/* lml_out[p] = logpdf(mvnormal(zeros(n), eval_cov(program_p, ts) + noise[p]*I), xs)
 * for p = 0..P-1 — the fused replacement of src/Model.jl:135-136 for P particles that share
 * (ts, xs), as they always do inside one SMC round (src/inference_smc_anneal_data.jl:127-141,
 * 212-217) and one MH/HMC sweep (src/inference_utils.jl:78-119).
 *
 *   prog_len[P]   nodes in particle p's program
 *   ops, param_off  concatenated programs (sum(prog_len) entries); param_off is relative to
 *                   the start of particle p's slice of params
 *   n_params[P]   doubles in particle p's slice of params; params = concatenation
 *   noise[P]      observation noise variance incl. JITTER (src/Model.jl:134)
 *   ts[n], xs[n]  time points (any order, duplicates allowed) and observations
 *   lml_out[P]    result; info_out[P]: 0 = ok, k>0 = leading minor k not positive definite
 *                 (LAPACK dpotrf convention; the Julia glue raises PosDefException(k) as PDMats
 *                 would) and lml_out[p] = NaN.
 * All pointers are HOST pointers; copies in/out and the synchronisation are inside the call. */
int agp_lml_batch(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops,
                  const int32_t* param_off, const int32_t* n_params, const double* params,
                  const double* noise, const double* ts, const double* xs, int32_t n,
                  double* lml_out, int32_t* info_out);

/* The same call split in three so a caller can keep inputs resident on the device and
 * overlap: upload (host -> device, compiles the programs), run (asynchronous on the handle's
 * stream; may be repeated), fetch (device -> host + synchronise). */
int agp_lml_upload(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops,
                   const int32_t* param_off, const int32_t* n_params, const double* params,
                   const double* noise, const double* ts, const double* xs, int32_t n);
int agp_lml_run(agp_handle* h);
int agp_lml_fetch(agp_handle* h, double* lml_out, int32_t* info_out);
/* Device pointers to the P results of the last run (valid until the next upload): lets the
 * multi-GPU caller all-gather log-weights straight from HBM (NCCL). */
int agp_lml_device_results(agp_handle* h, double** lml_dev, int32_t** info_dev);
/* Re-target the resident batch at a data prefix ts[0:n_prefix], xs[0:n_prefix] without a new
 * upload — the data-annealing step of src/inference_smc_anneal_data.jl:212-217. */
int agp_lml_set_prefix(agp_handle* h, int32_t n_prefix);

/* Continue the factorisation of the resident batch after the data prefix GREW (agp_lml_set_prefix
 * with a larger n), keeping every tile row of L, z and the running log-det / quadratic-form sums
 * that the previous run completed: only tile rows >= floor(n_old / 128) are recomputed, O(n^2 k)
 * instead of O(n^3) — the block-append path of BASELINE.json configs[4].  The reference has no
 * counterpart: it re-scores from scratch on every prefix (src/inference_smc_anneal_data.jl:
 * 127-141, src/api.jl:426-443); results are bitwise those of agp_lml_run on the same prefix.
 * Requires: same resident batch (no upload in between), the previous run fetched with info == 0
 * for every particle; otherwise AGP_ERR_STATE and the caller falls back to agp_lml_run. */
int agp_lml_run_append(agp_handle* h);

/* ---- gradient of site 2 ------------------------------------------------------------------------ */

/* lml_out[p] as agp_lml_batch, plus its gradient with respect to every kernel parameter and the noise:
 *   grad_params_out  concatenated like `params` (particle p's slice has n_params[p] entries, in the
 *                    same wire order: Julia fieldnames order per node, nodes in unroll order)
 *   grad_noise_out[P]
 * This is what Gen.hmc (src/inference_utils.jl:63-67) and Gen.map_optimize (src/Greedy.jl:95, 370)
 * obtain from mvnormal's logpdf_grad (cov_deriv = (alpha alpha' - K^{-1}) / 2, with a second Cholesky
 * and an explicit inverse) followed by ReverseDiff back through eval_cov; here
 *   dLML/dtheta = 1/2 sum_ik (alpha alpha' - K^{-1})_ik dK_ik/dtheta
 * comes from ONE identity-augmented factorisation (the rows [I 0] appended below K solve to L^{-T}
 * and leave -K^{-1} and -alpha behind: a blocked trtri + lauum on the same work queue) and a
 * reverse-mode walk of the kernel program per covariance entry; no n x n gradient leaves the GPU.
 * Gradients are with respect to the parameters as passed (amplitudes, lengthscales, ...); the
 * caller applies the chain rule of Model.transform_param.  info_out[p] != 0: lml and gradients NaN.
 * Kernels of up to 64 nodes and 64 parameters take the tuned path; a batch with a larger kernel (structure
 * learning with max_depth = -1 proposes them now and then) takes a slower general variant — any number of
 * parameters, up to about 250 nodes (512 tape levels: SE 1, GammaExponential 2, Periodic 3, Times 2,
 * ChangePoint 4); beyond that AGP_ERR_PROGRAM.  Replaces the resident batch of the handle. */
int agp_lml_grad_batch(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops,
                       const int32_t* param_off, const int32_t* n_params, const double* params,
                       const double* noise, const double* ts, const double* xs, int32_t n,
                       double* lml_out, double* grad_params_out, double* grad_noise_out,
                       int32_t* info_out);

/* The same call when only dLML/dnoise is wanted — the noise move of rejuvenate_particle_parameters
 * (`Gen.hmc(trace, Gen.select(:noise); L, eps)`, src/inference_smc_anneal_data.jl:65-67), half of all HMC
 * gradient evaluations of a run.  dK/dnoise = I, so
 *   dLML/dnoise = 1/2 tr(alpha alpha' - K^{-1}) = 1/2 (|alpha|^2 - |L^{-1}|_F^2):
 * factorisation + blocked trtri (2 n^3/3 flops), no lauum pass and no kernel-program walk — about 0.6 of the
 * time of agp_lml_grad_batch.  Same arguments and conventions otherwise (no program-size limit).
 * The value agrees with agp_lml_grad_batch's grad_noise_out to rounding (different summation order). */
int agp_lml_grad_noise_batch(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops,
                             const int32_t* param_off, const int32_t* n_params, const double* params,
                             const double* noise, const double* ts, const double* xs, int32_t n,
                             double* lml_out, double* grad_noise_out, int32_t* info_out);

/* ---- site 3: predictive distribution ---------------------------------------------------------- */

/* For each particle p: the conditional multivariate normal  X(ts_pred) | X(ts) = xs  of
 * Distributions.MvNormal(node, noise, ts, xs, ts_pred; noise_pred) (src/GP.jl:731-758, zero mean
 * function), used by predict / predict_mvn / predict_proba (src/api.jl:497-522, 633-699):
 *   mean_out[p*m + a]        = K_21 K_11^{-1} xs
 *   cov_out[p*m*m + a*m + b] = K_22 - K_21 K_11^{-1} K_12 + noise_pred[p] * (a == b)   (symmetric)
 * noise_pred may be NULL (= noise, the reference's default).  One factorisation serves both
 * solves (the reference factorises twice, src/GP.jl:753-754): the prediction points are appended
 * to the matrix as extra tile rows, the panel solves leave L_21 = K_21 L_11^{-T}, the forward solve
 * leaves -mean, and the trailing tiles receive the Schur complement.  info_out[p]: LAPACK code of
 * the training block (0 = ok).  Replaces the resident LML batch of the handle. */
int agp_predict_batch(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops,
                      const int32_t* param_off, const int32_t* n_params, const double* params,
                      const double* noise, const double* ts, const double* xs, int32_t n,
                      const double* ts_pred, int32_t m, const double* noise_pred, double* mean_out,
                      double* cov_out, int32_t* info_out);

/* The same conditional distribution reduced to its marginals: mean_out[P][m] and var_out[P][m] = diag(cov) — what
 * `predict` reads (src/api.jl:633-699: `Distributions.quantile(dist, p)`, src/GP.jl:1006-1012, uses mean and
 * sqrt(diag(cov)) only).  Only the diagonal
 * tiles of the Schur complement are formed and m values per particle leave the GPU instead of m^2 (the m x m
 * covariances dominate agp_predict_batch's end-to-end time from m ~ 256 on).  Bitwise equal to the diagonal of
 * agp_predict_batch's cov_out. */
int agp_predict_marginals_batch(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops,
                                const int32_t* param_off, const int32_t* n_params, const double* params,
                                const double* noise, const double* ts, const double* xs, int32_t n,
                                const double* ts_pred, int32_t m, const double* noise_pred, double* mean_out,
                                double* var_out, int32_t* info_out);

/* Joint posterior of the M summands of a sum kernel and of the observable at the prediction points —
 * `infer_gp_sum(nodes, noise, ts, xs, ts_pred; noise_pred)` (src/GP.jl:904-993), the decomposition of a
 * forecast into its additive components.  The observations follow k_1 + ... + k_M + noise I; summand c of
 * particle p is program p*M + c of the wire arrays (prog_len[P*M], n_params[P*M]; param_off relative to
 * the summand's own parameter slice; the slices of one particle are contiguous in params).  Outputs per
 * particle, d = (M+1) m:  mean_out[d] and cov_out[d][d] of  [F_1(T*); ...; F_M(T*); X(T*)]  given X(T) = xs:
 *   mu = S_ab K^{-1} xs,   Sigma = S_aa - S_ab K^{-1} S_ba            (:982-984)
 * with noise_pred (NULL: noise) on the X(T*) diagonal only (:967).  The caller adds the JITTER of :986 (GP.JITTER = 1e-8, :760) when it
 * builds the MvNormal.  One factorisation: the (M+1) m rows ride along as appended tile rows, like
 * agp_predict_batch.  M must be the same for all particles of a call.  Replaces the resident batch. */
int agp_predict_sum_batch(agp_handle* h, int32_t P, int32_t M, const int32_t* prog_len, const int32_t* ops,
                          const int32_t* param_off, const int32_t* n_params, const double* params,
                          const double* noise, const double* ts, const double* xs, int32_t n,
                          const double* ts_pred, int32_t m, const double* noise_pred, double* mean_out,
                          double* cov_out, int32_t* info_out);

/* ---- plumbing -------------------------------------------------------------------------- */

/* cudaStream_t of the handle (as void*), for event timing / stream ordering by the caller. */
void* agp_stream(agp_handle* h);
int agp_synchronize(agp_handle* h);
/* Number of kernels this handle has launched since creation (bench.py "gpu_launches"). */
int64_t agp_launch_count(const agp_handle* h);
/* Time `reps` back-to-back agp_lml_run() calls with CUDA events on the handle's stream;
 * returns total milliseconds in *ms_out. */
int agp_lml_time(agp_handle* h, int32_t reps, float* ms_out);
/* Per-kernel device time of ONE run (ms), CUDA events on the handle's stream around each launch:
 * {agp_gramfill_kernel, agp_chol_kernel, 0} — for the roofline line in bench.py.  stage_ms must
 * hold 3 floats. */
int agp_lml_stage_times(agp_handle* h, float* stage_ms);

/* (order >= 100: the identity-augmented schedule of agp_lml_grad_batch built on order - 100;
 *  order >= 200: that of agp_lml_grad_noise_batch — factorisation + trtri only — built on order - 200.)
 * The in-order work queue the persistent kernel executes for P particles x nt block columns
 * (host-only, no GPU needed): items_out receives up to `cap` items as 8 int32 each
 * {type | half << 8 | store_only << 9, particle, block column k, tile row i, j0, j1, extra_flag,
 * extra_need}, type 0 = DIAG, 1 = POTF2, 2 = PANEL; [j0, j1) = contraction range in block columns; extra_flag
 * (index into the counter array laid out for nt_stride = nt) / extra_need = the counter a
 * continuation item waits for; POTF2: extra_need = number of DIAG items of its tile.  Returns the
 * total item count.  tests/test_abi_host.py replays the queue against the kernel's wait rules and
 * checks that no item ever waits for a later one (the scheduler's deadlock-freedom argument). */
int64_t agp_queue_build(int32_t P, int32_t nt, int32_t order, int32_t* items_out, int64_t cap);
/* The schedule agp_lml_run uses for a plain batch: agp_queue_build's items plus one ITEM_GRAM item per lower tile half
 * (the kernel-tree evaluation of src/GP.jl:666-668 as work items of the same queue), each `lead` items ahead of the
 * first item that reads its tile half; those readers carry the unit's flag as a wait.  `order` as above (0..3). */
int64_t agp_queue_build_gram(int32_t P, int32_t nt, int32_t order, int32_t lead, int32_t* items_out, int64_t cap);
/* Diagnostics: whether agp_lml_run evaluates the Gram matrix of the resident batch as queue items (AGP_FUSE_GRAM: -1 = by
 * size, the default: up to 5 block columns; 0 = never; 1 = whenever possible) and with which lead (AGP_GRAM_LEAD,
 * default: the number of resident CTAs).  Without a resident batch *fused_out receives the AGP_FUSE_GRAM setting. */
int agp_gram_items(agp_handle* h, int32_t* fused_out, int32_t* lead_out);
/* Same for the continuation schedules: tile rows >= first_row only (agp_lml_run_append) and
 * nt_total - nt tile rows of prediction points below the factored block (agp_predict_batch). */
int64_t agp_queue_build_general(int32_t P, int32_t nt, int32_t nt_total, int32_t first_row,
                                int32_t* items_out, int64_t cap);
/* ... and of agp_predict_marginals_batch (first_row = 0, only the diagonal tiles of the trailing block). */
int64_t agp_queue_build_marginals(int32_t P, int32_t nt, int32_t nt_total, int32_t* items_out, int64_t cap);

/* Hybrid factorisation of plain LML runs and of the gradient calls (csrc/agp_ozaki.cu).  From `min_nt` block columns on
 * (default 14: n >= 1665) the block columns are grouped into super-columns of `width` (0 = by size, the default: 4, and 3
 * from 48 block columns on); before a super-column is factored, the contraction
 * of its tiles over ALL earlier block columns — the long sums of dpotrf's trailing update, which is what PDMats runs for
 * the mvnormal of src/Model.jl:136 — is computed by exact int8 digit-plane products on the 5th-generation tensor cores
 * (tcgen05.mma kind::i8, int32 accumulators in TMEM; tcgen05 has no f64 kind), the persistent FP64 kernel then does the
 * contractions inside the super-column, POTF2 and the panel solves.  Rows are scaled by a power of two >= sqrt(K_ii)
 * (an a-priori bound on |L_ij|) and cut into seven signed 8-bit digits; products below 2^-62 of the row scales are
 * dropped: the contraction error is of the order of FP64 accumulation (measured: tests/test_gpu_parity.py).
 * mode: -1 = by size (default), 0 = never (every run takes the single-launch FP64 schedule), 1 = whenever the batch has
 * more than `width` block columns.  Environment: AGP_OZAKI, AGP_OZ_W, AGP_OZ_MIN_NT. */
int agp_set_hybrid(agp_handle* h, int32_t mode, int32_t width, int32_t min_nt);
/* Diagnostics: whether agp_lml_run takes the hybrid schedule for the resident batch (without one: the mode), the
 * super-column width, and the per-stage device times (ms) of the last agp_lml_stage_times call on a hybrid run:
 * {Gram fill + row scales, persistent-kernel segments, int8 update launches, digit-plane launches}.  Any pointer may be NULL. */
int agp_hybrid_info(agp_handle* h, int32_t* active_out, int32_t* width_out, float* stage_ms4);
/* The hybrid schedule as agp_queue_build exports it, plus the first item of every super-column's segment (seg_out
 * receives up to seg_cap entries: one per segment and the total item count).  Contractions over block columns below a
 * segment's first one are the int8 kernel's, the items start at j0 = that column.  gram_lead > 0: with the Gram units as
 * ITEM_GRAM items (segment 0: its own tiles `gram_lead` items ahead of their readers, all later diagonal tiles and
 * super-column 1; segment s: super-column s + 1), 0: the Gram matrix comes from a launch of its own.  augmented != 0: the
 * schedule of the gradient calls (counters laid out for 2 nt tile rows): the panels of the appended rows nt + a ride in the
 * bulk of their block column with contraction ranges starting at max(a, first block column of the segment); the lauum pass
 * has no items (one int8 launch after the last segment).  augmented & 2: with the ITEM_SLICE items (type 4: the int8
 * digit planes of a finished tile half, field 5 >> 16 = the rowdone value of its tile row it waits for). */
int64_t agp_queue_build_hybrid(int32_t P, int32_t nt, int32_t width, int32_t gram_lead, int32_t augmented, int32_t* items_out, int64_t cap,
                               int32_t* seg_out, int32_t seg_cap);

/* Experiment behind DESIGN.md's co-residency note: the single-launch FP64 step of the resident batch with `ctas_per_sm`
 * CTAs per SM, `reps` int8 update launches over block columns [c0, c0 + 4) on a second stream (variant 2: one CTA per unit,
 * 3: CTA pairs), alone and at the same time.
 * ms_out[4] = {FP64 alone, int8 alone, FP64 next to int8, int8 next to FP64}. */
int agp_dev_overlap_probe(agp_handle* h, int32_t ctas_per_sm, int32_t variant, int32_t c0, int32_t reps, float* ms_out);

/* Diagnostics: one traced run of the resident batch.  trace_out receives 8 int64 per work item
 * (queue order): globaltimer ns at {pop, producers ready, contraction done, Gram done, L_kk
 * ready, item done}, then SM id and CTA id.  Returns the item count (trace_out may be NULL to
 * query it).  tools/trace_report.py turns this into per-phase / per-SM utilisation. */
int64_t agp_lml_trace(agp_handle* h, int64_t* trace_out, int64_t cap_items);

/* Diagnostics: copy the factor of one particle of the resident batch to the host: ld x ld doubles, row-major, lower
 * triangle = L (== the column-major upper factor U = L' that cholesky(Symmetric(K)) returns in Julia, which is what
 * PDMats keeps inside the MvNormal of src/Model.jl:136); rows/columns >= n are padding.  *ld_out receives ld
 * (factor_out may be NULL to query it).  Valid after agp_lml_run / agp_lml_batch. */
int agp_lml_copy_factor(agp_handle* h, int32_t particle, double* factor_out, int32_t* ld_out);

#ifdef __cplusplus
}
#endif
#endif /* AGP_B200_H */
