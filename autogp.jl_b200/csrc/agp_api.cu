// C-ABI of the engine (include/agp_b200.h): handle, workspaces, program upload, launches.
// No torch types, no global mutable state; every entry point selects the handle's device
// first because the reference calls this path from migrating Julia threads
// (src/inference_smc_anneal_data.jl:133, 240).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/agp_b200.h"
#include <chrono>
#include "agp_kernels.cuh"
#include "agp_ozaki.cuh"
#include "agp_program.h"

using agp::BatchView;
using agp::SchedView;
using agp::TB;

struct agp_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;

    // resident batch
    bool uploaded = false;
    int P = 0, n_full = 0, n_active = 0, ld = 0;
    int n_pred = 0;          // prediction points appended to the resident batch (agp_predict_batch)
    int n_factored = -1;     // observations covered by the factor resident in d_L (-1: none), for agp_lml_run_append
    bool factor_clean = false;  // the last fetch saw info == 0 for every particle
    double* d_pred = nullptr; size_t cap_pred = 0;  // predictive means + covariances
    unsigned char* d_comp = nullptr; size_t cap_comp = 0;  // summand programs of agp_predict_sum_batch
    agp::ComponentView comp{};   // comp.M > 0: rewrite the appended rows after the Gram fill (agp_predict_sum_batch)
    bool aug_identity = false;  // the resident batch is identity-augmented (agp_lml_grad_batch)
    bool trtri_only = false;    // ... and only L^{-T} is wanted, not -K^{-1} (agp_lml_grad_noise_batch)
    bool pred_diag_only = false;  // predictive marginals: only the diagonal tiles of the Schur complement (agp_predict_marginals_batch)
    double* d_grad = nullptr; size_t cap_grad = 0;  // per-CTA partial sums + gradients
    const int* d_param_prefix = nullptr;            // [P+1] prefix sums of n_params (inside the input arena)
    BatchView view{};
    agp::TmaMaps tma{};          // TMA descriptors over d_L for the current (pointer, ld, P)
    double* tma_L = nullptr; int tma_ld = 0; long long tma_rows = 0;

    // device workspaces (grow-only)
    double* d_L = nullptr;       size_t cap_L = 0;       // bytes
    unsigned char* d_in = nullptr;   size_t cap_in = 0;  // packed inputs arena
    unsigned char* d_work = nullptr; size_t cap_work = 0;  // y, z, logdet, zz, dinv
    unsigned char* d_res = nullptr;  size_t cap_res = 0;   // lml[P] + info[P]
    unsigned char* h_in = nullptr;   size_t cap_hin = 0;   // pinned staging
    unsigned char* h_res = nullptr;  size_t cap_hres = 0;  // pinned results
    // gram scratch (separate from the resident LML batch so the two paths do not disturb each other)
    double* d_K = nullptr;       size_t cap_K = 0;
    unsigned char* d_gin = nullptr;  size_t cap_gin = 0;
    unsigned char* h_gin = nullptr;  size_t cap_hgin = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // persistent dataflow kernel: work queues per batch shape + dependency counters
    int order = 3;        // queue order variant (AGP_ORDER)
    int ctas_per_sm = 0;  // AGP_CTAS_PER_SM forces 1 or 2 CTAs of the persistent kernel per SM; 0 = by batch shape (chol_ctas)
    // Plain LML runs evaluate the Gram matrix either by a launch of its own in front of the persistent kernel or as GRAM
    // items of its queue.  Measured (64 particles, profiles/r02_gram_items.txt): items win 5-8 % of the step up to 5 block
    // columns (n <= 640: the factorisation is a dependency chain there and idle CTAs pick the units up), lose 3-7 % from 8
    // block columns on (the FP64 pipe is shared: a unit next to a DMMA main loop runs 2.3x slower than next to another unit).
    int fuse_gram = -1;   // AGP_FUSE_GRAM: -1 = by size (items up to kGramItemsMaxNt block columns), 0 = own launch, 1 = items
    static constexpr int kGramItemsMaxNt = 5;
    int gram_lead = 0;    // AGP_GRAM_LEAD: a GRAM item sits this many items ahead of the first reader of its tile half (0: the resident CTAs)
    unsigned long long wait_timeout_ns = 2000000000ull;  // AGP_WAIT_TIMEOUT_MS (raise under profilers that replay slowly)
    int num_sms = 0;
    // Hybrid factorisation (agp_ozaki.cu): the block columns are grouped into super-columns of `oz_width`; the contraction
    // of a super-column's tiles over ALL earlier block columns runs as exact int8 digit-plane products on tcgen05
    // (one launch per super-column), the persistent DMMA kernel then factors the super-column (contractions inside it
    // only).  Plain LML runs from `oz_min_nt` block columns on.
    int oz_mode = -1;      // AGP_OZAKI: -1 = by size, 0 = never, 1 = whenever the batch is a plain LML run with >= 2 super-columns
    int oz_width = 0;      // AGP_OZ_W (0 = by size)
    int oz_min_nt = 12;    // AGP_OZ_MIN_NT (measured, 64 particles, FP64 single launch -> hybrid, W = 4: n = 1280 2.36 -> 2.43 ms, 1408 2.93 -> 2.87,
                           // 1536 3.57 -> 3.44, 1664 4.35 -> 4.09, 2048 7.30 -> 6.17)
    int oz_min_nt_aug = 8;   // AGP_OZ_MIN_NT_AUG: the gradient calls gain earlier (three passes of contractions, the lauum pass all int8): measured
                             // n = 1024 4.51 -> 4.22 ms, 1280 7.48 -> 6.84, 1536 11.49 -> 9.69, 2048 23.6 -> 17.3 (64 particles)
    double min_noise = 0.0;  // smallest noise of the resident batch (NaN counts as -1)
    // The appended rows of an identity-augmented batch are scaled by the a-priori bound 1 / sqrt(noise) (agp_ozaki.cuh); a noise
    // far below the smallest eigenvalue the kernel itself provides (a WhiteNoise node with noise ~ 0) makes that bound loose and
    // costs digits: below this value the gradient calls keep the FP64 schedule.  The reference never goes below its JITTER = 1e-5.
    static constexpr double kAugMinNoise = 1e-9;
    bool oz_aug = true;    // AGP_OZ_AUG: the identity-augmented batches of the gradient calls take the hybrid schedule too
    // AGP_OZ_SLICE_ITEMS=1: digit planes cut by SLICE items of the segments' queues instead of launches between them.  Built,
    // replay- and GPU-tested, measured and NOT the default: n = 2048 x 64 6.80 against 6.71 ms, n = 8192 203.9 against 201.9 ms,
    // gradient call 19.8 against 19.3 ms — the segments grow by more than the launches cost (the items' HBM traffic and
    // dependency polls sit next to the panel items the segment is bound by).
    bool oz_slice_items = false;
    bool oz_ride = false;  // AGP_OZ_RIDE: Gram units as items of the segments' queues (measured slower, see run_hybrid)
    // AGP_OZ_KERNEL: -1 = by size, 3 = CTA pairs (cta_group::2, 128-column accumulators, two passes), 2 = the same per CTA.  By size (measured, 64 particles): one CTA per unit below 24 block columns (n = 2048: 6.56 against 6.69 ms per step, gradient
    // call 18.7 against 19.1 ms — twice as many units in flight, no idle half pair on odd row counts), pairs from there on (n = 4096: 32.2
    // against 32.7 ms, n = 8192: int8 updates 128.5 against 135.8 ms — the shared-memory operand traffic of the long contractions)
    int oz_variant = -1;
    bool force_plain = false;  // diagnostics that index the single-launch schedule (agp_lml_trace)
    bool oz_nomem = false;     // the digit planes did not fit into device memory once: this handle stays on the FP64 schedule
    int8_t* d_S = nullptr;     size_t cap_S = 0;       // digit planes [8][P][ld][ld]
    double* d_rscale = nullptr; size_t cap_rscale = 0;  // [P][ld][2]
    agp::OzakiMaps ozmaps{};
    int8_t* oz_S = nullptr; int oz_ld = 0, oz_P = 0;
    float hybrid_ms[4] = {0.f, 0.f, 0.f, 0.f};  // last agp_lml_stage_times of a hybrid run: Gram, DMMA segments, int8 updates, digit planes
    struct Queue {
        int4* d_items = nullptr;
        size_t cap_bytes = 0;  // size of the device buffer (a power of two: evicted buffers are recycled, see queue_buffer)
        int n_items = 0;
        unsigned long long last_use = 0;
        std::vector<int> seg;  // hybrid schedule: first item of every super-column's segment (+ the total at the end)
    };
    unsigned long long queue_clock = 0;
    static constexpr size_t kMaxQueues = 256;  // least recently used queues beyond this are evicted (lock-step loops shrink
                                               // the batch one particle at a time, data annealing walks through every nt)
    // Buffers of evicted queues, kept for the next miss: cudaFree / cudaMalloc in the middle of an inference loop were
    // measured at 50 - 600 ms a piece on a handle that holds gigabytes (profiles/r02_fit_time.txt), so the steady state of
    // a loop allocates nothing.
    std::vector<std::pair<int4*, size_t>> spare_queues;
    static constexpr size_t kMaxSpareQueues = 32;
    // agp_reserve carves one slab into spare buffers up front (cudaMalloc alone was measured at 20 - 60 ms per queue miss);
    // chunks of the slab circulate between the cache and the spares and are freed with the slab
    unsigned char* queue_slab = nullptr;
    size_t queue_slab_bytes = 0;
    bool in_slab(const void* p) const {
        return queue_slab && (const unsigned char*)p >= queue_slab && (const unsigned char*)p < queue_slab + queue_slab_bytes;
    }
    std::map<std::tuple<int, int, int, int, int>, Queue> queues;  // (P, nt, nt_total, first row tile, nt_stride)
    int* d_sync = nullptr;   size_t cap_sync = 0;
    int* h_sync = nullptr;   // pinned, 2 ints: queue head, error flag
};

namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int fail(agp_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

#define AGP_CUDA(h, call)                                                                        \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            return fail(h, (e__ == cudaErrorMemoryAllocation) ? AGP_ERR_NOMEM : AGP_ERR_CUDA,    \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                    \
        }                                                                                        \
    } while (0)

// AGP_STALL_MS=<ms>: report (stderr) every instrumented section that takes longer — developer diagnostics
struct StallTimer {
    const char* what;
    size_t arg;
    std::chrono::steady_clock::time_point t0;
    static double limit_ms() {
        static const double v = [] { const char* e = getenv("AGP_STALL_MS"); return e ? atof(e) : -1.0; }();
        return v;
    }
    StallTimer(const char* w, size_t a = 0) : what(w), arg(a), t0(std::chrono::steady_clock::now()) {}
    ~StallTimer() {
        if (limit_ms() < 0) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > limit_ms()) fprintf(stderr, "[agp stall] %s(%zu): %.1f ms\n", what, arg, ms);
    }
};

template <typename T>
int grow_device(agp_handle* h, T** ptr, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return AGP_OK;
    StallTimer stall_("grow_device", bytes);
    // small buffers (counters, partial sums, staging arenas) at least double: a loop whose batches and series grow step by
    // step would otherwise free and allocate on every step; the big ones (factors, digit planes) are sized exactly
    if (bytes < ((size_t)256 << 20)) bytes = std::max(bytes, std::min<size_t>(2 * *cap, (size_t)256 << 20));
    if (*ptr) {
        AGP_CUDA(h, cudaStreamSynchronize(h->stream));
        AGP_CUDA(h, cudaFree(*ptr));
        *ptr = nullptr;
        *cap = 0;
    }
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(ptr), bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        return fail(h, AGP_ERR_NOMEM, "device allocation of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    }
    *cap = bytes;
    return AGP_OK;
}

int grow_pinned(agp_handle* h, unsigned char** ptr, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return AGP_OK;
    StallTimer stall_("grow_pinned", bytes);
    bytes = std::max(bytes, 2 * *cap);
    if (*ptr) {
        AGP_CUDA(h, cudaStreamSynchronize(h->stream));
        AGP_CUDA(h, cudaFreeHost(*ptr));
        *ptr = nullptr;
        *cap = 0;
    }
    AGP_CUDA(h, cudaMallocHost(reinterpret_cast<void**>(ptr), bytes));
    *cap = bytes;
    return AGP_OK;
}

int check_launch(agp_handle* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, AGP_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
    return AGP_OK;
}

}  // namespace

extern "C" {

int64_t agp_queue_build(int32_t P, int32_t nt, int32_t order, int32_t* items_out, int64_t cap);

const char* agp_version(void) { return "0.1.0+sm_100a"; }

int agp_create(int device, agp_handle** out) {
    if (!out) return AGP_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return (e != cudaSuccess) ? AGP_ERR_CUDA : AGP_ERR_ARG;
    }
    agp_handle* h = new agp_handle();
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess || agp::configure_fused() != cudaSuccess || agp::configure_fused_solo() != cudaSuccess ||
        cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&h->h_sync), 2 * sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        delete h;
        return AGP_ERR_CUDA;
    }
    if (const char* e = getenv("AGP_ORDER")) h->order = atoi(e);
    if (const char* e = getenv("AGP_WAIT_TIMEOUT_MS")) h->wait_timeout_ns = 1000000ull * (unsigned long long)atoll(e);
    if (const char* e = getenv("AGP_CTAS_PER_SM")) h->ctas_per_sm = atoi(e) >= 1 ? atoi(e) : 1;
    if (const char* e = getenv("AGP_FUSE_GRAM")) h->fuse_gram = atoi(e) < 0 ? -1 : atoi(e) != 0;
    if (const char* e = getenv("AGP_GRAM_LEAD")) h->gram_lead = atoi(e) > 0 ? atoi(e) : 0;
    if (const char* e = getenv("AGP_OZAKI")) h->oz_mode = atoi(e) < 0 ? -1 : atoi(e) != 0;
    if (const char* e = getenv("AGP_OZ_W")) h->oz_width = std::max(0, atoi(e));
    if (const char* e = getenv("AGP_OZ_MIN_NT")) h->oz_min_nt = std::max(2, atoi(e));
    if (const char* e = getenv("AGP_OZ_MIN_NT_AUG")) h->oz_min_nt_aug = std::max(2, atoi(e));
    if (const char* e = getenv("AGP_OZ_RIDE")) h->oz_ride = atoi(e) != 0;
    if (const char* e = getenv("AGP_OZ_AUG")) h->oz_aug = atoi(e) != 0;
    if (const char* e = getenv("AGP_OZ_SLICE_ITEMS")) h->oz_slice_items = atoi(e) != 0;
    if (const char* e = getenv("AGP_OZ_KERNEL")) h->oz_variant = (atoi(e) == 2 || atoi(e) == 3) ? atoi(e) : -1;
    if (agp::configure_ozaki() != cudaSuccess) {
        cudaGetLastError();
        agp_destroy(h);
        return AGP_ERR_CUDA;
    }
    *out = h;
    return AGP_OK;
}

void agp_destroy(agp_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_L);
    cudaFree(h->d_in);
    cudaFree(h->d_work);
    cudaFree(h->d_res);
    cudaFree(h->d_K);
    cudaFree(h->d_gin);
    cudaFree(h->d_sync);
    cudaFree(h->d_pred);
    cudaFree(h->d_comp);
    cudaFree(h->d_grad);
    cudaFree(h->d_S);
    cudaFree(h->d_rscale);
    for (auto& kv : h->queues)
        if (!h->in_slab(kv.second.d_items)) cudaFree(kv.second.d_items);
    for (auto& sp : h->spare_queues)
        if (!h->in_slab(sp.first)) cudaFree(sp.first);
    cudaFree(h->queue_slab);
    cudaFreeHost(h->h_sync);
    cudaFreeHost(h->h_gin);
    cudaFreeHost(h->h_in);
    cudaFreeHost(h->h_res);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* agp_last_error(const agp_handle* h) { return h ? h->err.c_str() : "null handle"; }

void* agp_stream(agp_handle* h) { return h ? (void*)h->stream : nullptr; }

int agp_synchronize(agp_handle* h) {
    if (!h) return AGP_ERR_ARG;
    AGP_CUDA(h, cudaSetDevice(h->device));
    AGP_CUDA(h, cudaStreamSynchronize(h->stream));
    return AGP_OK;
}

int64_t agp_launch_count(const agp_handle* h) { return h ? h->launches : 0; }

// ---- Gram ---------------------------------------------------------------------------------

static int gram_impl(agp_handle* h, const int32_t* ops, const int32_t* param_off, int32_t m, const double* params, int32_t n_params,
                     const double* ts, int32_t n, double noise, int32_t form, double* K_dev) {
    std::vector<AgpInstr> instr;
    int need = 1;
    std::string err;
    int rc = agp_compile_program(ops, param_off, m, params, n_params, instr, &need, err);
    if (rc != AGP_OK) return fail(h, rc, err);
    size_t prog_bytes = instr.size() * sizeof(AgpInstr);
    size_t ts_off = align_up(prog_bytes, 16);
    size_t total = ts_off + (size_t)n * 8;
    if ((rc = grow_pinned(h, &h->h_gin, &h->cap_hgin, total)) != AGP_OK) return rc;
    if ((rc = grow_device(h, &h->d_gin, &h->cap_gin, total)) != AGP_OK) return rc;
    if (!K_dev && (rc = grow_device(h, &h->d_K, &h->cap_K, (size_t)n * n * 8)) != AGP_OK) return rc;
    // staging may still be in flight from a previous async call on this handle
    AGP_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(h->h_gin, instr.data(), prog_bytes);
    memcpy(h->h_gin + ts_off, ts, (size_t)n * 8);
    unsigned char* d_args = h->d_gin;
    AGP_CUDA(h, cudaMemcpyAsync(d_args, h->h_gin, total, cudaMemcpyHostToDevice, h->stream));
    double* K_target = K_dev ? K_dev : h->d_K;
    agp::launch_gram(reinterpret_cast<const AgpInstr*>(d_args), (int)instr.size(), need, reinterpret_cast<const double*>(d_args + ts_off), n, noise,
                     form, K_target, h->stream);
    h->launches += 1;
    return check_launch(h, "gram");
}

int agp_gram(agp_handle* h, const int32_t* ops, const int32_t* param_off, int32_t m, const double* params, int32_t n_params, const double* ts,
             int32_t n, double noise, int32_t form, double* K_out) {
    if (!h) return AGP_ERR_ARG;
    if (!ops || !param_off || (!params && n_params > 0) || n < 0 || (n > 0 && (!ts || !K_out)) || (form != 0 && form != 1))
        return fail(h, AGP_ERR_ARG, "agp_gram: bad argument");
    AGP_CUDA(h, cudaSetDevice(h->device));
    if (n == 0) {
        // still validate the program, like the reference would construct the Node
        std::vector<AgpInstr> instr; int need; std::string err;
        int rc = agp_compile_program(ops, param_off, m, params, n_params, instr, &need, err);
        return rc == AGP_OK ? AGP_OK : fail(h, rc, err);
    }
    int rc = gram_impl(h, ops, param_off, m, params, n_params, ts, n, noise, form, nullptr);
    if (rc != AGP_OK) return rc;
    AGP_CUDA(h, cudaMemcpyAsync(K_out, h->d_K, (size_t)n * n * 8, cudaMemcpyDeviceToHost, h->stream));
    AGP_CUDA(h, cudaStreamSynchronize(h->stream));
    return AGP_OK;
}

int agp_gram_device(agp_handle* h, const int32_t* ops, const int32_t* param_off, int32_t m, const double* params, int32_t n_params,
                    const double* ts, int32_t n, double noise, int32_t form, double* K_out_dev) {
    if (!h) return AGP_ERR_ARG;
    if (!ops || !param_off || (!params && n_params > 0) || n <= 0 || !ts || !K_out_dev || (form != 0 && form != 1))
        return fail(h, AGP_ERR_ARG, "agp_gram_device: bad argument");
    AGP_CUDA(h, cudaSetDevice(h->device));
    return gram_impl(h, ops, param_off, m, params, n_params, ts, n, noise, form, K_out_dev);
}

// ---- LML batch ------------------------------------------------------------------------------

// Upload of a batch; with m > 0 the prediction points ts_pred are appended as extra rows that start
// at the next tile boundary after the observations (agp_predict_batch).
static int upload_impl(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops, const int32_t* param_off, const int32_t* n_params,
                       const double* params, const double* noise, const double* ts, const double* xs, int32_t n, const double* ts_pred, int32_t m,
                       const double* noise_pred, bool aug_identity = false) {
    if (!h) return AGP_ERR_ARG;
    h->uploaded = false;
    h->n_factored = -1;
    h->factor_clean = false;
    h->comp.M = 0;
    h->pred_diag_only = false;
    if (P < 0 || n < 0 || m < 0 || (P > 0 && (!prog_len || !ops || !param_off || !n_params || !noise)) || (n > 0 && (!ts || !xs)) || (m > 0 && !ts_pred && !aug_identity))
        return fail(h, AGP_ERR_ARG, "agp_lml_upload: bad argument");
    if (P > 65535) return fail(h, AGP_ERR_ARG, "agp_lml_upload: at most 65535 particles per batch");
    AGP_CUDA(h, cudaSetDevice(h->device));

    // compile programs
    std::vector<AgpInstr> instr;
    std::vector<int32_t> poff(P + 1, 0), pneed(P > 0 ? P : 1, 1);
    {
        size_t o = 0, po = 0;
        std::string err;
        for (int p = 0; p < P; ++p) {
            if (prog_len[p] <= 0 || n_params[p] < 0) return fail(h, AGP_ERR_PROGRAM, "particle " + std::to_string(p) + ": empty program");
            poff[p] = (int32_t)instr.size();
            int need = 1;
            int rc = agp_compile_program(ops + o, param_off + o, prog_len[p], params ? params + po : nullptr, n_params[p], instr, &need, err);
            if (rc != AGP_OK) return fail(h, rc, "particle " + std::to_string(p) + ": " + err);
            pneed[p] = need;
            o += prog_len[p];
            po += n_params[p];
        }
        poff[P] = (int32_t)instr.size();
    }

    const int ld_obs = (int)align_up((size_t)(n > 0 ? n : (m > 0 ? 0 : 1)), TB);
    if (aug_identity) m = ld_obs;  // one appended row per (padded) observation, carrying [I 0]
    const int ld = ld_obs + (int)align_up((size_t)m, TB);
    // packed input arena: ts[ld] xs[ld] noise[P] noise_pred[P] prog_off[P+1] prog_need[P] instr[]
    size_t off_ts = 0;
    size_t off_xs = off_ts + (size_t)ld * 8;
    size_t off_noise = off_xs + (size_t)ld * 8;
    size_t off_npred = off_noise + (size_t)P * 8;
    size_t off_poff = align_up(off_npred + (size_t)P * 8, 16);
    size_t off_need = align_up(off_poff + (size_t)(P + 1) * 4, 16);
    size_t off_pprefix = align_up(off_need + (size_t)P * 4, 16);
    size_t off_instr = align_up(off_pprefix + (size_t)(P + 1) * 4, 32);
    size_t in_bytes = off_instr + instr.size() * sizeof(AgpInstr);
    int rc;
    if ((rc = grow_pinned(h, &h->h_in, &h->cap_hin, in_bytes)) != AGP_OK) return rc;
    if ((rc = grow_device(h, &h->d_in, &h->cap_in, in_bytes)) != AGP_OK) return rc;
    // work arena: y[P][ld] z[P][ld] cum[P][ld/128][2] dinv[P][ld/128][4096] (one set of diagonal-block
    // inverses per block column: the persistent kernel factors column k+1 while column k is still
    // being solved)
    size_t off_y = 0;
    size_t off_z = off_y + (size_t)P * ld * 8;
    size_t off_cum = off_z + (size_t)P * ld * 8;
    size_t off_dinv = align_up(off_cum + (size_t)P * (ld / TB) * 2 * 8, 256);
    size_t work_bytes = off_dinv + (size_t)P * (ld / TB) * 4096 * 8;
    if ((rc = grow_device(h, &h->d_work, &h->cap_work, work_bytes)) != AGP_OK) return rc;
    size_t res_bytes = align_up((size_t)P * 8, 16) + (size_t)P * 4;
    {
        const size_t cap_before = h->cap_res;
        if ((rc = grow_device(h, &h->d_res, &h->cap_res, res_bytes > 0 ? res_bytes : 16)) != AGP_OK) return rc;
        // a fresh block: the fetch copies lml[P] | pad | info[P] in one piece, the pad bytes are never written by a kernel
        if (h->cap_res != cap_before) AGP_CUDA(h, cudaMemsetAsync(h->d_res, 0, h->cap_res, h->stream));
    }
    if ((rc = grow_pinned(h, &h->h_res, &h->cap_hres, res_bytes > 0 ? res_bytes : 16)) != AGP_OK) return rc;
    size_t L_bytes = (size_t)P * ld * ld * 8;
    if ((rc = grow_device(h, &h->d_L, &h->cap_L, L_bytes > 0 ? L_bytes : 16)) != AGP_OK) return rc;

    // the pinned staging buffer may still be the source of an in-flight copy
    AGP_CUDA(h, cudaStreamSynchronize(h->stream));
    memset(h->h_in, 0, off_noise);
    if (n > 0) {
        memcpy(h->h_in + off_ts, ts, (size_t)n * 8);
        memcpy(h->h_in + off_xs, xs, (size_t)n * 8);
    }
    if (m > 0 && ts_pred) memcpy(h->h_in + off_ts + (size_t)ld_obs * 8, ts_pred, (size_t)m * 8);
    {
        int32_t* pp = reinterpret_cast<int32_t*>(h->h_in + off_pprefix);
        pp[0] = 0;
        for (int p = 0; p < P; ++p) pp[p + 1] = pp[p] + n_params[p];
    }
    h->min_noise = 1e300;
    for (int p = 0; p < P; ++p) h->min_noise = (noise[p] == noise[p] && noise[p] < h->min_noise) ? noise[p] : (noise[p] == noise[p] ? h->min_noise : -1.0);
    if (P > 0) {
        memcpy(h->h_in + off_noise, noise, (size_t)P * 8);
        memcpy(h->h_in + off_npred, noise_pred ? noise_pred : noise, (size_t)P * 8);  // default: noise_pred = noise (src/GP.jl:738)
        memcpy(h->h_in + off_need, pneed.data(), (size_t)P * 4);
    }
    memcpy(h->h_in + off_poff, poff.data(), (size_t)(P + 1) * 4);
    if (!instr.empty()) memcpy(h->h_in + off_instr, instr.data(), instr.size() * sizeof(AgpInstr));
    AGP_CUDA(h, cudaMemcpyAsync(h->d_in, h->h_in, in_bytes, cudaMemcpyHostToDevice, h->stream));

    BatchView& v = h->view;
    v.L = h->d_L;
    v.mat_stride = (long long)ld * ld;
    v.ld = ld;
    v.n = n;
    v.nt = (n + TB - 1) / TB;
    v.n_pred = m;
    v.nt_total = v.nt + (m + TB - 1) / TB;
    v.ts = reinterpret_cast<const double*>(h->d_in + off_ts);
    v.xs = reinterpret_cast<const double*>(h->d_in + off_xs);
    v.noise = reinterpret_cast<const double*>(h->d_in + off_noise);
    v.prog_off = reinterpret_cast<const int*>(h->d_in + off_poff);
    v.prog_need = reinterpret_cast<const int*>(h->d_in + off_need);
    v.prog = reinterpret_cast<const AgpInstr*>(h->d_in + off_instr);
    v.max_prog_len = 0;
    for (int p = 0; p < P; ++p) v.max_prog_len = std::max(v.max_prog_len, (int)(poff[p + 1] - poff[p]));
    v.y = reinterpret_cast<double*>(h->d_work + off_y);
    v.z = reinterpret_cast<double*>(h->d_work + off_z);
    v.aug_identity = aug_identity ? 1 : 0;
    h->aug_identity = aug_identity;
    h->d_param_prefix = reinterpret_cast<const int*>(h->d_in + off_pprefix);
    v.cum = reinterpret_cast<double*>(h->d_work + off_cum);
    v.dinv = reinterpret_cast<double*>(h->d_work + off_dinv);
    v.lml = reinterpret_cast<double*>(h->d_res);
    v.info = reinterpret_cast<int*>(h->d_res + align_up((size_t)P * 8, 16));
    h->P = P;
    h->n_full = n;
    h->n_active = n;
    h->n_pred = m;
    h->ld = ld;
    h->uploaded = true;
    return AGP_OK;
}

static size_t sync_ints(int P, int nt_stride, bool fused);

int agp_reserve(agp_handle* h, int32_t max_n, int32_t max_pred, int32_t max_batch, int32_t with_gradient) {
    if (!h || max_n < 0 || max_pred < 0 || max_batch < 0) return AGP_ERR_ARG;
    if (max_batch > 65535) return fail(h, AGP_ERR_ARG, "agp_reserve: at most 65535 particles per batch");
    AGP_CUDA(h, cudaSetDevice(h->device));
    const size_t P = (size_t)max_batch;
    const size_t ld_obs = align_up((size_t)(max_n > 0 ? max_n : 1), TB);
    const size_t ld_pred = ld_obs + align_up((size_t)max_pred, TB), ld_aug = with_gradient ? 2 * ld_obs : 0;
    const size_t ld = std::max(ld_pred, ld_aug);
    int rc;
    if ((rc = grow_device(h, &h->d_L, &h->cap_L, std::max<size_t>(16, P * ld * ld * 8))) != AGP_OK) return rc;
    // work arena of upload_impl: y, z, cum, dinv
    const size_t work = align_up(2 * P * ld * 8 + P * (ld / TB) * 16, 256) + P * (ld / TB) * 4096 * 8;
    if ((rc = grow_device(h, &h->d_work, &h->cap_work, work)) != AGP_OK) return rc;
    const int nt = (int)(ld_obs / TB);
    if (with_gradient) {  // per-CTA partial sums of agp_grad_kernel + outputs (64 parameters per kernel assumed; grows if there are more)
        const size_t blocks = (size_t)nt * (nt + 1);
        if ((rc = grow_device(h, &h->d_grad, &h->cap_grad, (P * blocks * (agp::AGP_GRAD_MAX_PARAMS + 1) + P * (agp::AGP_GRAD_MAX_PARAMS + 1) + 2) * 8)) != AGP_OK) return rc;
    }
    {   // dependency counters of the persistent kernel (incl. the flags of Gram items)
        const size_t nts = ld / TB;
        if ((rc = grow_device(h, &h->d_sync, &h->cap_sync, sync_ints((int)P, (int)nts, true) * sizeof(int))) != AGP_OK) return rc;
    }
    if (!h->queue_slab) {  // spare queue buffers: 8 x 2 MB, 16 x 1 MB, 32 x 512 KB, 32 x 256 KB, 64 x 64 KB, 64 x 16 KB = 61 MB
        const size_t sizes[6] = {(size_t)2 << 20, (size_t)1 << 20, (size_t)512 << 10, (size_t)256 << 10, (size_t)64 << 10, (size_t)16 << 10};
        const int counts[6] = {8, 16, 32, 32, 64, 64};
        size_t total = 0;
        for (int a = 0; a < 6; ++a) total += sizes[a] * counts[a];
        if (cudaMalloc(reinterpret_cast<void**>(&h->queue_slab), total) == cudaSuccess) {
            h->queue_slab_bytes = total;
            size_t off = 0;
            for (int a = 0; a < 6; ++a)
                for (int c = 0; c < counts[a]; ++c, off += sizes[a]) h->spare_queues.emplace_back(reinterpret_cast<int4*>(h->queue_slab + off), sizes[a]);
        } else {
            cudaGetLastError();
            h->queue_slab = nullptr;
        }
    }
    // digit planes and row scales of the hybrid schedule, when a call of this size would take it
    const int min_nt = with_gradient ? std::min(h->oz_min_nt, h->oz_min_nt_aug) : h->oz_min_nt;
    if (h->oz_mode != 0 && nt >= min_nt) {
        const size_t ld_h = with_gradient ? ld_aug : ld_obs;
        if ((rc = grow_device(h, &h->d_S, &h->cap_S, (size_t)agp::OZ_SLICES * P * ld_h * ld_h)) != AGP_OK) {
            cudaGetLastError();  // the hybrid schedule is optional: the calls fall back to the FP64 schedule when the planes do not fit
            h->err.clear();
        } else if ((rc = grow_device(h, &h->d_rscale, &h->cap_rscale, P * ld_h * 16)) != AGP_OK) {
            return rc;
        }
    }
    return AGP_OK;
}

int agp_lml_upload(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops, const int32_t* param_off, const int32_t* n_params,
                   const double* params, const double* noise, const double* ts, const double* xs, int32_t n) {
    return upload_impl(h, P, prog_len, ops, param_off, n_params, params, noise, ts, xs, n, nullptr, 0, nullptr);
}

int agp_lml_set_prefix(agp_handle* h, int32_t n_prefix) {
    if (!h) return AGP_ERR_ARG;
    if (!h->uploaded) return fail(h, AGP_ERR_STATE, "agp_lml_set_prefix: no resident batch");
    if (h->n_pred > 0) return fail(h, AGP_ERR_STATE, "agp_lml_set_prefix: the resident batch carries prediction points");
    if (n_prefix < 0 || n_prefix > h->n_full) return fail(h, AGP_ERR_ARG, "agp_lml_set_prefix: prefix out of range");
    h->n_active = n_prefix;
    h->view.n = n_prefix;
    h->view.nt = (n_prefix + TB - 1) / TB;
    h->view.nt_total = h->view.nt;
    return AGP_OK;
}


// ---- persistent dataflow path: the in-order work queue -------------------------------------
//
// Any order is valid as long as every item's producers come EARLIER (agp_fused.cu); with
// PANEL(p,k,i,h)[j0,j1) / DIAG(p,k,h)[j0,j1) contracting block columns j0 <= j < j1:
//   any item [j0,j1)     <- final PANEL(p,j,i,*) and PANEL(p,j,k,*) for all j < j1
//   continuation j0 > 0  <- the partial item [0,j0) of the same tile (both halves)
//   POTF2(p,k)           <- every DIAG item of tile (k,k)
//   final PANEL(p,k,i,h) <- POTF2(p,k)
// order 0: block column by block column:  DIAG | POTF2 | PANEL.
// order 1: look-ahead: the panels of tile row k+1 go first in block column k, and DIAG(k+1),
//          POTF2(k+1) are interleaved into the bulk of column k's panels.
// order 2: order 1 + split contractions: the next diagonal tile (k+1,k+1) and the panel below it
//          (k+2,k+1) get a PARTIAL item over [0,k) one block column early, so only the last 128
//          columns of their contraction remain on the per-particle critical path
//          POTF2(k) -> PANEL(k,k+1) -> DIAG(k+1) -> POTF2(k+1).
// second int4 of a work item (agp_kernels.cuh)
static int4 pack_dep(int j0, int j1, int need_k, int need_i, int flag, int need) {
    return make_int4(j0 | (j1 << 16), need_k | (need_i << 16), flag, need);
}

struct QueueLayout {
    int P, nt, nt_stride;
    int flag_diagu(int p, int k) const { return 32 + P * nt_stride + p * nt_stride + k; }
    int flag_ppre(int p, int i) const { return 32 + 2 * P * nt_stride + p * nt_stride + i; }
};

// order 3: order 2 for the early block columns (bulk panels tile row by tile row, so the rows the next
//          look-ahead items read finish first), and the last AGP_LATE block columns RIGHT-LOOKING: when the
//          switch column is reached every tile of the trailing triangle is brought up to date over [0,k) by
//          one store-only item, and from then on every block column only adds single products
//          (final panel = last product + solve; trailing tiles one product each).  Late block columns have
//          too few tiles for all CTAs: their per-particle critical path then no longer waits for
//          contractions that could have been done earlier.
struct TileState {
    int P, nt, nt_stride, nrows;  // nrows = nt, or 2 nt with the appended rows of an identity-augmented batch (tile row nt + a is
                                  // structurally zero left of block column a: its coverage starts there)
    std::vector<int> cov, n_ppre, n_diag;  // coverage per tile, partial / diag item counts
    std::vector<int4>* items;
    TileState(int P_, int nt_, int nts, std::vector<int4>* it, int nrows_ = 0)
        : P(P_), nt(nt_), nt_stride(nts), nrows(nrows_ > nt_ ? nrows_ : nt_), cov((size_t)P_ * (nrows_ > nt_ ? nrows_ : nt_) * nt_, 0),
          n_ppre((size_t)P_ * (nrows_ > nt_ ? nrows_ : nt_), 0), n_diag((size_t)P_ * nt_, 0), items(it) {
        for (int p = 0; p < P; ++p)
            for (int i = nt; i < nrows; ++i)
                for (int k = 0; k < nt; ++k) cov[((size_t)p * nrows + i) * nt + k] = i - nt;
    }
    int fc(int i) const { return i >= nt ? i - nt : 0; }
    int& coverage(int p, int i, int k) { return cov[((size_t)p * nrows + i) * nt + k]; }
    int flag_diagu(int p, int k) const { return 32 + P * nt_stride + p * nt_stride + k; }
    int flag_ppre(int p, int i) const { return 32 + 2 * P * nt_stride + p * nt_stride + i; }
    void potf2(int p, int k) {
        items->push_back(make_int4(agp::ITEM_POTF2, p, k, k));
        items->push_back(make_int4(0, 0, -1, n_diag[(size_t)p * nt + k]));
    }
    // advance tile (i,k) to coverage j1 (final when j1 == k): both row halves
    void tile(int p, int i, int k, int j1) {
        int& c = coverage(p, i, k);
        const int j0 = c, f = fc(i);
        const bool fin = j1 == k;
        for (int h = 0; h < 2; ++h) {
            if (i == k) {
                const int cnt = n_diag[(size_t)p * nt + k];
                items->push_back(make_int4(agp::ITEM_DIAG | (h << 8) | (fin ? 0 : agp::ITEM_PARTIAL), p, k, i));
                items->push_back(make_int4(j0 | (j1 << 16), 2 * j1, j0 > 0 ? flag_diagu(p, k) : -1, j0 > 0 ? cnt : 0));
            } else {
                const int cnt = n_ppre[(size_t)p * nrows + i];
                items->push_back(make_int4(agp::ITEM_PANEL | (h << 8) | (fin ? 0 : agp::ITEM_PARTIAL) | (k == f ? agp::ITEM_YINIT : 0), p, k, i));
                items->push_back(make_int4(j0 | (j1 << 16), (2 * j1) | ((2 * (j1 - f)) << 16), j0 > f ? flag_ppre(p, i) : -1, j0 > f ? cnt : 0));
            }
        }
        if (i == k) n_diag[(size_t)p * nt + k] += 2;
        else if (!fin) n_ppre[(size_t)p * nrows + i] += 2;
        c = j1;
    }
};

static void build_queue_order3(int P, int nt, int nt_stride, int late, std::vector<int4>& items) {
    items.clear();
    TileState b(P, nt, nt_stride, &items);
    int split_from = 3;
    if (const char* e = getenv("AGP_SPLIT_FROM")) split_from = std::max(2, atoi(e));  // developer A/B
    auto split = [&](int k) { return k >= split_from && k < nt; };
    const int ks = std::max(nt - late, 1);  // first block column of the right-looking phase
    double pos_la = 0.0, pos_diag = 1.0 / 3, pos_potf2 = 0.5;
    if (const char* e = getenv("AGP_POS")) sscanf(e, "%lf,%lf,%lf", &pos_la, &pos_diag, &pos_potf2);  // developer A/B
    for (int p = 0; p < P; ++p) b.tile(p, 0, 0, 0);
    for (int p = 0; p < P; ++p) b.potf2(p, 0);
    struct T { int p, i, k, j1; };
    // Particle groups: within a block column the schedule below is emitted group by group (AGP_PGROUP particles each), so
    // that the items in flight at any time belong to few particles and the B operand of a particle — tile row k, read by
    // every panel item of the column — stays in L2 between them.  With all particles in one group (the round-1 order)
    // consecutive items belong to different particles and 296 resident items touch all 64 B panels plus their own A panels.
    int pg = P;
    if (const char* e = getenv("AGP_PGROUP")) pg = std::max(1, atoi(e));
    for (int k = 0; k < nt - 1; ++k)
      for (int p0 = 0; p0 < P; p0 += pg) {
        const int p1 = std::min(P, p0 + pg);
        for (int p = p0; p < p1; ++p) b.tile(p, k + 1, k, k);  // panels of tile row k+1 first
        if (k < ks) {
            std::vector<T> la;  // look-ahead store-only items over [0,k)
            if (split(k + 1) && k >= 1)
                for (int p = p0; p < p1; ++p) {
                    la.push_back({p, k + 1, k + 1, k});
                    if (k + 2 < nt) la.push_back({p, k + 2, k + 1, k});
                }
            if (k + 1 == ks && k >= 1)  // the switch: the rest of the trailing triangle
                for (int p = p0; p < p1; ++p)
                    for (int c = k + 1; c < nt; ++c)
                        for (int i = c; i < nt; ++i)
                            if (!(split(k + 1) && c == k + 1 && (i == k + 1 || i == k + 2))) la.push_back({p, i, c, k});
            std::vector<T> bulk;
            for (int i = k + 2; i < nt; ++i)
                for (int p = p0; p < p1; ++p) bulk.push_back({p, i, k, k});
            // the look-ahead items, DIAG(k+1) and POTF2(k+1) are interleaved into the bulk of column k at these
            // fractions (tuned with tools/queue_sim.py, confirmed on the device: profiles/r01_order_tuning.txt)
            const size_t nb = bulk.size();
            const size_t c0 = (size_t)(pos_la * nb), c1 = std::max(c0, (size_t)(pos_diag * nb)), c2 = std::max(c1, (size_t)(pos_potf2 * nb));
            for (size_t a = 0; a < c0; ++a) b.tile(bulk[a].p, bulk[a].i, bulk[a].k, bulk[a].j1);
            for (const T& a : la) b.tile(a.p, a.i, a.k, a.j1);
            for (size_t a = c0; a < c1; ++a) b.tile(bulk[a].p, bulk[a].i, bulk[a].k, bulk[a].j1);
            for (int p = p0; p < p1; ++p) b.tile(p, k + 1, k + 1, k + 1);
            for (size_t a = c1; a < c2; ++a) b.tile(bulk[a].p, bulk[a].i, bulk[a].k, bulk[a].j1);
            for (int p = p0; p < p1; ++p) b.potf2(p, k + 1);
            for (size_t a = c2; a < nb; ++a) b.tile(bulk[a].p, bulk[a].i, bulk[a].k, bulk[a].j1);
        } else {
            for (int p = p0; p < p1; ++p) b.tile(p, k + 1, k + 1, k + 1);
            for (int p = p0; p < p1; ++p) b.potf2(p, k + 1);
            for (int i = k + 2; i < nt; ++i)
                for (int p = p0; p < p1; ++p) b.tile(p, i, k, k);
            for (int c = k + 2; c < nt; ++c)
                for (int i = c; i < nt; ++i)
                    for (int p = p0; p < p1; ++p) b.tile(p, i, c, k + 1);
        }
      }
}

static void build_queue(int P, int nt, int nt_stride, int order, std::vector<int4>& items) {
    if (order >= 3 && nt > 1) {
        // right-looking tail: 4 block columns from nt = 16 on, a quarter of the block columns below that (measured at
        // n = 1024, 256 particles: 4.46 ms with 2, 4.56 with 4, 4.85 with 8; n = 512 is insensitive)
        int late = std::min(4, std::max(1, nt / 4));
        if (const char* e = getenv("AGP_LATE")) late = atoi(e);
        build_queue_order3(P, nt, nt_stride, late, items);
        return;
    }
    const QueueLayout lay{P, nt, nt_stride};
    const int split_from = 3;  // block columns below this are too short to be worth splitting
    auto split = [&](int k) { return order >= 2 && k >= split_from && k < nt; };  // tile (k,k) and (k+1,k) are split
    auto push = [&](int type, int h, int p, int k, int i, int j0, int j1, int flag, int need) {
        int flags = (type != agp::ITEM_POTF2 && j1 < k) ? agp::ITEM_PARTIAL : 0;
        if (type == agp::ITEM_PANEL && k == 0) flags |= agp::ITEM_YINIT;
        items.push_back(make_int4(type | (h << 8) | flags, p, k, i));
        items.push_back(pack_dep(j0, j1, type == agp::ITEM_POTF2 ? 0 : 2 * j1, type == agp::ITEM_PANEL ? 2 * j1 : 0, flag, need));
    };
    auto diag_full = [&](int p, int k) {
        for (int h = 0; h < 2; ++h) push(agp::ITEM_DIAG, h, p, k, k, 0, k, -1, 0);
    };
    auto potf2 = [&](int p, int k) { push(agp::ITEM_POTF2, 0, p, k, k, 0, 0, -1, split(k) ? 4 : 2); };
    auto panel_full = [&](int p, int k, int i) {
        for (int h = 0; h < 2; ++h) push(agp::ITEM_PANEL, h, p, k, i, 0, k, -1, 0);
    };
    items.clear();
    if (order == 0 || nt == 1) {
        for (int k = 0; k < nt; ++k) {
            for (int p = 0; p < P; ++p) diag_full(p, k);
            for (int p = 0; p < P; ++p) potf2(p, k);
            for (int p = 0; p < P; ++p)
                for (int i = k + 1; i < nt; ++i) panel_full(p, k, i);
        }
        return;
    }
    for (int p = 0; p < P; ++p) diag_full(p, 0);
    for (int p = 0; p < P; ++p) potf2(p, 0);
    for (int k = 0; k < nt - 1; ++k) {
        // panels of tile row k+1 first: they feed the next diagonal tile
        for (int p = 0; p < P; ++p) {
            if (split(k)) {
                for (int h = 0; h < 2; ++h) push(agp::ITEM_PANEL, h, p, k, k + 1, k - 1, k, lay.flag_ppre(p, k + 1), 2);
            } else {
                panel_full(p, k, k + 1);
            }
        }
        // look-ahead partials of block column k+1 (need block columns < k only)
        if (split(k + 1)) {
            for (int p = 0; p < P; ++p) {
                for (int h = 0; h < 2; ++h) push(agp::ITEM_DIAG, h, p, k + 1, k + 1, 0, k, -1, 0);
                if (k + 2 < nt)
                    for (int h = 0; h < 2; ++h) push(agp::ITEM_PANEL, h, p, k + 1, k + 2, 0, k, -1, 0);
            }
        }
        // bulk of column k, with DIAG(k+1) after the first third and POTF2(k+1) after the second
        std::vector<int4> bulk;
        for (int p = 0; p < P; ++p)
            for (int i = k + 2; i < nt; ++i)
                for (int h = 0; h < 2; ++h) {
                    bulk.push_back(make_int4(agp::ITEM_PANEL | (h << 8) | (k == 0 ? agp::ITEM_YINIT : 0), p, k, i));
                    bulk.push_back(pack_dep(0, k, 2 * k, 2 * k, -1, 0));
                }
        size_t n_bulk = bulk.size() / 2;
        size_t c1 = 2 * (n_bulk / 3), c2 = 2 * (2 * n_bulk / 3);
        items.insert(items.end(), bulk.begin(), bulk.begin() + c1);
        for (int p = 0; p < P; ++p) {
            if (split(k + 1)) {
                for (int h = 0; h < 2; ++h) push(agp::ITEM_DIAG, h, p, k + 1, k + 1, k, k + 1, lay.flag_diagu(p, k + 1), 2);
            } else {
                diag_full(p, k + 1);
            }
        }
        items.insert(items.end(), bulk.begin() + c1, bulk.begin() + c2);
        for (int p = 0; p < P; ++p) potf2(p, k + 1);
        items.insert(items.end(), bulk.begin() + c2, bulk.end());
    }
}

// Hybrid schedule (agp_ozaki.cu): block columns in super-columns of W.  Segment s covers block columns [c0, c1) =
// [s W, min(nt, (s + 1) W)) and is ONE launch of the persistent kernel; before it (s >= 1) the int8 update kernel has
// brought every lower tile of those block columns to coverage c0 (T_ik = K_ik - sum_{j < c0} L_ij L_kj^T), so the
// segment's items contract over [c0, k) only: DIAG(k)[c0, k), POTF2(k), PANEL(k, i)[c0, k) + solve for every i > k.
// Inside a segment the order is the look-ahead order of the single-launch schedule (panels of tile row k + 1 first,
// DIAG(k + 1) / POTF2(k + 1) interleaved into the bulk of column k).  The dependency counters are NOT reset between
// the segments (rowdone / fdone / diagu keep counting), only the queue head is.
static int gram_flag(int P, int nt_stride, int p, int i, int k, int hh);
static void fuse_gram_items(int P, int nt_stride, int lead, std::vector<int4>& items);

// gram_lead > 0: the Gram units ride in the queue as well (ITEM_GRAM).  Segment 0 carries the units of its own tiles `lead`
// items ahead of their first readers (fuse_gram_items, as in the single-launch schedule), the diagonal tiles of ALL later
// super-columns (the row scales of the digit planes are read off the Gram diagonal right after segment 0) and the other
// tiles of super-column 1; segment s >= 1 carries the tiles of super-column s + 1.  Nothing in a segment reads the units
// that ride for a later super-column, so they need no flag wait — the launch boundary orders them — and they are dealt
// out behind the POTF2 items of the segment's block columns, where the CTAs that hold no POTF2 item would otherwise
// spin on the factor of the diagonal tile.
// aug: the identity-augmented batch of the gradient calls (tile rows nt + a hold [I 0] and end as rows of L^{-T}): the
// panels of the appended rows ride in the bulk of their block column, their contraction over [a, c0) is the int8 kernel's
// like everyone else's; the lauum pass (-K^{-1} into the trailing tiles) has no FP64 items at all — one int8 launch after
// the last segment.
// slice_items: the int8 digit planes of a block column's finished panels (the rows later int8 launches read: below the
// super-column, and the appended rows) are cut by SLICE items of the segment itself, one block column behind the panels.
static void build_queue_hybrid(int P, int nt, int nt_stride, int W, int gram_lead, std::vector<int4>& items, std::vector<int>& seg, bool aug = false,
                               bool slice_items = false, bool with_lauum = true) {
    items.clear();
    seg.clear();
    std::vector<int4> cur;
    TileState b(P, nt, nt_stride, &cur, aug ? 2 * nt : nt);
    double pos_diag = 1.0 / 3, pos_potf2 = 0.5;
    if (const char* e = getenv("AGP_OZ_POS")) sscanf(e, "%lf,%lf", &pos_diag, &pos_potf2);  // developer A/B
    struct T { int p, i, k; };
    auto gram_unit = [&](std::vector<int4>& out, int p, int i, int k, int hh) {
        out.push_back(make_int4(agp::ITEM_GRAM | (hh << 8), p, k, i));
        out.push_back(make_int4(0, 0, gram_flag(P, nt_stride, p, i, k, i == k ? 0 : hh), 0));
    };
    for (int c0 = 0; c0 < nt; c0 += W) {
        const int c1 = std::min(nt, c0 + W);
        cur.clear();
        std::vector<int4> ride;  // Gram units of later super-columns that ride in this segment
        if (gram_lead > 0 && c1 < nt) {
            if (c0 == 0)
                for (int c = c1; c < nt; ++c)
                    for (int p = 0; p < P; ++p)
                        for (int hh = 0; hh < 2; ++hh) gram_unit(ride, p, c, c, hh);
            for (int k = c1; k < std::min(nt, c1 + W); ++k)
                for (int i = k + 1; i < nt; ++i)
                    for (int p = 0; p < P; ++p)
                        for (int hh = 0; hh < 2; ++hh) gram_unit(ride, p, i, k, hh);
        }
        auto slices_of_column = [&](int k) {
            if (!slice_items) return;
            auto push = [&](int p, int i, int need_i) {
                for (int hh = 0; hh < 2; ++hh) {
                    cur.push_back(make_int4(agp::ITEM_SLICE | (hh << 8), p, k, i));
                    cur.push_back(make_int4(0, (need_i > 0 ? need_i : 0) << 16, -1, 0));
                }
            };
            for (int i = c1; i < nt; ++i)
                for (int p = 0; p < P; ++p) push(p, i, 2 * (k + 1));
            // appended rows nt + a: tile (nt + a, k) is final after k - a + 1 panels of that row; a = k + 1 is structurally
            // zero there (zero planes, read by the CTA pair of row a - 1).  The last super-column's are read by the lauum pass only.
            if (aug && (c1 < nt || with_lauum))
                for (int a = 0; a < std::min(nt, c1 + 1); ++a)
                    for (int p = 0; p < P; ++p) push(p, nt + a, 2 * (k - a + 1));
        };
        size_t ride_pos = 0;
        auto deal = [&](int j) {  // behind the POTF2 items of the segment's j-th block column
            const size_t hi = (ride.size() / 2) * (size_t)(j + 1) / (size_t)(c1 - c0);
            for (; ride_pos < hi; ++ride_pos) {
                cur.push_back(ride[2 * ride_pos]);
                cur.push_back(ride[2 * ride_pos + 1]);
            }
        };
        if (c0 == 0) {
            for (int p = 0; p < P; ++p) b.tile(p, 0, 0, 0);  // starts the forward solve: y_0 = xs
        } else {
            for (int p = 0; p < P; ++p)
                for (int k = c0; k < c1; ++k) {
                    for (int i = k; i < nt; ++i) b.coverage(p, i, k) = c0;  // the int8 update
                    for (int a = 0; aug && a < c0; ++a) b.coverage(p, nt + a, k) = c0;
                }
        }
        for (int p = 0; p < P; ++p) b.potf2(p, c0);  // s >= 1: the diagonal tile needs no DIAG item (no contraction left)
        deal(0);
        for (int k = c0; k < c1; ++k) {
            if (k + 1 < nt)
                for (int p = 0; p < P; ++p) b.tile(p, k + 1, k, k);  // panels of tile row k + 1 first
            std::vector<T> bulk;
            for (int i = k + 2; i < nt; ++i)
                for (int p = 0; p < P; ++p) bulk.push_back({p, i, k});
            for (int a = 0; aug && a <= k; ++a)
                for (int p = 0; p < P; ++p) bulk.push_back({p, nt + a, k});
            const size_t nb = bulk.size();
            if (k + 1 < c1) {
                const size_t a1 = (size_t)(pos_diag * nb), a2 = std::max(a1, (size_t)(pos_potf2 * nb));
                for (size_t a = 0; a < a1; ++a) b.tile(bulk[a].p, bulk[a].i, bulk[a].k, bulk[a].k);
                for (int p = 0; p < P; ++p) b.tile(p, k + 1, k + 1, k + 1);
                for (size_t a = a1; a < a2; ++a) b.tile(bulk[a].p, bulk[a].i, bulk[a].k, bulk[a].k);
                for (int p = 0; p < P; ++p) b.potf2(p, k + 1);
                deal(k + 1 - c0);
                for (size_t a = a2; a < nb; ++a) b.tile(bulk[a].p, bulk[a].i, bulk[a].k, bulk[a].k);
            } else {
                for (size_t a = 0; a < nb; ++a) b.tile(bulk[a].p, bulk[a].i, bulk[a].k, bulk[a].k);
            }
            if (k > c0) slices_of_column(k - 1);
        }
        slices_of_column(c1 - 1);
        deal(c1 - c0 - 1);
        if (gram_lead > 0 && c0 == 0) fuse_gram_items(P, nt_stride, gram_lead, cur);
        seg.push_back((int)(items.size() / 2));
        items.insert(items.end(), cur.begin(), cur.end());
    }
    seg.push_back((int)(items.size() / 2));
}

// Gram units as queue items (ITEM_GRAM, agp_chol_gram.cu).  Takes a schedule of build_queue and (i) gives every DIAG /
// PANEL item that is the FIRST to touch its tile half (contraction range starting at 0: its accumulators start from
// minus the Gram tile) a wait for that half's flag, (ii) inserts the Gram unit of the tile half `lead` items ahead of
// that reader (all units whose reader sits among the first `lead` items open the queue, in reader order).  Producers
// stay earlier in the queue than their consumers, so in-order popping remains deadlock-free; the units inherit the
// readers' order, which spreads them over the factorisation in proportion to where tiles are first needed.
// Flags: one int per (particle, lower tile, half) behind fdone, index relative to SchedView::head (a diagonal tile uses
// the flag of its first half for both units).
static int gram_flag_base(int P, int nt_stride) { return 32 + 3 * P * nt_stride + P; }
static int gram_flag(int P, int nt_stride, int p, int i, int k, int hh) {
    const int T = nt_stride * (nt_stride + 1) / 2;
    return gram_flag_base(P, nt_stride) + ((p * T + i * (i + 1) / 2 + k) * 2 + hh);
}
static size_t sync_ints(int P, int nt_stride, bool fused) {
    return 32 + (size_t)3 * P * nt_stride + P + (fused ? (size_t)P * nt_stride * (nt_stride + 1) : 0);
}
static void fuse_gram_items(int P, int nt_stride, int lead, std::vector<int4>& items) {
    const size_t n = items.size() / 2;
    if (lead < 1) lead = 1;
    struct G { size_t reader; int4 it, dep; };
    std::vector<G> grams;
    for (size_t c = 0; c < n; ++c) {
        int4& it = items[2 * c];
        int4& dep = items[2 * c + 1];
        const int type = it.x & 0xff;
        if (type == agp::ITEM_POTF2 || type == agp::ITEM_SLICE) continue;
        const int j0 = dep.x & 0xffff;
        if (j0 != 0 || dep.z >= 0) continue;  // a continuation of an earlier item of the same tile half
        const int hh = (it.x >> 8) & 1, p = it.y, k = it.z, i = it.w;
        // a DIAG item reads 16x16 blocks from both row halves of its tile (agp_chol_diag.cu): the two units of a diagonal
        // tile share one flag and its first readers wait for both
        const bool dg = type == agp::ITEM_DIAG;
        const int flag = gram_flag(P, nt_stride, p, i, k, dg ? 0 : hh);
        dep.z = flag;
        dep.w = dg ? 2 : 1;
        grams.push_back({c, make_int4(agp::ITEM_GRAM | (hh << 8), p, k, i), make_int4(0, 0, flag, 0)});
    }
    std::vector<int4> out;
    out.reserve(items.size() + 2 * grams.size());
    size_t g = 0;
    for (size_t c = 0; c < n; ++c) {
        while (g < grams.size() && grams[g].reader <= c + (size_t)lead) {
            out.push_back(grams[g].it);
            out.push_back(grams[g].dep);
            ++g;
        }
        out.push_back(items[2 * c]);
        out.push_back(items[2 * c + 1]);
    }
    items.swap(out);
}

// General schedule (simple block-column order) for the two continuations of a factorisation:
//   first_row > 0        only tile rows >= first_row are (re)computed: the data prefix grew and rows
//                        above keep their factor (agp_lml_run_append)
//   nt_total > nt        tile rows >= nt hold prediction points: they are solved against every L_kk
//                        (PANEL) and their mutual tiles receive the Schur complement
//                        K_22 - L_21 L_21^T as store-only items over [0, nt) (agp_predict_batch)
static void build_queue_general(int P, int nt, int nt_total, int first_row, std::vector<int4>& items, bool trailing_diag_only = false) {
    auto push = [&](int type, int h, int partial, int p, int k, int i, int j0, int j1, int need) {
        int flags = partial ? agp::ITEM_PARTIAL : 0;
        if (type == agp::ITEM_PANEL && k == 0) flags |= agp::ITEM_YINIT;
        items.push_back(make_int4(type | (h << 8) | flags, p, k, i));
        items.push_back(pack_dep(j0, j1, type == agp::ITEM_POTF2 ? 0 : 2 * j1, type == agp::ITEM_PANEL ? 2 * j1 : 0, -1, need));
    };
    items.clear();
    if (first_row == 0 && nt_total > nt && nt > 0) {
        // prediction rows behind a full factorisation: the observation block keeps the tuned schedule, the appended tile
        // rows are solved block column by block column behind it (each panel needs tile row k final and its own
        // earlier tiles), then the Schur complement items
        int order = 3;
        if (const char* e = getenv("AGP_ORDER")) order = atoi(e);
        build_queue(P, nt, nt_total, order, items);
        for (int k = 0; k < nt; ++k)
            for (int p = 0; p < P; ++p)
                for (int i = nt; i < nt_total; ++i)
                    for (int h = 0; h < 2; ++h) push(agp::ITEM_PANEL, h, 0, p, k, i, 0, k, 0);
        for (int p = 0; p < P; ++p)
            for (int i = nt; i < nt_total; ++i)
                for (int k = trailing_diag_only ? i : nt; k <= i; ++k)  // marginals: only the diagonal tiles of the Schur complement
                    for (int h = 0; h < 2; ++h) push(i == k ? agp::ITEM_DIAG : agp::ITEM_PANEL, h, 1, p, k, i, 0, nt, 0);
        return;
    }
    for (int k = 0; k < nt; ++k) {
        if (k >= first_row) {
            for (int p = 0; p < P; ++p)
                for (int h = 0; h < 2; ++h) push(agp::ITEM_DIAG, h, 0, p, k, k, 0, k, 0);
            for (int p = 0; p < P; ++p) push(agp::ITEM_POTF2, 0, 0, p, k, k, 0, 0, 2);
        }
        const int i0 = (k + 1 > first_row) ? k + 1 : first_row;
        for (int p = 0; p < P; ++p)
            for (int i = i0; i < nt_total; ++i)
                for (int h = 0; h < 2; ++h) push(agp::ITEM_PANEL, h, 0, p, k, i, 0, k, 0);
    }
    for (int p = 0; p < P; ++p)
        for (int i = nt; i < nt_total; ++i)
            for (int k = nt; k <= i; ++k)
                for (int h = 0; h < 2; ++h) push(i == k ? agp::ITEM_DIAG : agp::ITEM_PANEL, h, 1, p, k, i, 0, nt, 0);
}

// Schedule of the identity-augmented factorisation (agp_lml_grad_batch): the matrix [[K, I], [I, 0]] is
// factored over its first nt tile rows; tile row nt + a then solves to block row a of L^{-T}
// (non-zero from block column a on: PANEL items over [a, k), k >= a, i.e. a blocked trtri at n^3/3 flops),
// and the trailing tile (nt + a, nt + b), b <= a, receives the Schur complement
// 0 - sum_{j >= a} L^{-T}[a][j] L^{-T}[b][j]^T = -K^{-1}[a][b] as a store-only item over [a, nt) (a blocked
// lauum at n^3/3 flops).  The forward-solve entries of the appended rows end as 0 - L^{-T} z = -alpha.
static void build_queue_inverse(int P, int nt, int nt_stride, int order, std::vector<int4>& items, bool with_lauum = true) {
    build_queue(P, nt, nt_stride, order, items);
    for (int k = 0; k < nt; ++k)
        for (int p = 0; p < P; ++p)
            for (int a = 0; a <= k; ++a)
                for (int h = 0; h < 2; ++h) {
                    items.push_back(make_int4(agp::ITEM_PANEL | (h << 8) | (k == a ? agp::ITEM_YINIT : 0), p, k, nt + a));
                    items.push_back(pack_dep(a, k, 2 * k, 2 * (k - a), -1, 0));
                }
    if (!with_lauum) return;  // dLML/dnoise needs |L^{-1}|_F only
    for (int p = 0; p < P; ++p)
        for (int a = 0; a < nt; ++a)
            for (int b = 0; b <= a; ++b)
                for (int h = 0; h < 2; ++h) {
                    items.push_back(make_int4((a == b ? agp::ITEM_DIAG : agp::ITEM_PANEL) | (h << 8) | agp::ITEM_PARTIAL, p, nt + b, nt + a));
                    items.push_back(pack_dep(a, nt, 2 * (nt - b), a == b ? 0 : 2 * (nt - a), -1, 0));
                }
}

// Plain LML run (no appended rows, no continuation, every program fits the item's shared-memory cache): may the Gram units
// ride in the queue, and does this handle want them to for this size?
// Device buffer for a new queue of `bytes`: the smallest spare that fits, else a fresh allocation rounded up to a power of
// two (so that it fits most later queues of its size class).
static int queue_buffer(agp_handle* h, size_t bytes, agp_handle::Queue* qu) {
    size_t best = h->spare_queues.size();
    for (size_t a = 0; a < h->spare_queues.size(); ++a)
        if (h->spare_queues[a].second >= bytes && (best == h->spare_queues.size() || h->spare_queues[a].second < h->spare_queues[best].second)) best = a;
    if (best < h->spare_queues.size()) {
        qu->d_items = h->spare_queues[best].first;
        qu->cap_bytes = h->spare_queues[best].second;
        h->spare_queues.erase(h->spare_queues.begin() + (long)best);
        return AGP_OK;
    }
    size_t cap = 4096;
    while (cap < bytes) cap *= 2;
    StallTimer st_("queue_buffer.cudaMalloc", cap);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&qu->d_items), cap);
    if (e != cudaSuccess) {
        cudaGetLastError();
        qu->d_items = nullptr;
        return fail(h, AGP_ERR_NOMEM, std::string("work queue allocation failed: ") + cudaGetErrorString(e));
    }
    qu->cap_bytes = cap;
    return AGP_OK;
}
// Evict the least recently used queue when the cache is full; its buffer becomes a spare.
static int evict_queue_if_full(agp_handle* h) {
    if (h->queues.size() < agp_handle::kMaxQueues) return AGP_OK;
    auto lru = h->queues.begin();
    for (auto jt = h->queues.begin(); jt != h->queues.end(); ++jt)
        if (jt->second.last_use < lru->second.last_use) lru = jt;
    AGP_CUDA(h, cudaStreamSynchronize(h->stream));  // a launch that reads it may still be in flight
    h->spare_queues.emplace_back(lru->second.d_items, lru->second.cap_bytes);
    h->queues.erase(lru);
    size_t own = 0, small = h->spare_queues.size();  // spares allocated one by one: drop the smallest beyond the limit
    for (size_t a = 0; a < h->spare_queues.size(); ++a) {
        if (h->in_slab(h->spare_queues[a].first)) continue;
        ++own;
        if (small == h->spare_queues.size() || h->spare_queues[a].second < h->spare_queues[small].second) small = a;
    }
    if (own > agp_handle::kMaxSpareQueues) {
        cudaFree(h->spare_queues[small].first);
        h->spare_queues.erase(h->spare_queues.begin() + (long)small);
    }
    return AGP_OK;
}

static bool gram_as_items(const agp_handle* h, int first_row) {
    const BatchView& v = h->view;
    if (h->aug_identity || first_row != 0 || v.nt_total != v.nt || h->comp.M != 0 || v.max_prog_len > 64) return false;
    return h->fuse_gram < 0 ? v.nt <= agp_handle::kGramItemsMaxNt : h->fuse_gram != 0;
}
// CTAs of the persistent kernel.  Two per SM hide each other's dependency waits and pipeline latencies when there is work for
// all of them; a small batch is bound by its per-particle chain (POTF2 -> panel -> diagonal tile), and every item of the
// chain runs faster with the SM to itself.  Measured (tools/ctas_ab.sh, one against two CTAs per SM, plain LML):
// P = 16: n = 512 -12 %, 1024 -18 %, 1536 -12 %, 2048 -7 %; P = 24: -15 / -15 / . / -2 %; P = 32: -13 / -9 / . / +2.5 %;
// P = 48: +3 / +4 / . / +7 %; n = 128, P = 64: -17 %; gradient calls (twice the tile rows): P = 16: n = 512 -8 %, 1024 -11 %,
// 2048 0; P = 32: 0 / +1 / +7 %.  The lock-step rejuvenation loops (rejuvenate.py) call with 10 - 20 active particles.
// One CTA per SM runs the kernel's second instantiation (launch bounds (256, 1), no 128-register cap: Makefile RDC_SOLO), another
// -3.5 .. -9 % on plain LML runs and bitwise the same results; with it: P = 40: n = 512 0.341 -> 0.293 ms, 1024 1.079 -> 1.007;
// P = 48: 0.356 -> 0.330, 1.185 -> 1.165; n = 2048: P = 24 3.03 -> 2.88, P = 32 3.63 -> 3.59 (two -> one CTA per SM).
static int chol_ctas(const agp_handle* h) {
    static const int forced_total = [] { const char* e = getenv("AGP_CTAS"); return e ? atoi(e) : 0; }();  // developer A/B: absolute grid size
    if (forced_total > 0) return std::min(forced_total, 2 * h->num_sms);
    if (h->ctas_per_sm > 0) return h->ctas_per_sm * h->num_sms;
    const int nt = h->view.nt_total > 0 ? h->view.nt_total : h->view.nt, P = h->P;
    bool one;
    if (h->aug_identity) one = P <= 24 - std::max(0, h->view.nt - 8);
    else if (nt <= 1) one = P <= h->num_sms;
    else if (nt == 2) one = P <= 64;
    else one = 2 * P <= 100 - 6 * std::max(0, nt - 8);  // with the uncapped one-CTA instantiation: 48 particles at n <= 1024, 32 at n = 2048 still win
    return (one ? 1 : 2) * h->num_sms;
}
static int gram_items_lead(const agp_handle* h) { return h->gram_lead > 0 ? h->gram_lead : chol_ctas(h); }
// one CTA per SM runs the instantiation compiled for it (no register cap; AGP_CHOL_SOLO=0: the shipped two-CTA build on a smaller grid)
static void launch_chol_by_grid(const agp_handle* h, const BatchView& v, const agp::SchedView& q) {
    static const bool solo_ok = [] { const char* e = getenv("AGP_CHOL_SOLO"); return !e || atoi(e) != 0; }();
    const int ctas = chol_ctas(h);
    if (solo_ok && ctas <= h->num_sms) agp::launch_chol_solo(v, q, h->tma, ctas, h->stream);
    else agp::launch_chol(v, q, h->tma, ctas, h->stream);
}

// Super-column width: AGP_OZ_W / agp_set_hybrid, or by size (measured, 64 particles: n = 2048: 4 best, 6.72 ms against
// 6.81 with 3 and 6.98 with 2; n = 4096: 2..4 within 0.5 %; n = 8192: 3 best, 205 ms against 207 / 208 with 2 / 4)
// by size (tools/width_sweep.py, final int8 kernel): n = 2048: W = 3 6.16, 4 6.16, 2 6.31, 5 6.30 ms; n = 4096: W = 2 29.4, 3 29.7,
// 4 30.0; n = 8192: W = 2 175.4, 3 177.1, 4 177.6 — the deeper the contractions, the less a launch's fixed costs weigh
static int hybrid_width(const agp_handle* h, int nt) { return h->oz_width > 0 ? h->oz_width : (nt >= 32 ? 2 : 4); }

// Plain LML run with at least two super-columns: does this handle factor it through the hybrid schedule?
static bool use_hybrid(const agp_handle* h, int first_row) {
    const BatchView& v = h->view;
    if (h->force_plain || first_row != 0 || h->comp.M != 0 || v.nt <= hybrid_width(h, v.nt)) return false;
    if (h->aug_identity ? (!h->oz_aug || v.nt_total != 2 * v.nt || !(h->min_noise >= agp_handle::kAugMinNoise)) : v.nt_total != v.nt) return false;
    if (h->oz_nomem) return false;
    // the thresholds are measured at 64 particles; a small batch of plain LML runs leaves the int8 launches too few units per
    // CTA for their fixed costs (tools/hybrid_small_batch.sh, FP64 schedule against hybrid: P = 8, n = 2048: 1.47 / 1.69 ms,
    // n = 4096: 8.70 / 6.26; P = 16: n = 1536 1.26 / 1.39, 1792 1.73 / 1.82, 2048 2.30 / 2.30; P = 32: n = 2048 3.95 / 3.64),
    // so the switch moves up by one block column per 8 particles below 48; the gradient calls win from 8 block columns on
    // at 16 particles as well (n = 1024: 1.60 / 1.55 ms, 1536: 3.59 / 3.10)
    const int small_batch = h->aug_identity ? 0 : std::max(0, (48 - h->P + 7) / 8);
    const int min_nt = (h->aug_identity ? std::min(h->oz_min_nt, h->oz_min_nt_aug) : h->oz_min_nt) + small_batch;
    return h->oz_mode < 0 ? (v.nt >= min_nt && h->fuse_gram != 1) : h->oz_mode != 0;
}

static int run_hybrid(agp_handle* h, float* kernel_ms, long long* d_trace = nullptr);

static int run_fused(agp_handle* h, long long* d_trace = nullptr, float* kernel_ms = nullptr, int first_row = 0) {
    if (use_hybrid(h, first_row)) return run_hybrid(h, kernel_ms, d_trace);
    const BatchView& v = h->view;
    const int P = h->P, nt = v.nt;
    const int nt_stride = h->ld / TB;
    const int nt_total = v.nt_total;
    // plain LML run (no appended rows, no continuation): the Gram units ride in the queue (agp_chol_gram.cu caches a program in shared memory)
    const bool fused = gram_as_items(h, first_row);
    const int lead = gram_items_lead(h);
    auto key = std::make_tuple(P, nt, nt_total, fused ? -4 : h->aug_identity ? (h->trtri_only ? -2 : -1) : (h->pred_diag_only && first_row == 0 && nt > 0 ? -3 : first_row), nt_stride);
    auto it = h->queues.find(key);
    if (it == h->queues.end()) {
        std::vector<int4> items;
        if (h->aug_identity) build_queue_inverse(P, nt, nt_stride, h->order, items, !h->trtri_only);
        else if (first_row == 0 && nt_total == nt) build_queue(P, nt, nt_stride, h->order, items);
        else build_queue_general(P, nt, nt_total, first_row, items, h->pred_diag_only);
        if (fused) fuse_gram_items(P, nt_stride, lead, items);
        agp_handle::Queue qu;
        qu.n_items = (int)(items.size() / 2);
        int qrc;
        if ((qrc = evict_queue_if_full(h)) != AGP_OK || (qrc = queue_buffer(h, items.size() * sizeof(int4), &qu)) != AGP_OK) return qrc;
        // pageable source: the copy is staged before the call returns, so `items` may go out of scope
        AGP_CUDA(h, cudaMemcpyAsync(qu.d_items, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
        AGP_CUDA(h, cudaStreamSynchronize(h->stream));
        it = h->queues.emplace(key, qu).first;
    }
    it->second.last_use = ++h->queue_clock;
    // counters: [0] head, [1] error, [32 ..] rowdone[P][nt_stride], diagu[P][nt_stride], ppre[P][nt_stride], fdone[P],
    // and with Gram items: gram flags [P][nt_stride (nt_stride + 1) / 2][2]
    const size_t n_sync = sync_ints(P, nt_stride, fused);
    int rc = grow_device(h, &h->d_sync, &h->cap_sync, n_sync * sizeof(int));
    if (rc != AGP_OK) return rc;
    AGP_CUDA(h, cudaMemsetAsync(h->d_sync, 0, n_sync * sizeof(int), h->stream));
    if (first_row > 0) {
        // continuation: tile rows above first_row are final for every block column, and the first
        // first_row diagonal tiles are factored
        std::vector<int> init(n_sync, 0);
        for (int p = 0; p < P; ++p) {
            for (int i = 0; i < first_row; ++i) init[32 + (size_t)p * nt_stride + i] = 2 * i;
            init[32 + (size_t)3 * P * nt_stride + p] = first_row;
        }
        AGP_CUDA(h, cudaMemcpyAsync(h->d_sync, init.data(), n_sync * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        AGP_CUDA(h, cudaStreamSynchronize(h->stream));  // `init` is pageable and goes out of scope
    }
    SchedView q;
    q.items = it->second.d_items;
    q.n_items = it->second.n_items;
    q.head = h->d_sync;
    q.err = h->d_sync + 1;
    q.rowdone = h->d_sync + 32;
    q.diagu = q.rowdone + (size_t)P * nt_stride;
    q.ppre = q.diagu + (size_t)P * nt_stride;
    q.fdone = q.ppre + (size_t)P * nt_stride;
    q.nt_stride = nt_stride;
    q.trace = d_trace;
    q.wait_timeout_ns = h->wait_timeout_ns;
    if (kernel_ms) AGP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if (fused) {
        h->launches -= 1;  // no Gram launch: the units are items of the persistent kernel
    } else if (h->aug_identity) {
        // the kernel tree is evaluated over the observation block only; the appended [I 0] rows are plain stores
        BatchView obs = v;
        obs.nt_total = v.nt;
        agp::launch_gramfill(obs, P, first_row, h->stream);
        agp::launch_augfill(v, P, h->stream);
        h->launches += 1;
    } else {
        agp::launch_gramfill(v, P, first_row, h->stream);
    }
    if (h->comp.M > 0) {
        agp::launch_component_fill(v, P, h->comp, h->stream);
        h->launches += 1;
    }
    if (kernel_ms) {
        AGP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
        AGP_CUDA(h, cudaEventSynchronize(h->ev1));
        AGP_CUDA(h, cudaEventElapsedTime(&kernel_ms[0], h->ev0, h->ev1));
        AGP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    }
    if (h->tma_L != v.L || h->tma_ld != h->ld || h->tma_rows != (long long)P * h->ld) {
        if (!agp::make_tma_maps(v.L, h->ld, (long long)P * h->ld, &h->tma)) return fail(h, AGP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the factor matrix");
        h->tma_L = v.L;
        h->tma_ld = h->ld;
        h->tma_rows = (long long)P * h->ld;
    }
    launch_chol_by_grid(h, v, q);
    if (kernel_ms) {
        AGP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
        AGP_CUDA(h, cudaEventSynchronize(h->ev1));
        AGP_CUDA(h, cudaEventElapsedTime(&kernel_ms[1], h->ev0, h->ev1));
    }
    h->launches += 2;
    h->n_factored = v.n;
    h->factor_clean = false;
    return check_launch(h, "chol");
}

// Hybrid factorisation: Gram fill, row scales, then per super-column  [digit planes of the previous super-column's
// panels -> int8 update of this super-column's tiles ->] one launch of the persistent kernel over the segment.
static int run_hybrid(agp_handle* h, float* kernel_ms, long long* d_trace) {
    StallTimer st_all_("run_hybrid", (size_t)h->P);
    const BatchView& v = h->view;
    const int P = h->P, nt = v.nt, ld = h->ld;
    const int nt_stride = ld / TB;
    const int W = hybrid_width(h, nt);
    // The Gram units can ride in the segments' queues (every program must fit the item's shared-memory cache).  Measured and
    // NOT the default (profiles/r02_hybrid.txt): n = 2048 x 64 7.15 ms against 6.72 with the Gram launch in front, n = 8192
    // 204 against 198 ms — next to a DMMA item a unit's DFMAs wait for the shared FP64 datapath, and the units that ride
    // behind the POTF2 items slow exactly the chain the segment is bound by.  AGP_OZ_RIDE=1 enables it.
    const bool aug = h->aug_identity;  // gradient calls: [I 0] rows appended (tile rows nt .. 2 nt), agp_lml_grad_batch / _noise_batch
    const bool ride = !aug && h->oz_ride && v.max_prog_len <= 64 && h->fuse_gram != 0;
    // digit planes cut by launches between the segments (default) or by SLICE items of the segments (AGP_OZ_SLICE_ITEMS=1)
    const bool slice_items = h->oz_slice_items && !ride;
    const bool lauum = aug && !h->trtri_only;
    auto key = std::make_tuple(P, nt, nt, -5 - W - (ride ? 1000 : 0) - (aug ? 2000 : 0) - (slice_items ? 4000 : 0) - (lauum ? 8000 : 0), nt_stride);
    auto it = h->queues.find(key);
    if (it == h->queues.end()) {
        StallTimer st_("run_hybrid.queue_miss", (size_t)h->queues.size());
        std::vector<int4> items;
        agp_handle::Queue qu;
        build_queue_hybrid(P, nt, nt_stride, W, ride ? gram_items_lead(h) : 0, items, qu.seg, aug, slice_items, lauum);
        qu.n_items = (int)(items.size() / 2);
        int qrc;
        if ((qrc = evict_queue_if_full(h)) != AGP_OK || (qrc = queue_buffer(h, items.size() * sizeof(int4), &qu)) != AGP_OK) return qrc;
        AGP_CUDA(h, cudaMemcpyAsync(qu.d_items, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
        AGP_CUDA(h, cudaStreamSynchronize(h->stream));
        it = h->queues.emplace(key, qu).first;
    }
    it->second.last_use = ++h->queue_clock;
    const agp_handle::Queue& qu = it->second;
    int rc;
    if ((rc = grow_device(h, &h->d_S, &h->cap_S, (size_t)agp::OZ_SLICES * P * ld * ld)) != AGP_OK) {
        if (rc != AGP_ERR_NOMEM) return rc;
        // the digit planes are as large as the factors themselves: without room for them the FP64 schedule does the job
        h->oz_nomem = true;
        h->err.clear();
        return run_fused(h, nullptr, kernel_ms);
    }
    if ((rc = grow_device(h, &h->d_rscale, &h->cap_rscale, (size_t)P * ld * 16)) != AGP_OK) return rc;
    if (h->oz_S != h->d_S || h->oz_ld != ld || h->oz_P != P) {
        if (!agp::make_ozaki_maps(h->d_S, ld, P, &h->ozmaps)) return fail(h, AGP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the digit planes");
        h->oz_S = h->d_S;
        h->oz_ld = ld;
        h->oz_P = P;
    }
    const size_t n_sync = sync_ints(P, nt_stride, ride);
    if ((rc = grow_device(h, &h->d_sync, &h->cap_sync, n_sync * sizeof(int))) != AGP_OK) return rc;
    AGP_CUDA(h, cudaMemsetAsync(h->d_sync, 0, n_sync * sizeof(int), h->stream));
    if (h->tma_L != v.L || h->tma_ld != ld || h->tma_rows != (long long)P * ld) {
        if (!agp::make_tma_maps(v.L, ld, (long long)P * ld, &h->tma)) return fail(h, AGP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the factor matrix");
        h->tma_L = v.L;
        h->tma_ld = ld;
        h->tma_rows = (long long)P * ld;
    }
    SchedView q;
    q.head = h->d_sync;
    q.err = h->d_sync + 1;
    q.rowdone = h->d_sync + 32;
    q.diagu = q.rowdone + (size_t)P * nt_stride;
    q.ppre = q.diagu + (size_t)P * nt_stride;
    q.fdone = q.ppre + (size_t)P * nt_stride;
    q.nt_stride = nt_stride;
    q.trace = nullptr;
    q.wait_timeout_ns = h->wait_timeout_ns;
    // stage timing (agp_lml_stage_times): every stage bracketed by events and a synchronisation, so the four sums are
    // per-kernel times, not a pipelined step time
    float acc_ms[4] = {0.f, 0.f, 0.f, 0.f};
    auto tic = [&]() -> int {
        if (kernel_ms) AGP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
        return AGP_OK;
    };
    auto toc = [&](int slot) -> int {
        if (!kernel_ms) return AGP_OK;
        float ms = 0.f;
        AGP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
        AGP_CUDA(h, cudaEventSynchronize(h->ev1));
        AGP_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        acc_ms[slot] += ms;
        return AGP_OK;
    };
    if (!ride) {
        if ((rc = tic()) != AGP_OK) return rc;
        if (aug) {
            // the kernel tree is evaluated over the observation block only; the appended [I 0] rows are plain stores
            BatchView obs = v;
            obs.nt_total = v.nt;
            agp::launch_gramfill(obs, P, 0, h->stream);
            agp::launch_augfill(v, P, h->stream);
            h->launches += 1;
        } else {
            agp::launch_gramfill(v, P, 0, h->stream);
        }
        h->launches += 1;
        if ((rc = toc(0)) != AGP_OK) return rc;
    }
    const int oz_variant = h->oz_variant < 0 ? (nt >= 24 ? 3 : 2) : h->oz_variant;
    const int n_seg = (int)qu.seg.size() - 1;
    BatchView vq = v;  // what the persistent kernel sees: + the digit-plane buffers for its SLICE items
    vq.oz_S = reinterpret_cast<signed char*>(h->d_S);
    vq.oz_rscale = h->d_rscale;
    vq.oz_plane = (long long)P * ld * ld;
    if (slice_items) {
        // the Gram diagonal is in place (launch above): row scales before the first segment
        if ((rc = tic()) != AGP_OK) return rc;
        agp::launch_ozaki_rowscale(v.L, v.mat_stride, ld, P, h->d_rscale, h->stream, nt * TB, aug ? v.noise : nullptr);
        h->launches += 1;
        if ((rc = toc(0)) != AGP_OK) return rc;
    }
    for (int s = 0; s < n_seg; ++s) {
        const int c0 = s * W, c1 = std::min(nt, c0 + W);
        if (s == 1 && !slice_items) {
            // every diagonal tile of K is in place (Gram launch, or the units that rode in segment 0): row scales
            if ((rc = tic()) != AGP_OK) return rc;
            agp::launch_ozaki_rowscale(v.L, v.mat_stride, ld, P, h->d_rscale, h->stream, nt * TB, aug ? v.noise : nullptr);
            h->launches += 1;
            if ((rc = toc(0)) != AGP_OK) return rc;
        }
        if (s > 0) {
            if (!slice_items) {
                if ((rc = tic()) != AGP_OK) return rc;
                // digit planes of the previous super-column's finished panels: tile rows below it, and (aug) the appended rows
                // nt + a, a <= c0 (row a = j + 1 is structurally zero at block column j: zero planes, read by the pair of row a - 1)
                agp::launch_ozaki_slice(v.L, v.mat_stride, ld, aug ? nt + std::min(nt, c0 + 1) : nt, P, h->d_rscale, h->d_S, c0 - W, c0, c0, h->stream);
                h->launches += 1;
                if ((rc = toc(3)) != AGP_OK) return rc;
            }
            if ((rc = tic()) != AGP_OK) return rc;
            agp::OzakiParams prm{v.L, v.mat_stride, ld, nt, P, h->d_rscale, c0, c1, q.err, h->wait_timeout_ns, 0, 0, 0, 0, 0};
            if (aug) prm.r_lo = c0, prm.r_hi = nt + c0, prm.k_lo = c0, prm.k_hi = c1, prm.chi = c0;  // appended rows a < c0 over [a, c0)
            agp::launch_ozaki_update(prm, h->ozmaps, h->num_sms, h->stream, oz_variant);
            if ((rc = toc(2)) != AGP_OK) return rc;
            AGP_CUDA(h, cudaMemsetAsync(h->d_sync, 0, sizeof(int), h->stream));  // the queue head; the dependency counters keep counting
            h->launches += 1;
        }
        if ((rc = tic()) != AGP_OK) return rc;
        q.items = qu.d_items + 2 * (size_t)qu.seg[s];
        q.n_items = qu.seg[s + 1] - qu.seg[s];
        q.trace = d_trace ? d_trace + 8 * (long long)qu.seg[s] : nullptr;
        launch_chol_by_grid(h, vq, q);
        h->launches += 1;
        if ((rc = toc(1)) != AGP_OK) return rc;
    }
    if (lauum) {
        // lauum pass: -K^{-1}[a][b] = 0 - sum_{j >= a} L^{-T}[a][j] L^{-T}[b][j]^T into the trailing tile (nt + a, nt + b), b <= a — a pure
        // contraction over block columns [a, nt): all of it on the int8 path, after the planes of the last super-column's
        // appended panels are cut
        const int cl = (n_seg - 1) * W;
        if (!slice_items) {
            if ((rc = tic()) != AGP_OK) return rc;
            agp::launch_ozaki_slice(v.L, v.mat_stride, ld, 2 * nt, P, h->d_rscale, h->d_S, cl, nt, nt, h->stream);
            h->launches += 1;
            if ((rc = toc(3)) != AGP_OK) return rc;
        }
        if ((rc = tic()) != AGP_OK) return rc;
        agp::OzakiParams prm{v.L, v.mat_stride, ld, nt, P, h->d_rscale, 0, 0, q.err, h->wait_timeout_ns, nt, 2 * nt, nt, 2 * nt, nt};
        agp::launch_ozaki_update(prm, h->ozmaps, h->num_sms, h->stream, oz_variant);
        if ((rc = toc(2)) != AGP_OK) return rc;
        h->launches += 1;
    }
    if (kernel_ms) {
        kernel_ms[0] = acc_ms[0];
        kernel_ms[1] = acc_ms[1] + acc_ms[2] + acc_ms[3];
        for (int e = 0; e < 4; ++e) h->hybrid_ms[e] = acc_ms[e];
    }
    h->n_factored = v.n;
    h->factor_clean = false;
    return check_launch(h, "hybrid factorisation");
}

static int run_impl(agp_handle* h, float* kernel_ms) {
    const BatchView& v = h->view;
    const int P = h->P;
    if (P == 0) return AGP_OK;
    if (v.n == 0) {
        // empty mvnormal: logpdf of a 0-vector is 0 (initial SMC state, inference_smc_anneal_data.jl:185-189)
        size_t lml_bytes = (size_t)P * 8, res_bytes = (lml_bytes + 15) / 16 * 16 + (size_t)P * 4;
        AGP_CUDA(h, cudaMemsetAsync(h->d_res, 0, res_bytes, h->stream));
        return AGP_OK;
    }
    return run_fused(h, nullptr, kernel_ms);
}

int agp_lml_run(agp_handle* h) {
    if (!h) return AGP_ERR_ARG;
    if (!h->uploaded) return fail(h, AGP_ERR_STATE, "agp_lml_run: no resident batch (call agp_lml_upload first)");
    AGP_CUDA(h, cudaSetDevice(h->device));
    return run_impl(h, nullptr);
}

int agp_lml_stage_times(agp_handle* h, float* stage_ms) {
    if (!h || !stage_ms) return AGP_ERR_ARG;
    if (!h->uploaded) return fail(h, AGP_ERR_STATE, "agp_lml_stage_times: no resident batch");
    AGP_CUDA(h, cudaSetDevice(h->device));
    stage_ms[0] = stage_ms[1] = stage_ms[2] = 0.f;
    return run_impl(h, stage_ms);
}

int agp_lml_time(agp_handle* h, int32_t reps, float* ms_out) {
    if (!h || !ms_out || reps <= 0) return AGP_ERR_ARG;
    if (!h->uploaded) return fail(h, AGP_ERR_STATE, "agp_lml_time: no resident batch");
    AGP_CUDA(h, cudaSetDevice(h->device));
    AGP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    for (int r = 0; r < reps; ++r) {
        int rc = run_impl(h, nullptr);
        if (rc != AGP_OK) return rc;
    }
    AGP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    AGP_CUDA(h, cudaEventSynchronize(h->ev1));
    AGP_CUDA(h, cudaEventElapsedTime(ms_out, h->ev0, h->ev1));
    return AGP_OK;
}

int agp_lml_fetch(agp_handle* h, double* lml_out, int32_t* info_out) {
    if (!h) return AGP_ERR_ARG;
    if (!h->uploaded) return fail(h, AGP_ERR_STATE, "agp_lml_fetch: no resident batch");
    const int P = h->P;
    if (P == 0) return AGP_OK;
    if (!lml_out || !info_out) return fail(h, AGP_ERR_ARG, "agp_lml_fetch: null output");
    AGP_CUDA(h, cudaSetDevice(h->device));
    size_t info_off = align_up((size_t)P * 8, 16);
    size_t res_bytes = info_off + (size_t)P * 4;
    AGP_CUDA(h, cudaMemcpyAsync(h->h_res, h->d_res, res_bytes, cudaMemcpyDeviceToHost, h->stream));
    h->h_sync[0] = h->h_sync[1] = 0;
    if (h->d_sync) AGP_CUDA(h, cudaMemcpyAsync(h->h_sync, h->d_sync, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    AGP_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->h_sync[1] != 0) return fail(h, AGP_ERR_CUDA, "agp_lml_fetch: the work-queue scheduler reported a dependency time-out");
    memcpy(lml_out, h->h_res, (size_t)P * 8);
    memcpy(info_out, h->h_res + info_off, (size_t)P * 4);
    h->factor_clean = true;
    for (int p = 0; p < P; ++p) h->factor_clean = h->factor_clean && info_out[p] == 0;
    return AGP_OK;
}

int agp_lml_copy_factor(agp_handle* h, int32_t particle, double* factor_out, int32_t* ld_out) {
    if (!h || !ld_out) return AGP_ERR_ARG;
    if (!h->uploaded || h->n_factored < 0) return fail(h, AGP_ERR_STATE, "agp_lml_copy_factor: no factor resident (run the batch first)");
    if (particle < 0 || particle >= h->P) return fail(h, AGP_ERR_ARG, "agp_lml_copy_factor: particle out of range");
    *ld_out = h->ld;
    if (!factor_out) return AGP_OK;
    AGP_CUDA(h, cudaSetDevice(h->device));
    AGP_CUDA(h, cudaMemcpyAsync(factor_out, h->view.L + (long long)particle * h->view.mat_stride, (size_t)h->ld * h->ld * 8, cudaMemcpyDeviceToHost,
                                h->stream));
    AGP_CUDA(h, cudaStreamSynchronize(h->stream));
    return AGP_OK;
}

int agp_lml_device_results(agp_handle* h, double** lml_dev, int32_t** info_dev) {
    if (!h || !lml_dev || !info_dev) return AGP_ERR_ARG;
    if (!h->uploaded) return fail(h, AGP_ERR_STATE, "agp_lml_device_results: no resident batch");
    *lml_dev = h->view.lml;
    *info_dev = h->view.info;
    return AGP_OK;
}

int agp_lml_run_append(agp_handle* h) {
    if (!h) return AGP_ERR_ARG;
    if (!h->uploaded) return fail(h, AGP_ERR_STATE, "agp_lml_run_append: no resident batch");
    if (h->n_pred > 0) return fail(h, AGP_ERR_STATE, "agp_lml_run_append: the resident batch carries appended rows");
    if (h->n_factored < 0 || !h->factor_clean)
        return fail(h, AGP_ERR_STATE, "agp_lml_run_append: no clean factor resident (run + fetch with info == 0 for every particle first)");
    if (h->view.n < h->n_factored) return fail(h, AGP_ERR_STATE, "agp_lml_run_append: the data prefix shrank; use agp_lml_run");
    AGP_CUDA(h, cudaSetDevice(h->device));
    if (h->P == 0 || h->view.n == 0) return run_impl(h, nullptr);
    // tile rows that were complete (all 128 rows observed) keep their factor, z and running sums
    const int first_row = h->n_factored / TB;
    if (first_row >= h->view.nt) return AGP_OK;  // nothing new: the resident results stand
    return run_fused(h, nullptr, nullptr, first_row);
}

static int predict_impl(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops, const int32_t* param_off, const int32_t* n_params,
                        const double* params, const double* noise, const double* ts, const double* xs, int32_t n, const double* ts_pred,
                        int32_t m, const double* noise_pred, double* mean_out, double* cov_out, int32_t* info_out, bool marginals) {
    if (!h) return AGP_ERR_ARG;
    if (m < 0 || (P > 0 && m > 0 && (!mean_out || !cov_out)) || (P > 0 && !info_out)) return fail(h, AGP_ERR_ARG, "agp_predict_batch: bad argument");
    int rc = upload_impl(h, P, prog_len, ops, param_off, n_params, params, noise, ts, xs, n, ts_pred, m, noise_pred);
    if (rc != AGP_OK) return rc;
    if (P == 0) return AGP_OK;
    h->pred_diag_only = marginals;
    const size_t mean_bytes = (size_t)P * m * 8, cov_bytes = (size_t)P * m * (marginals ? 1 : m) * 8;
    if ((rc = grow_device(h, &h->d_pred, &h->cap_pred, mean_bytes + cov_bytes + 16)) != AGP_OK) return rc;
    AGP_CUDA(h, cudaMemsetAsync(h->d_res, 0, align_up((size_t)P * 8, 16) + (size_t)P * 4, h->stream));  // info = 0 when nothing is factored
    {
        StallTimer st_("grad_impl.run_fused", (size_t)P);
        if ((rc = run_fused(h)) != AGP_OK) return rc;
    }
    h->n_factored = -1;  // the resident factor belongs to an augmented matrix
    if (m > 0) {
        const double* d_npred = h->view.noise + P;  // noise_pred[P] follows noise[P] in the input arena
        double* d_mean = h->d_pred;
        double* d_cov = h->d_pred + (size_t)P * m;
        if (marginals) agp::launch_predict_extract_marginals(h->view, P, d_npred, d_mean, d_cov, h->stream);
        else agp::launch_predict_extract(h->view, P, d_npred, d_mean, d_cov, h->stream);
        h->launches += 1;
        if ((rc = check_launch(h, "predict_extract")) != AGP_OK) return rc;
        AGP_CUDA(h, cudaMemcpyAsync(mean_out, d_mean, mean_bytes, cudaMemcpyDeviceToHost, h->stream));
        AGP_CUDA(h, cudaMemcpyAsync(cov_out, d_cov, cov_bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    std::vector<double> lml(P);
    rc = agp_lml_fetch(h, lml.data(), info_out);  // synchronises; info: LAPACK code of the training block
    h->factor_clean = false;
    return rc;
}

int agp_predict_batch(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops, const int32_t* param_off, const int32_t* n_params,
                      const double* params, const double* noise, const double* ts, const double* xs, int32_t n, const double* ts_pred,
                      int32_t m, const double* noise_pred, double* mean_out, double* cov_out, int32_t* info_out) {
    return predict_impl(h, P, prog_len, ops, param_off, n_params, params, noise, ts, xs, n, ts_pred, m, noise_pred, mean_out, cov_out, info_out, false);
}

int agp_predict_marginals_batch(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops, const int32_t* param_off,
                                const int32_t* n_params, const double* params, const double* noise, const double* ts, const double* xs, int32_t n,
                                const double* ts_pred, int32_t m, const double* noise_pred, double* mean_out, double* var_out, int32_t* info_out) {
    return predict_impl(h, P, prog_len, ops, param_off, n_params, params, noise, ts, xs, n, ts_pred, m, noise_pred, mean_out, var_out, info_out, true);
}

// Joint posterior of the summands of a sum kernel and of the observable at ts_pred (infer_gp_sum, src/GP.jl:904-993).
int agp_predict_sum_batch(agp_handle* h, int32_t P, int32_t M, const int32_t* prog_len, const int32_t* ops, const int32_t* param_off,
                          const int32_t* n_params, const double* params, const double* noise, const double* ts, const double* xs, int32_t n,
                          const double* ts_pred, int32_t m, const double* noise_pred, double* mean_out, double* cov_out, int32_t* info_out) {
    if (!h) return AGP_ERR_ARG;
    if (P < 0 || M < 1 || m < 0 || (P > 0 && (!prog_len || !ops || !param_off || !n_params || !noise || !info_out)) ||
        (P > 0 && m > 0 && (!mean_out || !cov_out || !ts_pred)))
        return fail(h, AGP_ERR_ARG, "agp_predict_sum_batch: bad argument");
    if (P == 0) return AGP_OK;
    AGP_CUDA(h, cudaSetDevice(h->device));
    // the observation kernel: k_1 + k_2 + ... + k_M as ONE program per particle (postfix: k_1 k_2 + k_3 + ...), its
    // parameter slice the concatenation of the summands' slices; and every summand compiled on its own
    std::vector<int32_t> c_len(P), c_ops, c_off, c_np(P);
    std::vector<AgpInstr> comp_instr;
    std::vector<int32_t> comp_off((size_t)P * M + 1, 0), comp_need((size_t)P * M, 1);
    {
        size_t o = 0, po = 0;
        std::string err;
        for (int p = 0; p < P; ++p) {
            int len = 0, shift = 0;
            for (int c = 0; c < M; ++c) {
                const int q = p * M + c;
                if (prog_len[q] <= 0 || n_params[q] < 0) return fail(h, AGP_ERR_PROGRAM, "agp_predict_sum_batch: particle " + std::to_string(p) + ": empty summand");
                for (int j = 0; j < prog_len[q]; ++j) {
                    c_ops.push_back(ops[o + j]);
                    c_off.push_back(param_off[o + j] + shift);
                }
                comp_off[q] = (int32_t)comp_instr.size();
                int need = 1;
                int rc = agp_compile_program(ops + o, param_off + o, prog_len[q], params ? params + po : nullptr, n_params[q], comp_instr, &need, err);
                if (rc != AGP_OK) return fail(h, rc, "agp_predict_sum_batch: particle " + std::to_string(p) + " summand " + std::to_string(c) + ": " + err);
                comp_need[q] = need;
                len += prog_len[q];
                if (c > 0) {
                    c_ops.push_back(AGP_OP_PLUS);
                    c_off.push_back(0);
                    ++len;
                }
                o += prog_len[q];
                po += n_params[q];
                shift += n_params[q];
            }
            c_len[p] = len;
            c_np[p] = shift;
        }
        comp_off[(size_t)P * M] = (int32_t)comp_instr.size();
    }
    const int mt = (M + 1) * m;  // appended rows: F_1(T*) ... F_M(T*), X(T*)
    std::vector<double> tp_ext((size_t)(mt > 0 ? mt : 1));
    for (int g = 0; g <= M; ++g)
        for (int a = 0; a < m; ++a) tp_ext[(size_t)g * m + a] = ts_pred[a];
    // the extraction kernel adds its noise_pred[p] to EVERY diagonal entry; only the X(T*) block carries it (:967), so
    // the resident copy is a zero vector and the host adds the X* diagonal after the copy back
    const std::vector<double> zeros((size_t)P, 0.0);
    int rc = upload_impl(h, P, c_len.data(), c_ops.data(), c_off.data(), c_np.data(), params, noise, ts, xs, n, tp_ext.data(), mt, zeros.data());
    if (rc != AGP_OK) return rc;
    // summand programs -> device
    const size_t off_b = align_up(comp_instr.size() * sizeof(AgpInstr), 16);
    const size_t need_b = align_up(off_b + comp_off.size() * 4, 16);
    const size_t comp_bytes = need_b + comp_need.size() * 4;
    if ((rc = grow_device(h, &h->d_comp, &h->cap_comp, comp_bytes)) != AGP_OK) return rc;
    {
        std::vector<unsigned char> img(comp_bytes, 0);  // pageable: staged before cudaMemcpyAsync returns
        memcpy(img.data(), comp_instr.data(), comp_instr.size() * sizeof(AgpInstr));
        memcpy(img.data() + off_b, comp_off.data(), comp_off.size() * 4);
        memcpy(img.data() + need_b, comp_need.data(), comp_need.size() * 4);
        AGP_CUDA(h, cudaMemcpyAsync(h->d_comp, img.data(), comp_bytes, cudaMemcpyHostToDevice, h->stream));
        AGP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    h->comp.prog = reinterpret_cast<const AgpInstr*>(h->d_comp);
    h->comp.off = reinterpret_cast<const int*>(h->d_comp + off_b);
    h->comp.need = reinterpret_cast<const int*>(h->d_comp + need_b);
    h->comp.M = (m > 0) ? M : 0;
    h->comp.m_each = m;
    const size_t mean_bytes = (size_t)P * mt * 8, cov_bytes = (size_t)P * mt * mt * 8;
    if ((rc = grow_device(h, &h->d_pred, &h->cap_pred, mean_bytes + cov_bytes + 16)) != AGP_OK) return rc;
    AGP_CUDA(h, cudaMemsetAsync(h->d_res, 0, align_up((size_t)P * 8, 16) + (size_t)P * 4, h->stream));
    rc = run_fused(h);
    h->comp.M = 0;
    if (rc != AGP_OK) return rc;
    h->n_factored = -1;  // the resident factor belongs to an augmented matrix
    if (mt > 0) {
        double* d_mean = h->d_pred;
        double* d_cov = h->d_pred + (size_t)P * mt;
        agp::launch_predict_extract(h->view, P, h->view.noise + P, d_mean, d_cov, h->stream);  // noise_pred slot: zeros
        h->launches += 1;
        if ((rc = check_launch(h, "predict_extract")) != AGP_OK) return rc;
        AGP_CUDA(h, cudaMemcpyAsync(mean_out, d_mean, mean_bytes, cudaMemcpyDeviceToHost, h->stream));
        AGP_CUDA(h, cudaMemcpyAsync(cov_out, d_cov, cov_bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    std::vector<double> lml(P);
    rc = agp_lml_fetch(h, lml.data(), info_out);  // synchronises; info: LAPACK code of the training block
    h->factor_clean = false;
    if (rc != AGP_OK) return rc;
    for (int p = 0; p < P; ++p) {
        const double np_ = noise_pred ? noise_pred[p] : noise[p];  // default: noise_pred = noise (:915)
        double* cov = cov_out + (size_t)p * mt * mt;
        for (int a = 0; a < m; ++a) cov[(size_t)(M * m + a) * mt + M * m + a] += np_;
    }
    return AGP_OK;
}

static int grad_impl(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops, const int32_t* param_off, const int32_t* n_params,
                     const double* params, const double* noise, const double* ts, const double* xs, int32_t n, double* lml_out,
                     double* grad_params_out, double* grad_noise_out, int32_t* info_out, bool noise_only) {
    if (!h) return AGP_ERR_ARG;
    if (P > 0 && (!lml_out || !grad_noise_out || !info_out || !prog_len || !n_params)) return fail(h, AGP_ERR_ARG, "agp_lml_grad_batch: bad argument");
    size_t total_params = 0;
    // Kernels of up to 64 nodes and 64 parameters take the hot variant of agp_grad_kernel; a batch with a larger one takes
    // the big variant (program from global memory, parameters in windows of 64, a tape of AGP_GRAD_TAPE_BIG levels).
    bool big = false;
    int max_params = 0;
    size_t op0 = 0;
    for (int p = 0; p < P && !noise_only; ++p) {
        if (prog_len[p] < 0 || (prog_len[p] > 0 && !ops)) return fail(h, AGP_ERR_ARG, "agp_lml_grad_batch: bad program");
        if (n_params[p] > agp::AGP_GRAD_MAX_PARAMS || prog_len[p] > 64) {
            big = true;
            int tape = 0;  // what the forward sweep pushes (agp_eval.cuh: eval_program_grad)
            for (int q = 0; q < prog_len[p]; ++q) {
                const int op = ops[op0 + q];
                tape += op == 3 ? 1 : op == 4 ? 2 : op == 5 ? 3 : op == 7 ? 2 : op == 8 ? 4 : 0;
            }
            if (tape > AGP_GRAD_TAPE_BIG)
                return fail(h, AGP_ERR_PROGRAM, "agp_lml_grad_batch: particle " + std::to_string(p) + ": the kernel needs " + std::to_string(tape) +
                                                    " tape levels for its gradient, at most " + std::to_string(AGP_GRAD_TAPE_BIG) +
                                                    " are supported (about 250 nodes)");
        }
        op0 += (size_t)prog_len[p];
        max_params = std::max(max_params, (int)n_params[p]);
        total_params += (size_t)(n_params[p] > 0 ? n_params[p] : 0);
    }
    if (total_params > 0 && !grad_params_out) return fail(h, AGP_ERR_ARG, "agp_lml_grad_batch: null gradient output");
    StallTimer stall_all_("grad_impl", (size_t)n);
    int rc;
    {
        StallTimer st_("grad_impl.upload", (size_t)P);
        rc = upload_impl(h, P, prog_len, ops, param_off, n_params, params, noise, ts, xs, n, nullptr, 0, nullptr, n > 0);
    }
    if (rc != AGP_OK) return rc;
    h->trtri_only = noise_only;
    if (P == 0) return AGP_OK;
    if (n == 0) {  // empty mvnormal: score 0, no dependence on anything
        for (int p = 0; p < P; ++p) lml_out[p] = 0.0, grad_noise_out[p] = 0.0, info_out[p] = 0;
        for (size_t j = 0; j < total_params; ++j) grad_params_out[j] = 0.0;
        return AGP_OK;
    }
    const int blocks = noise_only ? agp::noise_grad_blocks_per_particle(h->view) : agp::grad_blocks_per_particle(h->view);
    const size_t partial_doubles = (size_t)P * blocks * (noise_only ? 1 : agp::AGP_GRAD_MAX_PARAMS + 1);
    if ((rc = grow_device(h, &h->d_grad, &h->cap_grad, (partial_doubles + total_params + P + 2) * 8)) != AGP_OK) return rc;
    if ((rc = run_fused(h)) != AGP_OK) return rc;
    h->n_factored = -1;  // the resident factor belongs to an augmented matrix
    double* d_gparams = h->d_grad + partial_doubles;
    double* d_gnoise = d_gparams + total_params;
    if (noise_only) agp::launch_noise_grad(h->view, P, h->d_grad, d_gnoise, h->stream);
    else h->launches += agp::launch_grad(h->view, P, h->d_param_prefix, h->d_grad, d_gparams, d_gnoise, max_params, big, h->stream) - 2;
    h->launches += 2;
    if ((rc = check_launch(h, "grad")) != AGP_OK) return rc;
    if (total_params > 0) AGP_CUDA(h, cudaMemcpyAsync(grad_params_out, d_gparams, total_params * 8, cudaMemcpyDeviceToHost, h->stream));
    AGP_CUDA(h, cudaMemcpyAsync(grad_noise_out, d_gnoise, (size_t)P * 8, cudaMemcpyDeviceToHost, h->stream));
    {
        StallTimer st_("grad_impl.fetch", (size_t)P);
        rc = agp_lml_fetch(h, lml_out, info_out);  // synchronises
    }
    h->factor_clean = false;
    if (rc != AGP_OK) return rc;
    const double nan = std::nan("");
    size_t o = 0;
    for (int p = 0; p < P; ++p) {  // gradients of a failed factorisation are undefined
        if (info_out[p] != 0) {
            grad_noise_out[p] = nan;
            for (int j = 0; j < n_params[p] && !noise_only; ++j) grad_params_out[o + j] = nan;
        }
        o += n_params[p];
    }
    return AGP_OK;
}

int agp_lml_grad_batch(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops, const int32_t* param_off, const int32_t* n_params,
                       const double* params, const double* noise, const double* ts, const double* xs, int32_t n, double* lml_out,
                       double* grad_params_out, double* grad_noise_out, int32_t* info_out) {
    return grad_impl(h, P, prog_len, ops, param_off, n_params, params, noise, ts, xs, n, lml_out, grad_params_out, grad_noise_out, info_out, false);
}

int agp_lml_grad_noise_batch(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops, const int32_t* param_off, const int32_t* n_params,
                             const double* params, const double* noise, const double* ts, const double* xs, int32_t n, double* lml_out,
                             double* grad_noise_out, int32_t* info_out) {
    return grad_impl(h, P, prog_len, ops, param_off, n_params, params, noise, ts, xs, n, lml_out, nullptr, grad_noise_out, info_out, true);
}

int64_t agp_lml_trace(agp_handle* h, int64_t* trace_out, int64_t cap_items) {
    if (!h) return AGP_ERR_ARG;
    if (!h->uploaded) return fail(h, AGP_ERR_STATE, "agp_lml_trace: no resident batch");
    if (h->P == 0 || h->view.n == 0) return 0;
    if (h->n_pred > 0) return fail(h, AGP_ERR_STATE, "agp_lml_trace: plain LML batches only");
    AGP_CUDA(h, cudaSetDevice(h->device));
    const bool fused = gram_as_items(h, 0);  // as run_fused decides
    int64_t n_items = agp_queue_build(h->P, h->view.nt, h->order, nullptr, 0) + (fused ? (int64_t)h->P * h->view.nt * (h->view.nt + 1) : 0);
    if (use_hybrid(h, 0)) {  // the segments of the hybrid schedule, items in agp_queue_build_hybrid order (default switches)
        std::vector<int4> items;
        std::vector<int> seg;
        build_queue_hybrid(h->P, h->view.nt, h->ld / TB, hybrid_width(h, h->view.nt), 0, items, seg, false, h->oz_slice_items && !h->oz_ride, true);
        n_items = (int64_t)items.size() / 2;
    }
    if (!trace_out) return n_items;
    long long* d_trace = nullptr;
    AGP_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&d_trace), (size_t)n_items * 8 * sizeof(long long)));
    cudaMemsetAsync(d_trace, 0, (size_t)n_items * 8 * sizeof(long long), h->stream);
    int rc = run_fused(h, d_trace);
    if (rc == AGP_OK) {
        int64_t m = n_items < cap_items ? n_items : cap_items;
        cudaError_t e = cudaMemcpyAsync(trace_out, d_trace, (size_t)m * 8 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(h, AGP_ERR_CUDA, std::string("agp_lml_trace: ") + cudaGetErrorString(e));
    }
    cudaStreamSynchronize(h->stream);
    cudaFree(d_trace);
    return rc == AGP_OK ? n_items : rc;
}

static int64_t export_queue(const std::vector<int4>& items, int32_t* items_out, int64_t cap);

int64_t agp_queue_build_general(int32_t P, int32_t nt, int32_t nt_total, int32_t first_row, int32_t* items_out, int64_t cap) {
    if (P < 0 || nt < 0 || nt_total < nt || first_row < 0) return AGP_ERR_ARG;
    std::vector<int4> items;
    build_queue_general(P, nt, nt_total, first_row, items);
    return export_queue(items, items_out, cap);
}

int64_t agp_queue_build_marginals(int32_t P, int32_t nt, int32_t nt_total, int32_t* items_out, int64_t cap) {
    if (P < 0 || nt < 0 || nt_total < nt) return AGP_ERR_ARG;
    std::vector<int4> items;
    build_queue_general(P, nt, nt_total, 0, items, true);
    return export_queue(items, items_out, cap);
}

int64_t agp_queue_build(int32_t P, int32_t nt, int32_t order, int32_t* items_out, int64_t cap) {
    if (P < 0 || nt < 0) return AGP_ERR_ARG;
    std::vector<int4> items;
    if (order >= 300) {  // plain schedule with the Gram units as items (lead = 8 keeps the replay tests small and strict)
        build_queue(P, nt, nt, order - 300, items);
        fuse_gram_items(P, nt, 8, items);
    } else if (order >= 200) build_queue_inverse(P, nt, 2 * nt, order - 200, items, false);  // factorisation + trtri only (noise gradient)
    else if (order >= 100) build_queue_inverse(P, nt, 2 * nt, order - 100, items);  // identity-augmented schedule, counters laid out for 2 nt
    else build_queue(P, nt, nt, order, items);
    return export_queue(items, items_out, cap);
}

int64_t agp_queue_build_gram(int32_t P, int32_t nt, int32_t order, int32_t lead, int32_t* items_out, int64_t cap) {
    if (P < 0 || nt < 0 || lead < 1) return AGP_ERR_ARG;
    std::vector<int4> items;
    build_queue(P, nt, nt, order, items);
    fuse_gram_items(P, nt, lead, items);
    return export_queue(items, items_out, cap);
}

int64_t agp_queue_build_hybrid(int32_t P, int32_t nt, int32_t width, int32_t gram_lead, int32_t augmented, int32_t* items_out, int64_t cap, int32_t* seg_out,
                               int32_t seg_cap) {
    if (P < 0 || nt < 1 || width < 1 || gram_lead < 0 || (augmented && gram_lead > 0)) return AGP_ERR_ARG;
    const bool slice_items = (augmented & 2) != 0;  // bit 1: with the SLICE items
    augmented &= 1;
    std::vector<int4> items;
    std::vector<int> seg;
    build_queue_hybrid(P, nt, augmented ? 2 * nt : nt, width, gram_lead, items, seg, augmented != 0, slice_items, true);
    if (seg_out)
        for (size_t e = 0; e < seg.size() && (int32_t)e < seg_cap; ++e) seg_out[e] = seg[e];
    return export_queue(items, items_out, cap);
}

int agp_set_hybrid(agp_handle* h, int32_t mode, int32_t width, int32_t min_nt) {
    if (!h || width < 0 || min_nt < 2) return AGP_ERR_ARG;
    h->oz_mode = mode < 0 ? -1 : (mode != 0);
    h->oz_width = width;
    h->oz_min_nt = min_nt;
    return AGP_OK;
}

int agp_hybrid_info(agp_handle* h, int32_t* active_out, int32_t* width_out, float* stage_ms4) {
    if (!h) return AGP_ERR_ARG;
    if (active_out) *active_out = h->uploaded ? (use_hybrid(h, 0) ? 1 : 0) : h->oz_mode;
    if (width_out) *width_out = hybrid_width(h, h->uploaded ? h->view.nt : 16);
    if (stage_ms4)
        for (int e = 0; e < 4; ++e) stage_ms4[e] = h->hybrid_ms[e];
    return AGP_OK;
}

// Experiment (profiles/r02_overlap_probe.txt): do the FP64 persistent kernel and the int8 update kernel share an SM well?
// Times the single-launch FP64 step of the resident batch with `ctas_per_sm` CTAs per SM, `reps` int8 update launches
// (block columns [c0, c0 + 4), writing into a scratch matrix) on a second stream, and both at once.
// ms_out: {FP64 alone, int8 alone, FP64 with int8 next to it, int8 with FP64 next to it}.
int agp_dev_overlap_probe(agp_handle* h, int32_t ctas_per_sm, int32_t variant, int32_t c0, int32_t reps, float* ms_out) {
    if (!h || !ms_out || reps < 1) return AGP_ERR_ARG;
    if (!h->uploaded || h->aug_identity || h->view.nt_total != h->view.nt || h->view.nt < c0 + 1 || c0 < 1)
        return fail(h, AGP_ERR_STATE, "agp_dev_overlap_probe: needs a resident plain batch with more than c0 block columns");
    AGP_CUDA(h, cudaSetDevice(h->device));
    const BatchView& v = h->view;
    const int P = h->P, ld = h->ld, nt = v.nt;
    int rc;
    if ((rc = grow_device(h, &h->d_S, &h->cap_S, (size_t)agp::OZ_SLICES * P * ld * ld)) != AGP_OK) return rc;
    if ((rc = grow_device(h, &h->d_rscale, &h->cap_rscale, (size_t)P * ld * 16)) != AGP_OK) return rc;
    if (!agp::make_ozaki_maps(h->d_S, ld, P, &h->ozmaps)) return fail(h, AGP_ERR_CUDA, "tensor map");
    h->oz_S = h->d_S, h->oz_ld = ld, h->oz_P = P;
    double* scratch = nullptr;
    AGP_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&scratch), (size_t)P * ld * ld * 8));
    cudaStream_t s2;
    cudaEvent_t a0, a1, b0, b1;
    AGP_CUDA(h, cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEventCreate(&a0), cudaEventCreate(&a1), cudaEventCreate(&b0), cudaEventCreate(&b1);
    AGP_CUDA(h, cudaMemsetAsync(h->d_S, 1, (size_t)agp::OZ_SLICES * P * ld * ld, h->stream));
    AGP_CUDA(h, cudaMemsetAsync(scratch, 0, (size_t)P * ld * ld * 8, h->stream));
    const int saved_ctas = h->ctas_per_sm, saved_mode = h->oz_mode;
    h->ctas_per_sm = ctas_per_sm;
    h->oz_mode = 0;
    rc = run_fused(h);  // warm-up, queue build, row scales need the Gram diagonal
    agp::launch_gramfill(v, P, 0, h->stream);
    agp::launch_ozaki_rowscale(v.L, v.mat_stride, ld, P, h->d_rscale, h->stream);
    AGP_CUDA(h, cudaStreamSynchronize(h->stream));
    int* d_err2 = nullptr;
    AGP_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&d_err2), 4));
    AGP_CUDA(h, cudaMemset(d_err2, 0, 4));
    agp::OzakiParams prm{scratch, v.mat_stride, ld, nt, P, h->d_rscale, c0, std::min(nt, c0 + 4), d_err2, h->wait_timeout_ns};
    auto int8_burst = [&]() {
        for (int r = 0; r < reps; ++r) agp::launch_ozaki_update(prm, h->ozmaps, h->num_sms, s2, variant);
    };
    for (int pass = 0; pass < 2 && rc == AGP_OK; ++pass) {
        // alone
        cudaEventRecord(a0, h->stream);
        rc = run_fused(h);
        cudaEventRecord(a1, h->stream);
        cudaStreamSynchronize(h->stream);
        cudaEventRecord(b0, s2);
        int8_burst();
        cudaEventRecord(b1, s2);
        cudaStreamSynchronize(s2);
        cudaEventElapsedTime(&ms_out[0], a0, a1);
        cudaEventElapsedTime(&ms_out[1], b0, b1);
        // together
        cudaEventRecord(b0, s2);
        cudaEventRecord(a0, h->stream);
        int8_burst();
        if (rc == AGP_OK) rc = run_fused(h);
        cudaEventRecord(a1, h->stream);
        cudaEventRecord(b1, s2);
        cudaStreamSynchronize(h->stream);
        cudaStreamSynchronize(s2);
        cudaEventElapsedTime(&ms_out[2], a0, a1);
        cudaEventElapsedTime(&ms_out[3], b0, b1);
    }
    h->ctas_per_sm = saved_ctas;
    h->oz_mode = saved_mode;
    cudaFree(scratch);
    cudaFree(d_err2);
    cudaStreamDestroy(s2);
    cudaEventDestroy(a0), cudaEventDestroy(a1), cudaEventDestroy(b0), cudaEventDestroy(b1);
    return rc;
}

int agp_gram_items(agp_handle* h, int32_t* fused_out, int32_t* lead_out) {
    if (!h) return AGP_ERR_ARG;
    if (fused_out) *fused_out = h->uploaded ? (gram_as_items(h, 0) ? 1 : 0) : h->fuse_gram;
    if (lead_out) *lead_out = gram_items_lead(h);
    return AGP_OK;
}

static int64_t export_queue(const std::vector<int4>& items, int32_t* items_out, int64_t cap) {
    const int64_t n_items = (int64_t)items.size() / 2;
    if (items_out) {
        int64_t m = n_items < cap ? n_items : cap;
        for (int64_t q = 0; q < 2 * m; ++q) {
            items_out[4 * q + 0] = items[q].x;
            items_out[4 * q + 1] = items[q].y;
            items_out[4 * q + 2] = items[q].z;
            items_out[4 * q + 3] = items[q].w;
        }
    }
    return n_items;
}

int agp_lml_batch(agp_handle* h, int32_t P, const int32_t* prog_len, const int32_t* ops, const int32_t* param_off, const int32_t* n_params,
                  const double* params, const double* noise, const double* ts, const double* xs, int32_t n, double* lml_out,
                  int32_t* info_out) {
    int rc = agp_lml_upload(h, P, prog_len, ops, param_off, n_params, params, noise, ts, xs, n);
    if (rc != AGP_OK) return rc;
    rc = agp_lml_run(h);
    if (rc != AGP_OK) return rc;
    return agp_lml_fetch(h, lml_out, info_out);
}

}  // extern "C"
