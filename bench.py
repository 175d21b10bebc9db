#!/usr/bin/env python
"""Benchmark of the GP log-marginal-likelihood hot path (BASELINE.json metric:
particle-LMLs/sec at n=2048 x 64 particles, depth-3 Sum(Product(SE,Periodic),Linear) tree, FP64).

  python bench.py [--gpus N --steps K --warmup W]            our CUDA path (one rank per GPU)
  python bench.py --impl reference [...]                     the reference's CPU path, restated
                                                             (oracle/: NumPy Gram with one temporary
                                                             per op + LAPACK dpotrf/dtrsv, all host cores)

A step = one pass of the hot path over one batch: Gram build -> Cholesky -> solve -> logdet for
`--particles` particles per GPU sharing (ts, xs) (+ the single all-gather of log-weights when
N > 1).  `value` is timed with the batch resident in HBM; `e2e` goes through the C-ABI call with
host buffers (program encoding, H2D of programs/ts/xs and D2H of the results inside the timed region).
After the headline line's numbers are taken, the rest of BASELINE.json's target is measured into the
`also` object of the same JSON line: n = 512 and n = 8192 x 64 particles, one LML-gradient and one
noise-gradient call at n = 2048, and the data-annealing schedule of configs[3] (10 prefixes of
linear_schedule(2048, 0.10), all-gather + ESS + replicated resample every round) at the launched N.
Rank 0 prints ONE JSON line.  Nothing here reads /root/reference; the oracle (the checker) is
imported only by the cpu_baseline leg and by --impl reference.
"""
import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-LMLs/sec (n×n Gram+Cholesky+solve) at n=2048, 64 particles; %FP64 roofline"
UNIT = "particle-LMLs/s"
TREE = "se*per+lin"
# FP64 dense peak measured on this pool's B200 (profiles/r01_fp64_lib_probe.txt,
# profiles/r01_fp64_probe.txt): cuBLAS DGEMM 8192^3 sustained 35.4 TFLOP/s (the convention
# MEASURED_PEAKS.json uses for bf16); the DMMA issue-rate ceiling is 37.2 TFLOP/s.
# MEASURED_PEAKS.json itself has no FP64 entry.
FP64_PEAK_TFLOPS = 35.4
KERNEL_SOURCES = ("agp_chol_kernel.cu", "agp_chol_diag.cu", "agp_chol_potf2.cu", "agp_chol_gram.cu", "agp_chol_common.cuh", "agp_ozaki.cu")
# dense int8 tensor rate measured on this pool (profiles/r02_i8_probe.txt, r02_i8_shape_probe.txt): one M128 N128 K32
# tcgen05.mma kind::i8 per 64.0 clocks and SM = 8192 MAC/clk/SM = 4117 TOP/s at the clocks of that run
INT8_PEAK_TOPS = 4117.0
INT8_PRODUCTS = 28  # digit-plane products per FP64 product (7 planes of 8 bits, p + q <= 6)


def workload_name(n, P):
    return f"n={n}, {P} particles/GPU, tree Plus(Times(SE,Periodic),Linear), FP64"


def kernel_source_md5():
    h = hashlib.md5()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "autogp.jl_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference's CPU path (oracle port): cpu_baseline leg and --impl reference ---------------------------------

def load_oracle():
    """The CPU oracle, the checker of this repository: imported ONLY by the cpu_baseline leg and by --impl reference."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import autogp_oracle as o
    return o


def cpu_reference_rate(o, n, particle_ids, workers, blas_threads):
    """The reference's CPU path restated (oracle port): particles in a thread pool like Threads.@threads
    (src/inference_smc_anneal_data.jl:133), LAPACK with `blas_threads` threads per call; returns (LML/s, seconds)."""
    from concurrent.futures import ThreadPoolExecutor

    from threadpoolctl import threadpool_limits

    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p, TREE) for p in particle_ids]
    with threadpool_limits(limits=blas_threads):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(lambda pr: o.log_marginal_likelihood(pr[0], pr[1], ts, xs), parts))
        dt = time.perf_counter() - t0
    return len(parts) / dt, dt


def cpu_fused_rate(o, n, particle_ids, workers):
    """BASELINE.md §2 context number, NOT the reference's path: fused scalar-C Gram (upper triangle, no n x n temporaries)
    + LAPACK dpotrf/dtrtrs, one particle per pool worker."""
    from concurrent.futures import ThreadPoolExecutor

    import c_oracle
    from threadpoolctl import threadpool_limits

    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p, TREE) for p in particle_ids]
    progs = [(o.encode_program(nd), nz) for nd, nz in parts]
    c_oracle.lml_cpu_best(progs[0][0], ts[:64], xs[:64], progs[0][1])
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(lambda pr: c_oracle.lml_cpu_best(pr[0], ts, xs, pr[1], 1), progs))
        dt = time.perf_counter() - t0
    return len(parts) / dt, dt


def pick_cpu_threads(o, n, cores):
    """'All the host threads it can use': the faster of (pool of `cores` workers x 1 LAPACK thread) and (the same pool x
    `cores` LAPACK threads), probed on `cores` particles each.  torchrun's OMP_NUM_THREADS=1 no longer decides it."""
    best = None
    for blas in (1, cores):
        rate, _ = cpu_reference_rate(o, n, range(cores), cores, blas)
        if best is None or rate > best[0]:
            best = (rate, blas)
    return best[1]


def run_reference(args, rank, world):
    if rank != 0:
        return
    o = load_oracle()
    cores = os.cpu_count() or 1
    n, P = args.n, args.particles
    o.log_marginal_likelihood(*o.synthetic_particle(0, TREE), *[a[:256] for a in o.synthetic_series(n)])  # warm imports / BLAS
    blas = pick_cpu_threads(o, n, cores)
    # a step = the same batch as the CUDA arm's: P particles sharing (ts, xs).  Bounded: the probe above gives the rate,
    # and the number of steps is capped so that the timed region stays under ~150 s of CPU (the JSON line says so).
    rate0, _ = cpu_reference_rate(o, n, range(min(P, cores)), cores, blas)
    est_step = P / rate0
    steps = max(1, min(args.steps, int(150.0 / max(est_step, 1e-3))))
    warmup = 1 if args.warmup > 0 else 0
    for _ in range(warmup):
        cpu_reference_rate(o, n, range(P), cores, blas)
    t_tot = 0.0
    for _ in range(steps):
        _, dt = cpu_reference_rate(o, n, range(P), cores, blas)
        t_tot += dt
    value = P * steps / t_tot
    sample = (f"{P} particles per step (the CUDA arm's batch), {steps} of the {args.steps} requested steps"
              f"{' (capped: ~150 s of CPU)' if steps < args.steps else ''}, pool of {cores} workers x {blas} LAPACK thread(s): "
              "NumPy Gram with one temporary per op + LAPACK dpotrf/dtrsv = the reference's Julia path restated (Julia is not installed)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * t_tot / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, P) + " (BASELINE.json configs[1])" if n == 2048 and P == 64 else workload_name(n, P),
                   "n": n, "particles_per_gpu": P, "tree": TREE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- the CUDA arm ---------------------------------------------------------------------------------------------------

class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU path)")
        torch.cuda.set_device(self.local_rank)
        self.dev = f"cuda:{self.local_rank}"
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        import autogp.jl_b200 as agp
        from autogp.jl_b200 import smc, workloads

        self.agp, self.smc, self.wl = agp, smc, workloads
        self.eng = agp.Engine(self.local_rank)
        self.stream = torch.cuda.ExternalStream(self.eng.stream, device=self.local_rank)
        self.args = args

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def device_lml(self, P):
        """The resident batch's P log-weights in HBM as a torch tensor (zero copy)."""
        ptr, _ = self.eng.device_results()

        class _Wrap:
            __cuda_array_interface__ = {"shape": (P,), "typestr": "<f8", "data": (ptr, False), "version": 2}

        return self.torch.as_tensor(_Wrap(), device=self.dev)

    def batch(self, n, P):
        ts, xs = self.wl.synthetic_series(n)
        nodes, noises = self.wl.synthetic_batch(P, TREE, first=self.rank * P)
        return ts, xs, nodes, noises

    # -- one size: device-resident rate + stage times ---------------------------------------------------------------
    def size_point(self, n, P, steps, warmup):
        ts, xs, nodes, noises = self.batch(n, P)
        self.eng.upload(nodes, noises, ts, xs)
        for _ in range(warmup):
            self.eng.run()
        self.barrier()
        ms = self.max_over_ranks(self.eng.time_runs(steps)) / steps
        lml, info = self.eng.fetch()
        assert np.all(info == 0) and np.all(np.isfinite(lml)), f"n={n}: batch failed to factor"
        st = [self.eng.stage_times() for _ in range(3)]
        chol_ms = float(np.median([b for _, b, _ in st]))
        flops = P * n ** 3 / 3.0
        fused, lead = self.eng.gram_items()
        hyb_on, hyb_w, hyb_ms = self.eng.hybrid_info()
        extra = {}
        if hyb_on:
            extra["hybrid"] = self.hybrid_record(n, P, hyb_w, hyb_ms)
            self.eng.set_hybrid(0)
            self.eng.run()
            self.barrier()
            extra["fp64_single_launch_ms_per_step"] = self.max_over_ranks(self.eng.time_runs(max(2, steps // 2))) / max(2, steps // 2)
            self.eng.set_hybrid(-1)
        return {**extra, "workload": workload_name(n, P), "value": self.world * P / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
                "gram_units": f"items of the persistent kernel's queue (lead {lead}): one launch per step, chol_kernel_ms is that launch, timed "
                              "with an event pair and a host sync around it" if fused else "own launch in front of the persistent kernel",
                "chol_kernel_ms": chol_ms, "chol_frac_of_fp64_peak": flops / (chol_ms * 1e-3) * 1e-12 / FP64_PEAK_TFLOPS,
                "whole_step_frac_of_fp64_peak": flops / (ms * 1e-3) * 1e-12 / FP64_PEAK_TFLOPS}

    def hybrid_record(self, n, P, width, stage_ms):
        """Stage times of a hybrid step (agp_hybrid_info; each stage bracketed by events and a sync) and the int8 kernel's rate."""
        nt = -(-n // 128)
        mac = 0.0   # multiply-accumulates of the int8 updates in FP64 terms: tile (i, k) of super-column [c0, c1) over depth c0
        for c0 in range(width, nt, width):
            for k in range(c0, min(nt, c0 + width)):
                mac += (nt - k) * 128.0 * 128.0 * (c0 * 128.0)
        int8_tops = 2.0 * INT8_PRODUCTS * P * mac / (stage_ms[2] * 1e-3) * 1e-12 if stage_ms[2] > 0 else None
        return {"super_column_width": width, "gram_and_row_scales_ms": stage_ms[0], "fp64_segments_ms": stage_ms[1], "int8_updates_ms": stage_ms[2],
                "digit_planes_ms": stage_ms[3],
                "share_of_flops_on_int8": 2.0 * mac / (n ** 3 / 3.0) if n % 128 == 0 else None,
                "int8_kernel": {"kernel": "agp_ozaki_update2_kernel (tcgen05.mma kind::i8, TMEM accumulators; one CTA per unit below 24 block columns, CTA pairs with cta_group::2 from there on)",
                                "achieved": int8_tops, "peak": INT8_PEAK_TOPS, "unit": "TOP/s (int8, 28 digit-plane products per FP64 product)",
                                "frac": None if int8_tops is None else int8_tops / INT8_PEAK_TOPS,
                                "peak_source": "measured on this pool: tcgen05.mma kind::i8 issue rate with operands resident in shared memory, 8192 MAC/clk/SM "
                                               "at 1965 MHz (profiles/r02_i8_probe.txt); 2 x MEASURED_PEAKS.json's dense bf16 burst figure would be 3326",
                                "fp64_equivalent_tflops": None if int8_tops is None else int8_tops / INT8_PRODUCTS}}

    # -- gradient calls, end to end through the C-ABI ----------------------------------------------
    def grad_point(self, n, P, reps):
        ts, xs, nodes, noises = self.batch(n, P)
        out = {}
        for name, fn in (("lml_grad_batch", self.eng.lml_grad_batch), ("lml_grad_noise_batch", self.eng.lml_grad_noise_batch)):
            fn(nodes, noises, ts, xs)
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                res = fn(nodes, noises, ts, xs)
            dt = self.max_over_ranks(time.perf_counter() - t0) / reps
            assert np.all(res[-1] == 0)
            out[name] = {"ms_per_call": 1e3 * dt, "gradients_per_s": self.world * P / dt, "calls": reps}
        out["workload"] = workload_name(n, P) + "; host buffers in and out (what Gen.hmc / the noise move consume per leapfrog step)"
        return out

    # -- configs[3]: the data-annealing schedule ----------------------------------------------------------------------
    def schedule_point(self, n, P, passes):
        """10 rounds of linear_schedule(n, 0.10) (src/Schedule.jl:24-39) as run_smc_anneal_data drives them
        (src/inference_smc_anneal_data.jl:206-234): re-score every particle on the grown prefix, all-gather the
        log-weights, ESS, multinomial resampling (replicated from a shared seed) when ESS < P_total / 2.
        full: every round recomputes the factorisation (the reference's behaviour) and a resampling round re-uploads the
        permuted programs.  append: the resident factor is continued (agp_lml_run_append), no resampling (a resampled
        particle would need its parent's factor)."""
        torch, dist, smc, eng = self.torch, self.dist, self.smc, self.eng
        ts, xs = self.wl.synthetic_series(n)
        Ptot = self.world * P
        all_nodes, all_noises = self.wl.synthetic_batch(Ptot, TREE)
        sched = smc.linear_schedule(n, 0.10)
        lo = self.rank * P
        gathered = torch.empty(Ptot, dtype=torch.float64, device=self.dev)
        pinned = torch.empty(Ptot, dtype=torch.float64).pin_memory()
        result = {"workload": f"n={n}, {Ptot} particles over {self.world} GPU(s), prefixes {sched[0]}..{sched[-1]} (BASELINE.json configs[3] shape)",
                  "rounds_per_pass": len(sched)}

        def one_pass(mode, timed):
            state = smc.ParticleState(list(all_nodes), list(all_noises))
            eng.upload(state.nodes[lo:lo + P], state.noises[lo:lo + P], ts, xs)
            lml_dev = self.device_lml(P)
            t_dev = t_exch = 0.0
            resampled = 0
            self.barrier()
            t_start = time.perf_counter()
            for r, n_r in enumerate(sched):
                t0 = time.perf_counter()
                eng.set_prefix(n_r)
                if mode == "append" and r > 0:
                    eng.run_append()
                else:
                    eng.run()
                if timed == "split":
                    torch.cuda.synchronize()
                t1 = time.perf_counter()
                local, info = eng.fetch()  # D2H of this shard's scores + LAPACK info (what smc_step! needs to raise PosDefException)
                assert np.all(info == 0), "schedule benchmark: a particle failed to factor"
                if self.world > 1:
                    with torch.cuda.stream(self.stream):
                        dist.all_gather_into_tensor(gathered, lml_dev)
                        pinned.copy_(gathered, non_blocking=True)
                    self.stream.synchronize()
                    scores = pinned.numpy().copy()
                else:
                    scores = local
                state.log_weights = state.log_weights + (scores - state.scores)
                state.scores = scores
                if mode == "full" and n_r < sched[-1] and smc.maybe_resample(state, Ptot / 2, seed=1234 + r):
                    resampled += 1
                    eng.upload(state.nodes[lo:lo + P], state.noises[lo:lo + P], ts, xs)
                    lml_dev = self.device_lml(P)
                t2 = time.perf_counter()
                t_dev += t1 - t0
                t_exch += t2 - t1
            torch.cuda.synchronize()
            return time.perf_counter() - t_start, t_dev, t_exch, resampled

        for mode in ("full", "append"):
            one_pass(mode, "total")  # warm: queues for every prefix, NCCL
            tot = 0.0
            for _ in range(passes):
                dt, _, _, resampled = one_pass(mode, "total")
                tot += dt
            dt = self.max_over_ranks(tot / passes)
            _, t_dev, t_exch, _ = one_pass(mode, "split")
            result[mode] = {"ms_per_pass": 1e3 * dt, "rounds_per_s": len(sched) / dt, "particle_rounds_per_s": Ptot * len(sched) / dt,
                            "resampling_rounds": resampled, "passes": passes,
                            "split_pass": {"lml_ms": 1e3 * self.max_over_ranks(t_dev), "exchange_and_host_ms": 1e3 * self.max_over_ranks(t_exch),
                                           "note": "one extra pass with a device sync after every round's LML: lml = set_prefix + launches + kernels; "
                                                   "exchange = all-gather of the log-weights + D2H + ESS / resample decision (+ re-upload when resampled)"}}
        return result

    # -- the callers of the path: the whole data-annealing SMC loop ------------------------
    def fit_point(self, n, P, n_mcmc, n_hmc):
        """smc.run_smc_anneal_data = run_smc_anneal_data of src/inference_smc_anneal_data.jl:143-273 with the reference's
        tree prior and subtree-replace / detach-attach proposals (tree_moves.py): every reweighting step, MH proposal and
        leapfrog step of all particles is one batched C-ABI call.  Wall clock of one fit, host work included."""
        from autogp.jl_b200 import tree_moves

        rng = np.random.default_rng(4)
        ts = rng.permutation(np.arange(n) / (n - 1))
        xs = 0.8 * np.sin(2 * np.pi * ts / 0.125) + 0.5 * ts + 0.05 * rng.standard_normal(n)
        sched = self.smc.linear_schedule(n, 0.10)
        calls = {"n": 0}
        t0 = time.perf_counter()
        state = self.smc.run_smc_anneal_data(ts, xs, config=tree_moves.GPConfig(), n_particles=P, n_mcmc=n_mcmc, n_hmc=n_hmc, schedule=sched,
                                             seed=0, engine=self.eng,
                                             callback_fn=lambda **kw: calls.__setitem__("n", calls["n"] + 1))
        dt = time.perf_counter() - t0
        w = self.smc.compute_particle_weights(state.log_weights)
        return {"workload": f"n={n}, {P} particles, {len(sched)} prefixes of linear_schedule(n, .10), n_mcmc={n_mcmc}, n_hmc={n_hmc} (L = 10 + 10 leapfrog steps), "
                            "prior over kernels = GPConfig defaults (max_depth = -1), synthetic periodic + trend series",
                "seconds_per_fit": dt, "rounds": calls["n"] - 1, "log_ml_est": float(state.log_ml_est),
                "posterior_weight_on_periodic_kernels": float(sum(wi for wi, nd in zip(w, state.nodes)
                                                                  if any(type(a).__name__ == "Periodic" for a in self.agp.unroll(nd)))),
                "note": "one GPU; includes the host side (proposals, latents, program encoding) and one-time imports"}

    def run(self):
        args, torch, dist, eng = self.args, self.torch, self.dist, self.eng
        rank, world = self.rank, self.world
        n, P = args.n, args.particles
        ts, xs, nodes, noises = self.batch(n, P)
        eng.upload(nodes, noises, ts, xs)
        lml_dev = self.device_lml(P)
        gathered = torch.empty(world * P, dtype=torch.float64, device=self.dev)

        def step():
            eng.run()
            if world > 1:
                with torch.cuda.stream(self.stream):
                    dist.all_gather_into_tensor(gathered, lml_dev)

        # ---- device-resident throughput (`value`) ------------------------------------------------
        warmup = max(args.warmup, 3)
        for _ in range(warmup):
            step()
        self.barrier()
        sampler = ClockSampler(self.local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.15)
        launches0 = eng.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        for _ in range(args.steps):
            step()
        e1.record(self.stream)
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        launches = eng.launch_count - launches0 + (args.steps if world > 1 else 0)
        clocks = sampler.stop() if rank == 0 else None
        ms_per_step = ms / args.steps
        value = world * P / (ms_per_step * 1e-3)
        lml_res, info = eng.fetch()
        assert np.all(info == 0) and np.all(np.isfinite(lml_res)), "benchmark batch failed to factor"

        # ---- end to end through the public call with host buffers (`e2e`) --------------------------
        # every step: encode the kernel trees (a new set of programs per MH step in real use), one H2D of the packed
        # arena, the two launches, one D2H of the results (+ the all-gather and a host read of it when N > 1)
        e2e_steps = max(5, min(args.steps, 30))
        for _ in range(3):
            eng.lml_batch(nodes, noises, ts, xs)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out, _ = eng.lml_batch(nodes, noises, ts, xs)
            if world > 1:
                with torch.cuda.stream(self.stream):
                    dist.all_gather_into_tensor(gathered, lml_dev)
                float(gathered[0].item())
        torch.cuda.synchronize()
        dt = self.max_over_ranks(time.perf_counter() - t0)
        e2e_value = world * P * e2e_steps / dt
        ld = -(-n // 128) * 128
        packed = eng.pack_batch(nodes, noises)
        n_instr = int(packed[0].sum())
        h2d = 2 * ld * 8 + 2 * P * 8 + 2 * (P + 1) * 4 + P * 4 + n_instr * 48  # the packed input arena (agp_api.cu: upload_impl)
        d2h = P * 8 + P * 4

        # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream ----
        # One step = agp_gramfill_kernel (kernel-tree interpreter -> K tiles in HBM, issue / FP64-pipe bound)
        # followed by ONE launch of agp_chol_kernel (persistent dataflow kernel: FP64 DMMA contraction,
        # diagonal Cholesky, panel solves, forward solve, log det), which dominates.  Algorithmic work of
        # that launch = P * n^3 / 3 flops.
        eng.upload(nodes, noises, ts, xs)
        st = [eng.stage_times() for _ in range(7)]
        gram_ms, chol_ms = float(np.median([a for a, _, _ in st])), float(np.median([b for _, b, _ in st]))
        flops = P * n ** 3 / 3.0
        achieved = flops / (chol_ms * 1e-3) * 1e-12
        traffic, traffic_note = None, "no capture on file"
        tpath = os.path.join(ROOT, "profiles", "chol_kernel_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if tj.get("n") == n and tj.get("particles") == P and tj.get("kernel_source_md5") == kernel_source_md5():
                    traffic = tj.get("dram_bytes_per_launch")
                    traffic_note = tj.get("source", "")
                else:
                    traffic_note = "the ncu capture on file is of another kernel source or workload"
            except Exception:
                pass
        n_entries = P * (n * (n + 1) / 2.0)
        hyb_on, hyb_w, hyb_ms = eng.hybrid_info()
        n_launch = 1
        kernel_name = "agp_chol_kernel (persistent: FP64 DMMA contraction + potf2 + panel solve + forward solve)"
        if hyb_on:
            n_seg = -(-(-(-n // 128)) // hyb_w)
            n_launch = 3 * n_seg - 2
            kernel_name = (f"the factorisation as {n_launch} launches: {n_seg} segments of agp_chol_kernel (FP64 DMMA: potf2, panel solves, contractions inside a "
                           f"super-column of {hyb_w} block columns), {n_seg - 1} x agp_ozaki_update2_kernel (the contractions over all earlier block columns as exact "
                           f"int8 digit-plane products on tcgen05, kind::i8) and {n_seg - 1} x agp_ozaki_slice_kernel; avg_launch_ms = their sum")
            traffic, traffic_note = None, "no capture of the multi-launch factorisation on file"
            hpath = os.path.join(ROOT, "profiles", "hybrid_traffic.json")
            if os.path.exists(hpath):
                try:
                    tj = json.load(open(hpath))
                    if tj.get("n") == n and tj.get("particles") == P and tj.get("kernel_source_md5") == kernel_source_md5():
                        traffic = tj.get("dram_bytes_per_step")
                        traffic_note = tj.get("source", "")
                    else:
                        traffic_note = "the ncu capture on file is of another kernel source or workload"
                except Exception:
                    pass
        roofline = {
            "bound": "tensor",
            "kernel": kernel_name,
            "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
            "traffic": traffic, "traffic_source": traffic_note,
            "peak_source": "measured on this pool: cuBLAS DGEMM 8192^3 sustained (profiles/r01_fp64_lib_probe.txt); "
                           "MEASURED_PEAKS.json has no FP64 entry; DMMA issue ceiling 37.2",
            "launches_per_step": n_launch, "avg_launch_ms": chol_ms,
            "algorithmic_flops_per_launch": flops,
            "gramfill_kernel": {"avg_launch_ms": gram_ms, "entries_per_s": n_entries / (gram_ms * 1e-3),
                                "hbm_write_GBps": 8.0 * n_entries / (gram_ms * 1e-3) * 1e-9,
                                "bound": "instruction issue / FP64 pipe (2 exp + 1 sin + 1 division per entry in FP64; ncu: FP64 pipe ~48 % busy, issue slots ~64 %), not HBM"},
            "whole_step": {"flops": world * flops, "achieved": flops / (ms_per_step * 1e-3) * 1e-12,
                           "frac": flops / (ms_per_step * 1e-3) * 1e-12 / FP64_PEAK_TFLOPS, "unit": "TFLOP/s per GPU"},
        }

        if hyb_on:
            roofline["hybrid"] = self.hybrid_record(n, P, hyb_w, hyb_ms)
            roofline["note"] = ("peak is the FP64 tensor (DMMA) peak; a frac above the FP64 segments' share is possible because the long contractions "
                                "run on the int8 tensor path at FP64-grade accuracy (error-free digit-plane split)")
            eng.set_hybrid(0)
            eng.run()
            self.barrier()
            k = max(3, args.steps // 4)
            roofline["fp64_single_launch_ms_per_step"] = self.max_over_ranks(eng.time_runs(k)) / k
            eng.set_hybrid(-1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(n, P) + (" (BASELINE.json configs[1])" if n == 2048 and P == 64 else ""),
                       "n": n, "particles_per_gpu": P, "tree": TREE, "parallelism": f"particles sharded x{world}",
                       "l2": f"no flush: working set {P * ld * ld * 8 / 1e9:.2f} GB of factors per GPU exceeds the 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "includes": "kernel-tree encoding on the host, H2D, both launches, D2H" + (", all-gather + host read" if world > 1 else "")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
        }

        # ---- the rest of BASELINE.json's target, into the same line --------------------------------
        if not args.no_also:
            also = {}
            try:
                also["n512"] = self.size_point(512, P, steps=50, warmup=5)
                also["n8192"] = self.size_point(8192, P, steps=3, warmup=1)
                also["grad_n2048"] = self.grad_point(2048, P, reps=3)
                also["anneal_schedule_n2048"] = self.schedule_point(2048, P, passes=3)
                if world == 1:
                    also["structure_learning_n512"] = self.fit_point(512, P, n_mcmc=3, n_hmc=2)
            except Exception as e:  # the headline line must survive a failure of the extras
                also["error"] = f"{type(e).__name__}: {e}"
            line["also"] = also

        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            o = load_oracle()
            cores = os.cpu_count() or 1
            blas = pick_cpu_threads(o, n, cores)
            sample = max(2 * cores, 16)
            rate, secs = cpu_reference_rate(o, n, range(sample), cores, blas)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} particles of the same workload ({secs:.1f} s), pool of {cores} workers x {blas} LAPACK thread(s): "
                                              "NumPy Gram (one temporary per op, like eval_cov) + LAPACK dpotrf/dtrsv"}
            try:
                frate, fsecs = cpu_fused_rate(o, n, range(sample), cores)
                line["cpu_baseline"]["context_fused_c"] = {
                    "value": frate, "unit": UNIT,
                    "what": f"not the reference's path: fused scalar-C Gram without n x n temporaries + LAPACK, {sample} particles over {cores} workers "
                            f"({fsecs:.1f} s); scalar libm exp/sin lose to NumPy's SIMD loops what the saved temporaries gain"}
            except Exception as e:
                line["cpu_baseline"]["context_fused_c"] = {"error": f"{type(e).__name__}: {e}"}
        if rank == 0:
            print(json.dumps(line), flush=True)
        eng.close()
        if world > 1:
            dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--particles", type=int, default=64, help="particles per GPU (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the other sizes / gradient / schedule measurements")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    Bench(args).run()


if __name__ == "__main__":
    main()
