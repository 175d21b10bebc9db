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
host buffers (H2D of programs/ts/xs and D2H of the results inside the timed region).
Rank 0 prints ONE JSON line.  Nothing here reads /root/reference.
"""
import argparse
import json
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


def load_oracle():
    """The CPU oracle — only for the cpu_baseline / --impl reference legs (never on the product path)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import autogp_oracle as o
    return o


def workload(o, n, particle_ids):
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p, TREE) for p in particle_ids]
    return ts, xs, parts


def to_agp(agp, o, nd):
    cls = getattr(agp, type(nd).__name__)
    if isinstance(nd, o.LEAVES):
        return cls(**nd.__dict__)
    if isinstance(nd, o.ChangePoint):
        return cls(to_agp(agp, o, nd.left), to_agp(agp, o, nd.right), nd.location, nd.scale)
    return cls(to_agp(agp, o, nd.left), to_agp(agp, o, nd.right))


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


def update_stage_flops(n, tb=128):
    """Algorithmic flops of the trailing-update (SYRK/GEMM) stage of one n x n Cholesky with block
    width tb: off-diagonal tiles full, diagonal tiles lower half."""
    nt = (n + tb - 1) // tb
    fma = 0.0
    for k in range(nt):
        fma += (nt - k - 1) * k * tb ** 3 + k * tb ** 3 / 2
    return 2.0 * fma


def cpu_reference_rate(o, n, n_particles, threads):
    """The reference's CPU path restated (oracle port): particles in a thread pool like
    Threads.@threads (src/inference_smc_anneal_data.jl:133); returns (LML/s, seconds)."""
    from concurrent.futures import ThreadPoolExecutor

    ts, xs, parts = workload(o, n, range(n_particles))
    o.log_marginal_likelihood(*parts[0], ts[:256], xs[:256])  # warm imports / BLAS
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda pr: o.log_marginal_likelihood(pr[0], pr[1], ts, xs), parts))
    dt = time.perf_counter() - t0
    return n_particles / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    o = load_oracle()
    cores = os.cpu_count() or 1
    sample = max(cores, 8)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_rate(o, args.n, min(sample, 4), cores)
    steps = max(1, min(args.steps, 5))  # each step = `sample` particles (~2-4 s of CPU); bounded
    t_tot, done = 0.0, 0
    for _ in range(steps):
        _, dt = cpu_reference_rate(o, args.n, sample, cores)
        t_tot += dt
        done += sample
    value = done / t_tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": 1e3 * t_tot / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"n={args.n}, {args.particles} particles/GPU, tree Plus(Times(SE,Periodic),Linear), FP64",
                   "n": args.n, "particles_per_gpu": args.particles, "tree": TREE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} particles of the same workload per step, {steps} steps, thread pool of {cores} "
                                   "(NumPy Gram with one temporary per op + LAPACK dpotrf/dtrsv: the reference's Julia path restated; Julia is not installed)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--particles", type=int, default=64, help="particles per GPU (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import autogp.jl_b200 as agp
    o = load_oracle()  # workload definition + the CPU baseline leg only

    n, P = args.n, args.particles
    ts, xs, parts = workload(o, n, range(rank * P, (rank + 1) * P))
    nodes = [to_agp(agp, o, nd) for nd, _ in parts]
    noises = [nz for _, nz in parts]
    eng = agp.Engine(local_rank)
    packed = eng.pack_batch(nodes, noises)
    stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)

    # device-resident results as a torch tensor (zero copy) for the all-gather of log-weights
    eng.upload_packed(packed, ts, xs)
    lml_ptr, _ = eng.device_results()

    class _Wrap:
        __cuda_array_interface__ = {"shape": (P,), "typestr": "<f8", "data": (lml_ptr, False), "version": 2}

    lml_dev = torch.as_tensor(_Wrap(), device=f"cuda:{local_rank}")
    gathered = torch.empty(world * P, dtype=torch.float64, device=f"cuda:{local_rank}")

    def step():
        eng.run()
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, lml_dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0 + (args.steps if world > 1 else 0)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * P / (ms_per_step * 1e-3)

    # sanity: results finite and equal to what the C-ABI host path returns
    lml_res, info = eng.fetch()
    assert np.all(info == 0) and np.all(np.isfinite(lml_res)), "benchmark batch failed to factor"

    # ---- end to end through the C-ABI with host buffers (`e2e`) ----------------------------------
    e2e_steps = max(5, min(args.steps, 30))
    for _ in range(3):
        eng.lml_batch_packed(packed, ts, xs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out, _ = eng.lml_batch_packed(packed, ts, xs)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, lml_dev)
            float(gathered[0].item())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_value = world * P * e2e_steps / dt
    ld = -(-n // 128) * 128
    n_instr = int(packed[0].sum())
    h2d = 2 * ld * 8 + 2 * P * 8 + 2 * (P + 1) * 4 + P * 4 + n_instr * 48  # the packed input arena (agp_api.cu: upload_impl)
    d2h = P * 8 + P * 4

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream ----
    # One step = agp_gramfill_kernel (kernel-tree interpreter -> K tiles in HBM, issue / FP64-pipe bound)
    # followed by ONE launch of agp_chol_kernel (persistent dataflow kernel: FP64 DMMA contraction,
    # diagonal Cholesky, panel solves, forward solve, log det), which dominates.  Algorithmic work of
    # that launch = P * n^3 / 3 flops.
    gram_ms, chol_ms = [], []
    for _ in range(7):
        a, b, _c = eng.stage_times()
        gram_ms.append(a), chol_ms.append(b)
    gram_ms, chol_ms = float(np.median(gram_ms)), float(np.median(chol_ms))
    flops = P * n ** 3 / 3.0
    achieved = flops / (chol_ms * 1e-3) * 1e-12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "chol_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("n") == n and tj.get("particles") == P:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    n_entries = P * (n * (n + 1) / 2.0)
    roofline = {
        "bound": "tensor",
        "kernel": "agp_chol_kernel (persistent: FP64 DMMA contraction + potf2 + panel solve + forward solve)",
        "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
        "traffic": traffic,
        "peak_source": "measured on this pool: cuBLAS DGEMM 8192^3 sustained (profiles/r01_fp64_lib_probe.txt); "
                       "MEASURED_PEAKS.json has no FP64 entry; DMMA issue ceiling 37.2",
        "launches_per_step": 1, "avg_launch_ms": chol_ms,
        "algorithmic_flops_per_launch": flops,
        "gramfill_kernel": {"avg_launch_ms": gram_ms, "entries_per_s": n_entries / (gram_ms * 1e-3),
                            "hbm_write_GBps": 8.0 * n_entries / (gram_ms * 1e-3) * 1e-9,
                            "bound": "instruction issue / FP64 pipe (2 exp + 1 sin + 1 division per entry in FP64; ncu: FP64 pipe ~48 % busy, issue slots ~64 %), not HBM"},
        "whole_step": {"flops": world * flops, "achieved": flops / (ms_per_step * 1e-3) * 1e-12,
                       "frac": flops / (ms_per_step * 1e-3) * 1e-12 / FP64_PEAK_TFLOPS, "unit": "TFLOP/s per GPU"},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"n={n}, {P} particles/GPU, tree Plus(Times(SE,Periodic),Linear), FP64 (BASELINE.json configs[1])",
                   "n": n, "particles_per_gpu": P, "tree": TREE, "parallelism": f"particles sharded x{world}",
                   "l2": f"no flush: working set {P * ld * ld * 8 / 1e9:.2f} GB of factors per GPU exceeds the 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = max(2 * cores, 16)
        rate, secs = cpu_reference_rate(o, n, sample, cores)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{sample} particles of the same workload ({secs:.1f} s), thread pool of {cores}: NumPy Gram "
                                          "(one temporary per op, like eval_cov) + LAPACK dpotrf/dtrsv"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
