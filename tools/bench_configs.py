"""Measurements for the other BASELINE.json configs and the §8(f) rows (GPU box).  One JSON line each.

  configs[2]  n=8192, 64 particles (device-resident run)
  configs[3]  data-annealing schedule linear_schedule(2048, 0.10), 64 particles on this GPU:
              per round full re-score (reference behaviour) vs block-append continuation
  configs[4]  online shape: n = 256 -> 4096 step 256, 32 particles: full re-score vs block-append
  f-1         LML + gradient, n=2048 (and 512), 64 particles (end to end through the C-ABI)
  f-3         predictive MVN, n=2048, m=256, 64 particles (end to end)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import autogp_oracle as o  # noqa: E402  (workload definition only)
import autogp.jl_b200 as agp  # noqa: E402
from tools.dev_check import to_agp  # noqa: E402

PEAK = 35.4


def emit(d):
    print(json.dumps(d), flush=True)


def batch(P, tree="se*per+lin"):
    parts = [o.synthetic_particle(p, tree) for p in range(P)]
    return [to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]


def main():
    eng = agp.Engine(0)
    # ---- resident-run throughput at the three sizes
    for n, P, reps in ((512, 64, 50), (2048, 64, 20), (8192, 64, 3)):
        ts, xs = o.synthetic_series(n)
        nodes, noises = batch(P)
        eng.upload(nodes, noises, ts, xs)
        eng.run(); eng.synchronize()
        ms = eng.time_runs(reps) / reps
        g, c, _ = eng.stage_times()
        _, info = eng.fetch()
        fl = P * n ** 3 / 3
        emit({"what": "lml resident run", "n": n, "particles": P, "ms_per_step": ms, "lml_per_s": P / ms * 1e3,
              "tflops": fl / ms * 1e-9, "frac_of_fp64_peak": fl / ms * 1e-9 / PEAK, "gramfill_ms": g, "chol_kernel_ms": c,
              "chol_kernel_tflops": fl / c * 1e-9, "info_ok": bool(np.all(info == 0))})

    # ---- configs[0] (n = 128, 4 particles, Plus(SE, WN): the reference's own CPU-runnable case) and the small-n regime:
    # latency of one end-to-end call with host buffers (launch + copies + one tile per particle), and the same with a
    # batch large enough to fill the GPU
    for n, P, tree in ((128, 4, "se+wn"), (128, 64, "se*per+lin"), (128, 1024, "se*per+lin"), (512, 256, "se*per+lin")):
        ts, xs = o.synthetic_series(n)
        nodes, noises = batch(P, tree)
        packed = eng.pack_batch(nodes, noises)
        eng.lml_batch_packed(packed, ts, xs)
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            lml, info = eng.lml_batch_packed(packed, ts, xs)
        dt = (time.perf_counter() - t0) / reps
        emit({"what": "lml end-to-end call (host buffers, programs packed once)", "n": n, "particles": P, "tree": tree,
              "us_per_call": dt * 1e6, "lml_per_s": P / dt, "info_ok": bool(np.all(info == 0))})

    # ---- prefix-growing schedules: full re-score vs block-append
    def schedule_run(name, n_full, P, prefixes):
        ts, xs = o.synthetic_series(n_full)
        nodes, noises = batch(P)
        eng.upload(nodes, noises, ts, xs)
        res = {}
        for mode in ("recompute", "append"):
            for timed in (False, True):   # first pass warms every per-shape work queue and workspace
                eng.set_prefix(prefixes[0]); eng.run(); eng.fetch()
                eng.synchronize()
                t0 = time.perf_counter()
                out = []
                for i, n in enumerate(prefixes):
                    eng.set_prefix(n)
                    if mode == "append" and i > 0:
                        eng.run_append()
                    else:
                        eng.run()
                    lml, info = eng.fetch()
                    out.append(lml.copy())
                res[mode] = (time.perf_counter() - t0, out)
        same = all(np.array_equal(a, b) for a, b in zip(res["recompute"][1], res["append"][1]))
        emit({"what": name, "n_full": n_full, "particles": P, "rounds": len(prefixes), "prefixes": prefixes[:3] + ["..."] + prefixes[-1:],
              "recompute_s": res["recompute"][0], "append_s": res["append"][0], "speedup": res["recompute"][0] / res["append"][0],
              "rounds_per_s_append": len(prefixes) / res["append"][0], "particle_lml_per_s_append": P * len(prefixes) / res["append"][0],
              "bitwise_equal": bool(same)})

    schedule_run("configs[3] data annealing linear_schedule(2048,0.10), one GPU shard", 2048, 64, o.linear_schedule(2048, 0.10))
    schedule_run("configs[4] online n=256..4096 step 256", 4096, 32, list(range(256, 4097, 256)))

    # ---- gradient
    for n, P in ((512, 64), (2048, 64)):
        ts, xs = o.synthetic_series(n)
        nodes, noises = batch(P)
        eng.lml_grad_batch(nodes, noises, ts, xs)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            lml, grads, gnoise, info = eng.lml_grad_batch(nodes, noises, ts, xs)
        dt = (time.perf_counter() - t0) / reps
        emit({"what": "f-1 lml + gradient, end to end (host buffers)", "n": n, "particles": P, "ms_per_call": dt * 1e3,
              "grads_per_s": P / dt, "algorithmic_tflops (n^3: chol + trtri + lauum)": P * n ** 3 / dt * 1e-12, "info_ok": bool(np.all(info == 0))})

        eng.lml_grad_noise_batch(nodes, noises, ts, xs)
        t0 = time.perf_counter()
        for _ in range(reps):
            lml, gnoise, info = eng.lml_grad_noise_batch(nodes, noises, ts, xs)
        dt = (time.perf_counter() - t0) / reps
        emit({"what": "f-1 lml + noise gradient only, end to end (host buffers)", "n": n, "particles": P, "ms_per_call": dt * 1e3,
              "grads_per_s": P / dt, "algorithmic_tflops (2 n^3 / 3: chol + trtri)": P * 2 * n ** 3 / 3 / dt * 1e-12,
              "info_ok": bool(np.all(info == 0))})

    # ---- a9 / f-4: one round of rejuvenate_particle_parameters for all particles in lock step (L = 10 + 10 leapfrog steps)
    from autogp.jl_b200 import rejuvenate as rj
    n, P = 2048, 64
    ts, xs = o.synthetic_series(n)
    nodes, noises = batch(P)
    ch = rj.Chains(list(nodes), np.array([agp.untransform_param("noise", nz - agp.JITTER) for nz in noises]))
    rj.refresh(ch, ts, xs, eng)
    c0 = ch.n_calls
    t0 = time.perf_counter()
    acc, trials = rj.rejuvenate_parameters_lockstep(ch, np.arange(P), 1, ts, xs, rngs=rj.particle_rngs(0, P), engine=eng)
    dt = time.perf_counter() - t0
    emit({"what": "a9 lock-step parameter rejuvenation, one HMC round (params L=10, noise L=10), wall", "n": n, "particles": P,
          "ms": dt * 1e3, "batched_calls": ch.n_calls - c0, "noise_only_calls": ch.n_noise_only_calls,
          "accepted": int(sum(acc.values())), "leapfrog_evaluations_per_s": P * (ch.n_calls - c0) / dt})

    # ---- predictive MVN
    n, m, P = 2048, 256, 64
    ts, xs = o.synthetic_series(n)
    nodes, noises = batch(P)
    tp = np.linspace(1.0, 1.2, m)
    eng.predict_batch(nodes, noises, ts, xs, tp)
    t0 = time.perf_counter()
    for _ in range(3):
        mean, cov, info = eng.predict_batch(nodes, noises, ts, xs, tp)
    dt = (time.perf_counter() - t0) / 3
    emit({"what": "f-3 predictive MVN, end to end (host buffers)", "n": n, "m": m, "particles": P, "ms_per_call": dt * 1e3,
          "predictions_per_s": P / dt, "d2h_bytes": int(mean.nbytes + cov.nbytes), "info_ok": bool(np.all(info == 0))})
    # ---- infer_gp_sum: joint posterior of the two summands of the bench tree + the observable at 128 points
    n, m, P = 2048, 128, 64
    ts, xs = o.synthetic_series(n)
    tp = np.linspace(1.0, 1.2, m)
    nodes, noises = batch(P)
    sums = [[nd.left, nd.right] for nd in nodes]      # Plus(Times(SE, PER), LIN) = (SE x PER) + LIN
    eng.predict_sum_batch(sums, noises, ts, xs, tp)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        mean, cov, info = eng.predict_sum_batch(sums, noises, ts, xs, tp)
    dt = (time.perf_counter() - t0) / reps
    emit({"what": "f-3 infer_gp_sum (2 summands + observable), end to end (host buffers)", "n": n, "m": m, "particles": P,
          "ms_per_call": dt * 1e3, "posteriors_per_s": P / dt, "d2h_bytes": int(mean.nbytes + cov.nbytes), "info_ok": bool(np.all(info == 0))})
    eng.close()

if __name__ == "__main__":
    main()
