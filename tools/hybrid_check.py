"""Developer check of the hybrid (int8 tcgen05 + FP64) factorisation against the single-launch FP64 schedule and the
oracle: LML and factor differences at several sizes, then timings by super-column width."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import autogp_oracle as o  # noqa: E402
import autogp.jl_b200 as agp  # noqa: E402
from tools.dev_check import to_agp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="640,1024,1100,2048")
ap.add_argument("--P", type=int, default=64)
ap.add_argument("--time-n", default="2048")
ap.add_argument("--widths", default="2,3,4,6,8")
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
eng = agp.Engine(0)
trees = ["se*per+lin", "ge+per*lin"]
for n in [int(x) for x in a.sizes.split(",") if x]:
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p, trees[p % 2]) for p in range(4)]
    nodes, noises = [to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
    eng.set_hybrid(0)
    lml0, info0 = eng.lml_batch(nodes, noises, ts, xs)
    L0 = eng.factor(0)
    ref = np.array([o.log_marginal_likelihood(nd, nz, ts, xs) for nd, nz in parts[:2]])
    for W in (2, 4):
        eng.set_hybrid(1, W, 2)
        lml1, info1 = eng.lml_batch(nodes, noises, ts, xs)
        act = eng.hybrid_info()[0]
        L1 = eng.factor(0)
        dL = np.max(np.abs(L1 - L0)) / np.max(np.abs(L0))
        print(f"n={n} W={W} hybrid={act}: info {info1.tolist()} rel |lml_h - lml_fp64| = {np.max(np.abs(lml1 - lml0) / np.abs(lml0)):.2e}, "
              f"vs oracle: fp64 {np.max(np.abs(lml0[:2] - ref) / np.abs(ref)):.2e} hybrid {np.max(np.abs(lml1[:2] - ref) / np.abs(ref)):.2e}, "
              f"max |dL| / max |L| = {dL:.2e}", flush=True)
for n in [int(x) for x in a.time_n.split(",") if x]:
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p, "se*per+lin") for p in range(a.P)]
    nodes, noises = [to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
    eng.set_hybrid(0)
    eng.upload(nodes, noises, ts, xs)
    eng.run(); eng.run(); eng.synchronize()
    ms0 = eng.time_runs(a.reps) / a.reps
    st = eng.stage_times()
    lml0, _ = eng.fetch()
    print(f"n={n} P={a.P} FP64 single launch: {ms0:.3f} ms/run (Gram {st[0]:.3f} + chol {st[1]:.3f})", flush=True)
    for W in [int(x) for x in a.widths.split(",")]:
        eng.set_hybrid(1, W, 2)
        eng.upload(nodes, noises, ts, xs)
        eng.run(); eng.run(); eng.synchronize()
        ms = eng.time_runs(a.reps) / a.reps
        eng.stage_times()
        hs = eng.hybrid_info()[2]
        lml1, info1 = eng.fetch()
        print(f"n={n} P={a.P} hybrid W={W}: {ms:.3f} ms/run ({ms0 / ms:.2f}x)  stages: Gram {hs[0]:.3f}, FP64 segments {hs[1]:.3f}, int8 updates {hs[2]:.3f}, "
              f"digit planes {hs[3]:.3f}; rel diff to FP64 {np.max(np.abs(lml1 - lml0) / np.abs(lml0)):.2e} info_ok={bool(np.all(info1 == 0))}", flush=True)
