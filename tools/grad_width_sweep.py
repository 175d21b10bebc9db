"""Gradient calls by super-column width of the hybrid schedule (developer tool, GPU box):  python tools/grad_width_sweep.py n P [widths...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200.workloads import synthetic_batch, synthetic_series  # noqa: E402

n, P = int(sys.argv[1]), int(sys.argv[2])
widths = [int(w) for w in sys.argv[3:]] or [2, 3, 4, 5]
eng = agp.Engine(0)
ts, xs = synthetic_series(n)
nodes, noises = synthetic_batch(P)
for w in [0] + widths:
    eng.set_hybrid(0 if w == 0 else 1, max(w, 1), 2)
    out = []
    for fn in (eng.lml_grad_batch, eng.lml_grad_noise_batch):
        fn(nodes, noises, ts, xs)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            res = fn(nodes, noises, ts, xs)
            best = min(best, time.perf_counter() - t0)
        assert np.all(res[-1] == 0)
        out.append(best * 1e3)
    print(f"n={n} P={P} {'FP64 schedule' if w == 0 else f'hybrid W={w}'}: lml_grad_batch {out[0]:.2f} ms, lml_grad_noise_batch {out[1]:.2f} ms per call", flush=True)
