"""Looks for sporadic long calls in the call pattern of the SMC loop (developer tool, GPU box)."""
import gc
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200.workloads import synthetic_batch, synthetic_series  # noqa: E402

eng = agp.Engine(0)
ts, xs = synthetic_series(2048)
nodes, noises = synthetic_batch(64)
if os.environ.get("RESERVE"):
    eng.reserve(2048, 64, gradient=True)
if os.environ.get("NOGC"):
    gc.disable()
log = []
for rnd in range(2):
    for n in range(205, 2049, 205):
        for name, fn, P in (("lml", eng.lml_batch, 64), ("grad", eng.lml_grad_batch, 64), ("grad", eng.lml_grad_batch, 15), ("noise", eng.lml_grad_noise_batch, 15),
                            ("grad", eng.lml_grad_batch, 9), ("grad", eng.lml_grad_batch, 64)):
            for rep in range(3):
                t0 = time.perf_counter()
                fn(nodes[:P], noises[:P], ts[:n], xs[:n])
                log.append((time.perf_counter() - t0, rnd, n, name, P, rep))
by = {}
for dt, rnd, n, name, P, rep in log:
    by.setdefault((n, name, P), []).append(dt)
print("calls", len(log), "total", round(sum(l[0] for l in log), 3), "s; sum of per-shape minima x count", round(sum(min(v) * len(v) for v in by.values()), 3), "s")
for dt, rnd, n, name, P, rep in sorted(log, reverse=True)[:12]:
    print(f"  {dt * 1e3:8.1f} ms  pass {rnd} n={n} {name} P={P} rep {rep}   (best of shape {min(by[(n, name, P)]) * 1e3:.1f} ms)")
