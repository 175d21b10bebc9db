"""Latency floor of the batched C-ABI calls: tiny problems, host buffers in and out (developer tool, GPU box)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200.workloads import synthetic_batch, synthetic_series  # noqa: E402

eng = agp.Engine(0)
for n, P in ((8, 16), (128, 16), (128, 4), (512, 16), (1024, 16)):
    ts, xs = synthetic_series(max(n, 2))
    ts, xs = ts[:n], xs[:n]
    nodes, noises = synthetic_batch(P)
    packed = eng.pack_batch(nodes, noises)
    row = []
    for name, fn in (("lml_batch", lambda: eng.lml_batch(nodes, noises, ts, xs)), ("lml_batch_packed", lambda: eng.lml_batch_packed(packed, ts, xs)),
                     ("lml_grad_batch", lambda: eng.lml_grad_batch(nodes, noises, ts, xs)),
                     ("lml_grad_noise_batch", lambda: eng.lml_grad_noise_batch(nodes, noises, ts, xs))):
        for _ in range(5):
            fn()
        t0 = time.perf_counter()
        for _ in range(200):
            fn()
        row.append(f"{name} {(time.perf_counter() - t0) / 200 * 1e6:.0f} us")
    l0 = eng.launch_count
    eng.lml_grad_batch(nodes, noises, ts, xs)
    l1 = eng.launch_count
    eng.lml_batch(nodes, noises, ts, xs)
    l2 = eng.launch_count
    print(f"n={n} P={P}: " + ", ".join(row) + f"; launches per grad call {l1 - l0}, per lml call {l2 - l1}", flush=True)
