"""First use of the general gradient variant (bigger per-thread stack): how long does the first call take? (developer tool)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200.workloads import synthetic_batch, synthetic_series  # noqa: E402

eng = agp.Engine(0)
ts, xs = synthetic_series(1435)
nodes, noises = synthetic_batch(64)
eng.reserve(2048, 64, gradient=True)
big = agp.Periodic(0.5, 0.3, 0.2)
for j in range(36):
    leaf = agp.Periodic(0.5 + 0.01 * j, 0.3, 0.05) if j % 2 else agp.Linear(0.3, 0.1, 0.05)
    big = agp.Plus(big, leaf) if j % 3 else agp.Times(big, leaf)
for label, nd in (("tuned variant", nodes), ("tuned variant", nodes), ("general variant", [big] + nodes[1:]), ("general variant", [big] + nodes[1:]),
                  ("tuned variant", nodes), ("general variant", [big] + nodes[1:])):
    t0 = time.perf_counter()
    eng.lml_grad_batch(nd, noises, ts, xs)
    print(f"{label}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
