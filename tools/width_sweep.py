"""Sweep of the hybrid schedule's super-column width (and of the int8 kernel variant) at a given size (developer tool, GPU box).

    python tools/width_sweep.py n P [widths...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200.workloads import synthetic_batch, synthetic_series  # noqa: E402

n, P = int(sys.argv[1]), int(sys.argv[2])
widths = [int(w) for w in sys.argv[3:]] or [2, 3, 4, 5, 6]
eng = agp.Engine(0)
ts, xs = synthetic_series(n)
nodes, noises = synthetic_batch(P)
eng.upload(nodes, noises, ts, xs)
reps = 10 if n <= 4096 else 3
base = None
for w in widths:
    eng.set_hybrid(1, w, 2)
    for _ in range(2):
        eng.run()
    eng.synchronize()
    ms = min(eng.time_runs(reps) / reps for _ in range(3))
    lml, info = eng.fetch()
    assert np.all(info == 0)
    base = lml if base is None else base
    on, width, stages = eng.hybrid_info()
    print(f"n={n} P={P} W={w}: {ms:.3f} ms per step; stages (gram+scales, fp64 segments, int8 updates, digit planes) ms = "
          f"{', '.join(f'{s:.3f}' for s in stages)}; max |dLML| vs first width {np.max(np.abs(lml - base)):.2e}", flush=True)
