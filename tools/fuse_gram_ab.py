"""A/B of the Gram units as queue items against the Gram fill as its own launch (developer tool, GPU box).

    python tools/fuse_gram_ab.py n P [lead ...]

Every variant is a fresh engine (AGP_FUSE_GRAM / AGP_GRAM_LEAD are read by agp_create); the LMLs of all variants must be
bitwise the same (the units run the same code either way)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200.workloads import synthetic_batch, synthetic_series  # noqa: E402


def run(n, P, fuse, lead, reps):
    os.environ["AGP_FUSE_GRAM"] = "1" if fuse else "0"
    os.environ["AGP_GRAM_LEAD"] = str(lead)
    eng = agp.Engine(0)
    ts, xs = synthetic_series(n)
    nodes, noises = synthetic_batch(P)
    eng.upload(nodes, noises, ts, xs)
    for _ in range(3):
        eng.run()
    eng.synchronize()
    ms = min(eng.time_runs(reps) / reps for _ in range(3))
    lml, info = eng.fetch()
    st = eng.stage_times()
    eng.close()
    return ms, lml.copy(), info.copy(), st


def main():
    n, P = int(sys.argv[1]), int(sys.argv[2])
    leads = [int(a) for a in sys.argv[3:]] or [0]
    reps = 50 if n <= 512 else 20 if n <= 1024 else 10 if n <= 2048 else 2
    ms0, lml0, info0, st0 = run(n, P, False, 0, reps)
    print(f"n={n} P={P} reps={reps}: own launch {ms0:.4f} ms/step (stages: fill {st0[0]:.3f} + chol {st0[1]:.3f}) info_ok={bool(np.all(info0 == 0))}", flush=True)
    for lead in leads:
        ms, lml, info, st = run(n, P, True, lead, reps)
        same = np.array_equal(lml, lml0) and np.array_equal(info, info0)
        print(f"  gram items lead={lead or 'default'}: {ms:.4f} ms/step ({100 * (ms / ms0 - 1):+.1f} %)  bitwise_same={same}", flush=True)


if __name__ == "__main__":
    main()
