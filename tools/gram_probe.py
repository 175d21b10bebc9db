"""Cost of the kernel-tree interpreter per node type: times agp_gram_device on several trees (GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autogp.jl_b200 as agp  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    eng = agp.Engine(0)
    ts = np.random.default_rng(0).permutation(n) / (n - 1.0)
    K = torch.empty(n * n, dtype=torch.float64, device="cuda")
    stream = torch.cuda.ExternalStream(eng.stream)
    A = agp
    trees = {
        "const": A.Constant(0.5),
        "lin": A.Linear(0.1, 1.3, 0.7),
        "se": A.SquaredExponential(0.47, 0.13),
        "per": A.Periodic(0.96, 0.21, 1.1),
        "ge": A.GammaExponential(0.42, 0.58, 3.2),
        "wn": A.WhiteNoise(1.0),
        "const+const": A.Plus(A.Constant(0.5), A.Constant(0.25)),
        "se*per": A.Times(A.SquaredExponential(0.47, 0.13), A.Periodic(0.96, 0.21, 1.1)),
        "se*per+lin": A.Plus(A.Times(A.SquaredExponential(0.47, 0.13), A.Periodic(0.96, 0.21, 1.1)), A.Linear(0.1, 1.3, 0.7)),
        "cp(lin,se)": A.ChangePoint(A.Linear(0.1, 1.3, 0.7), A.SquaredExponential(0.47, 0.13), 0.5, 0.05),
        "deep": A.Plus(A.Plus(A.Plus(A.Constant(1.0), A.Constant(2.0)), A.Plus(A.Constant(1.0), A.Constant(2.0))),
                       A.Plus(A.Plus(A.Constant(1.0), A.Constant(2.0)), A.Plus(A.Constant(1.0), A.Constant(2.0)))),
    }
    for name, tree in trees.items():
        for form in (0,):
            for _ in range(2):
                eng.gram_device(tree, 0.1, ts, K.data_ptr(), form)
            eng.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record(stream)
            for _ in range(reps):
                eng.gram_device(tree, 0.1, ts, K.data_ptr(), form)
            e1.record(stream)
            eng.synchronize()
            ms = e0.elapsed_time(e1) / reps
            uniq = n * (n + 1) / 2
            print(f"{name:14s} n={n}: {ms*1e3:8.1f} us  {uniq/ms*1e-6:8.1f} Gentry/s (unique)  write {8*n*n/ms*1e-6:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
