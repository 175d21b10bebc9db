"""Developer check on a GPU box: Gram + LML parity vs the oracle and quick timings."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import autogp_oracle as o  # noqa: E402
import autogp.jl_b200 as agp  # noqa: E402


def to_agp(nd):
    name = type(nd).__name__
    cls = getattr(agp, name)
    if isinstance(nd, o.LEAVES):
        return cls(**nd.__dict__)
    if isinstance(nd, o.ChangePoint):
        return cls(to_agp(nd.left), to_agp(nd.right), nd.location, nd.scale)
    return cls(to_agp(nd.left), to_agp(nd.right))


def main():
    eng = agp.Engine(0)
    base = [o.WhiteNoise(1), o.Constant(0.5), o.Linear(0.1, 1.3, 0.7), o.SquaredExponential(0.47, 0.13),
            o.GammaExponential(0.42, 0.58, 3.2), o.Periodic(0.96, 0.21, 1.1)]
    ts = np.linspace(0, 1, 100)
    worst = 0
    for b1 in base:
        for b2 in base:
            for op in (o.Plus, o.Times, lambda x, y: o.ChangePoint(x, y, 0.5, 0.95)):
                nd = op(b1, b2)
                K0 = o.compute_cov_matrix_vectorized(nd, 0.1, ts)
                K1 = eng.gram(to_agp(nd), 0.1, ts)
                err = np.max(np.abs(K0 - K1) / np.maximum(np.abs(K0), 1e-300))
                worst = max(worst, err)
    print("gram fixture sweep: worst rel err", worst)

    for n, P, tree in ((128, 4, "se+wn"), (100, 3, "se*per+lin"), (300, 5, "ge+per*lin"), (512, 8, "se*per+lin"),
                       (700, 4, "cp(lin,se)"), (2048, 4, "se*per+lin")):
        ts, xs = o.synthetic_series(n)
        parts = [o.synthetic_particle(p, tree) for p in range(P)]
        ref = np.array([o.log_marginal_likelihood(nd, nz, ts, xs) for nd, nz in parts])
        t0 = time.time()
        lml, info = eng.lml_batch([to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs)
        dt = time.time() - t0
        rel = np.max(np.abs(lml - ref) / np.abs(ref))
        print(f"n={n} P={P} {tree}: max rel err {rel:.3e} info={info.tolist()} ({dt*1e3:.1f} ms)  ref[0]={ref[0]:.6f} got[0]={lml[0]:.6f}")

    # non-PD: negative noise
    ts, xs = o.synthetic_series(200)
    lml, info = eng.lml_batch([agp.Constant(1.0), agp.SquaredExponential(0.1, 1.0)], [-2.0, 0.1], ts, xs)
    print("non-PD check:", lml, info)

    for n, P in ((512, 64), (2048, 64)):
        ts, xs = o.synthetic_series(n)
        parts = [o.synthetic_particle(p) for p in range(P)]
        eng.upload([to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs)
        eng.run()
        eng.synchronize()
        ms = eng.time_runs(5) / 5
        fl = P * n ** 3 / 3
        print(f"n={n} P={P}: {ms:.3f} ms/run  {P/ms*1e3:.0f} LML/s  {fl/ms*1e-9:.2f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
