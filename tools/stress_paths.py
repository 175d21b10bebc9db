"""Repeated-run stability of the continuation paths (gradient, predictive MVN, block-append) on the persistent kernel:
every repetition must reproduce the first one bit for bit (developer check)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import autogp_oracle as o  # noqa: E402
import autogp.jl_b200 as agp  # noqa: E402
from tools.dev_check import to_agp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
eng = agp.Engine(0)
ts, xs = o.synthetic_series(n)
parts = [o.synthetic_particle(p) for p in range(P)]
nodes, noises = [to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
m = 200
tp = np.linspace(1.0, 1.2, m)


def same(a, b):
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


first = {}
bad = {"grad": 0, "predict": 0, "append": 0}
for r in range(reps):
    lml, grads, gnoise, ginfo = eng.lml_grad_batch(nodes, noises, ts, xs)
    g = np.concatenate([np.ravel(x) for x in grads] + [gnoise, lml])
    mean, cov, pinfo = eng.predict_batch(nodes[:16], noises[:16], ts, xs, tp)
    pr = np.concatenate([np.ravel(mean), np.ravel(cov)])
    eng.upload(nodes, noises, ts, xs)
    eng.set_prefix(n // 2 + 37)
    eng.run()
    eng.fetch()
    eng.set_prefix(n)
    eng.run_append()
    ap, ainfo = eng.fetch()
    cur = {"grad": g, "predict": pr, "append": ap}
    ok_info = bool(np.all(ginfo == 0) and np.all(pinfo == 0) and np.all(ainfo == 0))
    if r == 0:
        first = cur
    for k in cur:
        if not same(cur[k], first[k]) or not ok_info:
            bad[k] += 1
print(f"n={n} P={P} reps={reps}: runs differing from the first -> {bad}", flush=True)
