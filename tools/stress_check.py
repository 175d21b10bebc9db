"""Repeated full-size runs: info flags, finiteness, run-to-run bitwise stability (developer check)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import autogp_oracle as o
import autogp.jl_b200 as agp
from tools.dev_check import to_agp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
eng = agp.Engine(0)
ts, xs = o.synthetic_series(n)
parts = [o.synthetic_particle(p) for p in range(P)]
nodes, noises = [to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
first = None
for r in range(reps):
    lml, info = eng.lml_batch(nodes, noises, ts, xs)
    bad = np.nonzero(info)[0]
    if first is None:
        first = lml.copy()
    print(f"order={os.environ.get('AGP_ORDER','3')} n={n} run {r}: bad particles {bad.tolist()} info {info[bad].tolist()} "
          f"finite={bool(np.all(np.isfinite(lml)))} same_as_first={bool(np.array_equal(lml, first))} "
          f"max_rel_diff={np.nanmax(np.abs(lml - first) / np.abs(first)):.2e} n_diff={int(np.sum(lml != first))}", flush=True)
ref = o.log_marginal_likelihood(*parts[0], ts, xs)
print("relerr[0]", abs(first[0] - ref) / abs(ref))
