"""Timing of the resident-batch LML run for a given (n, P); prints ms/run and stage times."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import autogp_oracle as o  # noqa: E402
import autogp.jl_b200 as agp  # noqa: E402
from tools.dev_check import to_agp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--P", type=int, default=64)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--tree", default="se*per+lin")
ap.add_argument("--check", type=int, default=1)
a = ap.parse_args()
eng = agp.Engine(0)
ts, xs = o.synthetic_series(a.n)
parts = [o.synthetic_particle(p, a.tree) for p in range(a.P)]
eng.upload([to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs)
eng.run(); eng.run(); eng.synchronize()
ms = eng.time_runs(a.reps) / a.reps
lml, info = eng.fetch()
msg = ""
for p in range(min(a.check, a.P)):
    ref = o.log_marginal_likelihood(*parts[p], ts, xs)
    msg += f" relerr[{p}]={abs(lml[p]-ref)/abs(ref):.1e}"
fl = a.P * a.n ** 3 / 3
print(f"order={os.environ.get('AGP_ORDER','3')} n={a.n} P={a.P}: {ms:.3f} ms/run {a.P/ms*1e3:.0f} LML/s "
      f"{fl/ms*1e-9:.2f} TF/s info_ok={bool(np.all(info==0))}{msg}", flush=True)
