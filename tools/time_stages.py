"""Stage times (Gram fill, Cholesky kernel) of the resident-batch LML run; prints one line."""
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
a = ap.parse_args()
eng = agp.Engine(0)
ts, xs = o.synthetic_series(a.n)
parts = [o.synthetic_particle(p, a.tree) for p in range(a.P)]
eng.upload([to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs)
eng.run(); eng.run(); eng.synchronize()
st = np.array([eng.stage_times() for _ in range(a.reps)])
ms = eng.time_runs(a.reps) / a.reps
lml, info = eng.fetch()
ref = o.log_marginal_likelihood(*parts[0], ts, xs)
print(f"lib={os.environ.get('AGP_LIB','default')} order={os.environ.get('AGP_ORDER','3')} n={a.n} P={a.P} tree={a.tree}: "
      f"{ms:.3f} ms/run  gramfill {np.median(st[:,0]):.3f} ms  chol {np.median(st[:,1]):.3f} ms  info_ok={bool(np.all(info==0))} "
      f"relerr[0]={abs(lml[0]-ref)/abs(ref):.1e}", flush=True)
