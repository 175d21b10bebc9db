import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import autogp_oracle as o
import autogp.jl_b200 as agp
from tools.dev_check import to_agp
eng = agp.Engine(0)
for n, P in ((512, 2), (640, 2), (768, 2), (1024, 1), (1024, 4), (2048, 1), (2048, 4), (2048, 8)):
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p, "se*per+lin") for p in range(P)]
    ref = np.array([o.log_marginal_likelihood(nd, nz, ts, xs) for nd, nz in parts])
    outs = []
    for rep in range(3):
        lml, info = eng.lml_batch([to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs)
        outs.append(lml)
    rel = [np.max(np.abs(l - ref) / np.abs(ref)) for l in outs]
    print(f"order={os.environ.get('AGP_ORDER','2')} n={n} P={P}: rel errs {['%.1e' % r for r in rel]}", flush=True)
