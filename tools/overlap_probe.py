"""Do the FP64 persistent kernel and the int8 update kernel share an SM well?  (agp_dev_overlap_probe)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import autogp_oracle as o  # noqa: E402
import autogp.jl_b200 as agp  # noqa: E402
from tools.dev_check import to_agp  # noqa: E402

n, P = 2048, 64
eng = agp.Engine(0)
ts, xs = o.synthetic_series(n)
parts = [o.synthetic_particle(p, "se*per+lin") for p in range(P)]
eng.upload([to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs)
ms = (C.c_float * 4)()
for ctas, variant, reps in [(2, 3, 6), (1, 3, 6), (1, 2, 6), (2, 2, 6)]:
    rc = eng._lib.agp_dev_overlap_probe(eng._h, ctas, variant, 8, reps, ms)
    print(f"FP64 kernel {ctas} CTA/SM, int8 variant {variant} x {reps} launches: rc={rc}  FP64 alone {ms[0]:.3f} ms, int8 alone {ms[1]:.3f} ms; "
          f"together: FP64 {ms[2]:.3f} ms, int8 {ms[3]:.3f} ms (sum of alone {ms[0] + ms[1]:.3f})", flush=True)
