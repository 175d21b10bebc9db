"""Per-kernel split of one agp_lml_grad_batch call at n = 2048, 64 particles: run under
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/grad_split.py
(one warm call, one measured call; the launch list then holds both)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autogp.jl_b200 as agp
from autogp.jl_b200.workloads import synthetic_batch, synthetic_series
eng = agp.Engine(0)
n, P = 2048, 64
ts, xs = synthetic_series(n)
nodes, noises = synthetic_batch(P)
for _ in range(2):
    eng.lml_grad_batch(nodes, noises, ts, xs)
eng.close()
