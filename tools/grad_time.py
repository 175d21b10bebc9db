"""Wall time of agp_lml_grad_batch (LML + gradient, host buffers) at n = 2048, 64 particles; run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split (Gram fill / factorisation + inverse / gradient interpreter)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'oracle'))
import autogp_oracle as o
import autogp.jl_b200 as agp
from tools.dev_check import to_agp
eng=agp.Engine(0)
n,P=(int(sys.argv[1]) if len(sys.argv)>1 else 2048),64
ts,xs=o.synthetic_series(n)
parts=[o.synthetic_particle(p) for p in range(P)]
nodes,noises=[to_agp(nd) for nd,_ in parts],[nz for _,nz in parts]
for _ in range(3): eng.lml_grad_batch(nodes,noises,ts,xs)
t0=time.perf_counter()
for _ in range(5): eng.lml_grad_batch(nodes,noises,ts,xs)
print('grad ms/call',(time.perf_counter()-t0)/5*1e3)
for _ in range(3): eng.lml_grad_noise_batch(nodes,noises,ts,xs)
t0=time.perf_counter()
for _ in range(5): eng.lml_grad_noise_batch(nodes,noises,ts,xs)
print('noise-only grad ms/call',(time.perf_counter()-t0)/5*1e3)
for _ in range(3): eng.lml_batch(nodes,noises,ts,xs)
t0=time.perf_counter()
for _ in range(5): eng.lml_batch(nodes,noises,ts,xs)
print('lml ms/call',(time.perf_counter()-t0)/5*1e3)
