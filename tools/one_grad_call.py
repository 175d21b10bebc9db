"""Two agp_lml_grad_batch calls at n = 2048 x 64 (the second one warm) — run under `ncu --metrics gpu__time_duration.sum` for the launch
list of a gradient call on the hybrid schedule (profiles/r02_ncu_launches_grad_hybrid.csv)."""
import os, sys
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'oracle'))
import autogp_oracle as o
import autogp.jl_b200 as agp
from tools.dev_check import to_agp
eng=agp.Engine(0)
n,P=2048,64
ts,xs=o.synthetic_series(n)
parts=[o.synthetic_particle(p) for p in range(P)]
nodes,noises=[to_agp(nd) for nd,_ in parts],[nz for _,nz in parts]
eng.lml_grad_batch(nodes,noises,ts,xs)
eng.lml_grad_batch(nodes,noises,ts,xs)
