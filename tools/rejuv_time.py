"""Wall time of the lock-step rejuvenation loops (autogp.jl_b200/rejuvenate.py) at bench scale: how much of a
parameter-rejuvenation round is GPU time inside agp_lml_grad_batch and how much is host bookkeeping."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import autogp_oracle as o  # noqa: E402  (synthetic inputs only)
import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200 import rejuvenate as rj  # noqa: E402
from tools.dev_check import to_agp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--P", type=int, default=64)
ap.add_argument("--n_hmc", type=int, default=1)
ap.add_argument("--n_mcmc", type=int, default=2)
a = ap.parse_args()
eng = agp.Engine(0)
ts, xs = o.synthetic_series(a.n)
parts = [o.synthetic_particle(p) for p in range(a.P)]
# the bench tree has a SquaredExponential leaf, which the prior never draws but HMC moves like any other leaf
nodes = [to_agp(nd) for nd, _ in parts]
zn = np.array([agp.untransform_param("noise", nz - agp.JITTER) for _, nz in parts])
ch = rj.Chains(list(nodes), zn.copy())
rngs = rj.particle_rngs(0, a.P)
rj.refresh(ch, ts, xs, eng)            # warm-up: allocations, queues
t0 = time.perf_counter()
rj.refresh(ch, ts, xs, eng)
t_call = time.perf_counter() - t0
c0 = ch.n_calls
t0 = time.perf_counter()
acc, trials = rj.rejuvenate_parameters_lockstep(ch, np.arange(a.P), a.n_hmc, ts, xs, rngs=rngs, engine=eng)
t_hmc = time.perf_counter() - t0
calls = ch.n_calls - c0
print(f"n={a.n} P={a.P}: one agp_lml_grad_batch call {t_call * 1e3:.1f} ms; parameter rejuvenation n_hmc={a.n_hmc} "
      f"(params L=10 + noise L=10): {calls} batched calls ({ch.n_noise_only_calls} of them noise-gradient-only), "
      f"{t_hmc * 1e3:.1f} ms wall = {t_hmc / calls * 1e3:.1f} ms per call, accepted {sum(acc.values())}/{sum(trials.values())}; "
      f"{a.P * calls / t_hmc:.0f} particle leapfrog evaluations/s")
c0 = ch.n_calls
t0 = time.perf_counter()
stats = rj.rejuvenate_structure_lockstep(ch, a.n_mcmc, a.n_hmc, rj.leaf_swap_proposal, ts, xs, seed=1, engine=eng, rngs=rngs)
t_s = time.perf_counter() - t0
print(f"structure rejuvenation n_mcmc={a.n_mcmc} n_hmc={a.n_hmc}: {ch.n_calls - c0} batched calls, {t_s * 1e3:.1f} ms wall, stats {stats}")
