"""Stage split of agp_lml_grad_batch at n = 2048, 64 particles without a profiler: wall time of the whole call, of the
noise-only call (no kernel-tree walk, no lauum pass) and of a plain LML step, min of 5 (developer tool, GPU box;
AGP_LIB selects an A/B build)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autogp.jl_b200 as agp
from autogp.jl_b200.workloads import synthetic_batch, synthetic_series
eng = agp.Engine(0)
n, P = 2048, 64
ts, xs = synthetic_series(n)
nodes, noises = synthetic_batch(P)
def t(fn, reps=5):
    fn(); fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
g = t(lambda: eng.lml_grad_batch(nodes, noises, ts, xs))
gn = t(lambda: eng.lml_grad_noise_batch(nodes, noises, ts, xs))
l = t(lambda: eng.lml_batch(nodes, noises, ts, xs))
print(f"{os.environ.get('AGP_LIB', 'product')}: grad {g:.3f} ms/call  noise-only {gn:.3f}  lml {l:.3f}")
