"""Wall time of the whole structure-learning loop (smc.run_smc_anneal_data: the reference's run_smc_anneal_data,
src/inference_smc_anneal_data.jl:143-273, with its own tree proposals) on one GPU: how much of it is inside the batched
C-ABI calls and how much is host bookkeeping (proposals, latents, program encoding).

    python tools/fit_time.py --n 1024 --P 64 --n_mcmc 5 --n_hmc 2 --percent 0.1
"""
import argparse
import os
import sys
import time
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200 import smc, tree_moves as tm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--P", type=int, default=64)
ap.add_argument("--n_mcmc", type=int, default=5)
ap.add_argument("--n_hmc", type=int, default=2)
ap.add_argument("--percent", type=float, default=0.1)
ap.add_argument("--max_depth", type=int, default=-1)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--eps", type=float, default=0.02, help="HMC step size (eps_param = eps_noise; reference default .02)")
ap.add_argument("--show_failures", type=int, default=0, help="print this many proposals the engine could not score")
a = ap.parse_args()

rng = np.random.default_rng(4)
ts = rng.permutation(np.arange(a.n) / (a.n - 1))
xs = 0.8 * np.sin(2 * np.pi * ts / 0.125) + 0.5 * ts + 0.05 * rng.standard_normal(a.n)


class Timed:
    """Engine proxy: wall time and call counts of every batched call."""

    def __init__(self, eng):
        self.eng, self.t, self.calls, self.evals, self.failures = eng, Counter(), Counter(), Counter(), []
        self.log = []

    def __getattr__(self, name):
        f = getattr(self.eng, name)
        if name not in ("lml_batch", "lml_grad_batch", "lml_grad_noise_batch"):
            return f

        def wrapped(nodes, *args, **kw):
            t0 = time.perf_counter()
            out = f(nodes, *args, **kw)
            self.t[name] += time.perf_counter() - t0
            if os.environ.get("FIT_TIME_CALLS") and time.perf_counter() - t0 > 0.1:   # a stall: is it the input or the moment?
                again = []
                for _ in range(2):
                    t1 = time.perf_counter()
                    f(nodes, *args, **kw)
                    again.append(round(1e3 * (time.perf_counter() - t1), 1))
                info = out[-1]
                print(f"stall: {name} batch {len(nodes)} n {len(args[1])}: {1e3 * (time.perf_counter() - t0):.0f} ms incl. two repeats of {again} ms; "
                      f"info != 0 for {int(np.count_nonzero(info))} particles; hybrid {self.eng.hybrid_info()[:2]}", flush=True)
                import pickle
                pickle.dump((name, [repr(nd) for nd in nodes], list(args[0]), len(args[1])), open(os.path.join(ROOT, "gpurun_out", f"stall_{len(self.log)}.pkl"), "wb"))
            self.log.append((time.perf_counter() - t0, name, len(nodes), len(args[1]), max(agp.size(nd) for nd in nodes)))
            self.calls[name] += 1
            self.evals[name] += len(nodes)
            if name == "lml_batch" and len(self.failures) < a.show_failures:
                lml, info = out
                for i in np.nonzero((info != 0) | ~np.isfinite(lml))[0]:
                    self.failures.append((len(args[1]), int(info[i]), float(lml[i]), float(args[0][i]), nodes[i]))
            return out

        return wrapped


eng = Timed(agp.Engine(0))
cfg = tm.GPConfig(max_depth=a.max_depth)
rounds = []
t_last = [time.perf_counter()]


def cb(**kw):
    now = time.perf_counter()
    rounds.append((kw["step"], now - t_last[0], kw["resampled"], kw["stats"]))
    t_last[0] = now


t0 = time.perf_counter()
state = smc.run_smc_anneal_data(ts, xs, config=cfg, n_particles=a.P, n_mcmc=a.n_mcmc, n_hmc=a.n_hmc,
                                schedule=smc.linear_schedule(a.n, a.percent), seed=a.seed, engine=eng, callback_fn=cb,
                                hmc_config={"eps_param": a.eps, "eps_noise": a.eps})
total = time.perf_counter() - t0
in_calls = sum(eng.t.values())
print(f"n={a.n} P={a.P} n_mcmc={a.n_mcmc} n_hmc={a.n_hmc} schedule {a.percent}: {total:.2f} s wall, "
      f"{in_calls:.2f} s ({100 * in_calls / total:.0f} %) inside the batched calls, {total - in_calls:.2f} s host")
for k in eng.t:
    print(f"  {k:22s} {eng.calls[k]:5d} calls {eng.evals[k]:7d} particle evaluations {eng.t[k]:8.2f} s "
          f"({1e3 * eng.t[k] / eng.calls[k]:.2f} ms per call, mean batch {eng.evals[k] / eng.calls[k]:.1f})")
for step, dt, res, st in rounds[1:]:
    print(f"  prefix {step:5d}: {dt:6.2f} s resampled={res} {st}")
w = smc.compute_particle_weights(state.log_weights)
best = int(np.argmax(w))
print(f"log_ml_est {state.log_ml_est:.3f}; heaviest particle (w = {w[best]:.3f}, LML {state.scores[best]:.2f}, noise {state.noises[best]:.4g}): {state.nodes[best]}")
print("sizes of the final kernels:", sorted(agp.size(nd) for nd in state.nodes))
for n_obs, info, lml, noise, nd in eng.failures[:a.show_failures]:
    print(f"  unscoreable proposal at n = {n_obs}: info {info} lml {lml} noise {noise:.4g} size {agp.size(nd)}: {nd}")
if os.environ.get("FIT_TIME_CALLS"):
    print("slowest calls (ms, call, batch, n, largest kernel):", [(round(1e3 * d, 1), nm, P, n, sz) for d, nm, P, n, sz in sorted(eng.log, reverse=True)[:25]])
