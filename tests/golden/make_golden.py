"""Generates the committed golden fixtures for the GP-LML hot path.

The reference (Julia) cannot run in this image and ships no golden vectors for this path, so
the fixtures are produced by the CPU oracle (oracle/autogp_oracle.py) on the reference's own
test fixtures, with an mpmath 50-digit evaluation as ground truth for the LML values:

  * kernels  : test/test_GP.jl:24-33 (6 base kernels) and :54-56 (all 36 ordered pairs under
               +, *, ChangePoint(.,.,0.5,0.95)) -> 114 trees
  * grid     : test/test_GP.jl:38-40  ds_raw = range(-10, 10, length=100) mapped to [0,1];
               every 7th point (15 points) keeps the file small
  * LML cases: test/experiment_hmc.jl:180-184 benchmark (kernel, noise) pairs

Run:  python tests/golden/make_golden.py     (writes gram_golden.npz, lml_golden.json here)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import autogp_oracle as o  # noqa: E402


def base_kernels():
    return [o.WhiteNoise(1), o.Constant(0.5), o.Linear(0.1, 1.3, 0.7), o.SquaredExponential(0.47, 0.13),
            o.GammaExponential(0.42, 0.58, 3.2), o.Periodic(0.96, 0.21, 1.1)]


def fixture_kernels():
    base = base_kernels()
    out = list(base)
    for b1 in base:
        for b2 in base:
            out.append(o.Plus(b1, b2))
            out.append(o.Times(b1, b2))
            out.append(o.ChangePoint(b1, b2, 0.5, 0.95))
    return out


def fixture_grid():
    ds_raw = np.linspace(-10.0, 10.0, 100)
    # Transforms.LinearTransform(ds_raw, 0, 1): slope = 1/20, intercept = -slope*tmin + 0
    slope = (1.0 - 0.0) / (ds_raw.max() - ds_raw.min())
    intercept = -slope * ds_raw.min() + 0.0
    ds = slope * ds_raw + intercept
    return ds_raw, ds


def fixture_xs(ts):
    return np.sin(3.0 * ts) + 0.2 * np.cos(11.0 * ts)


def hmc_benchmarks():
    return [(o.SquaredExponential(2.0), 0.01),
            (o.Plus(o.Linear(0.5), o.Periodic(2.0, 1.0)), 0.05),
            (o.ChangePoint(o.Linear(0.5), o.Linear(1.5), 1.0, 0.001), 0.001)]


def main():
    _, ds = fixture_grid()
    sub = ds[::7]
    kernels = fixture_kernels()
    grams = np.stack([o.compute_cov_matrix_vectorized(k, 0.0, sub) for k in kernels])
    np.savez_compressed(os.path.join(HERE, "gram_golden.npz"), ts=sub, grams=grams)

    xs = fixture_xs(sub)
    noise = 0.1
    lml = {"noise": noise, "fixture_mp": [], "fixture_f64": []}
    for k in kernels:
        lml["fixture_mp"].append(o.log_marginal_likelihood_mp(k, noise, sub, xs))
        lml["fixture_f64"].append(o.log_marginal_likelihood(k, noise, sub, xs))
    ts_h = np.linspace(0.0, 10.0, 1000)[::25]  # 40 of the experiment_hmc.jl:197 grid
    xs_h = fixture_xs(ts_h / 10.0)
    lml["hmc_mp"] = [o.log_marginal_likelihood_mp(k, nz + o.JITTER, ts_h, xs_h) for k, nz in hmc_benchmarks()]
    # configs[0] of BASELINE.json: n=128, 4 particles, SE + WhiteNoise (FP64 oracle values)
    ts, xs128 = o.synthetic_series(128)
    lml["config0_f64"] = [o.log_marginal_likelihood(*o.synthetic_particle(p, "se+wn"), ts, xs128) for p in range(4)]
    ts, xs256 = o.synthetic_series(256)
    lml["n256_se_per_lin_f64"] = [o.log_marginal_likelihood(*o.synthetic_particle(p), ts, xs256) for p in range(4)]
    with open(os.path.join(HERE, "lml_golden.json"), "w") as f:
        json.dump(lml, f, indent=1)
    print("wrote", len(kernels), "kernels")


def sum_fixture_inputs():
    """infer_gp_sum / noise-gradient / marginals fixture: the Plus benchmark of experiment_hmc.jl:181 read as two summands."""
    ts_h = np.linspace(0.0, 10.0, 1000)[::25]
    xs_h = fixture_xs(ts_h / 10.0)
    return [o.Linear(0.5), o.Periodic(2.0, 1.0)], 0.05 + o.JITTER, ts_h, xs_h, np.array([0.3, 4.9, 10.0, 10.5, 12.0])


def main_sum():
    """Separate file (the fixtures above are frozen): values of the oracle restatements added for SURVEY §8 f."""
    nodes, noise, ts, xs, tp = sum_fixture_inputs()
    mu, cov, _ = o.infer_gp_sum(nodes, noise, ts, xs, tp, noise_pred=0.0)
    whole = o.Plus(nodes[0], nodes[1])
    mu_p, cov_p = o.predictive_mvn(whole, noise, ts, xs, tp)
    g, gn = o.lml_grad_dense_fd(whole, noise, ts, xs)
    out = {"infer_gp_sum_mean": mu.tolist(), "infer_gp_sum_cov": cov.tolist(), "predictive_mean": mu_p.tolist(),
           "predictive_var": np.diag(cov_p).tolist(), "lml_grad_noise": gn, "lml": o.log_marginal_likelihood(whole, noise, ts, xs)}
    with open(os.path.join(HERE, "sum_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote sum_golden.json")


if __name__ == "__main__":
    import sys

    if len(sys.argv) > 1 and sys.argv[1] == "sum":
        main_sum()
    else:
        main()
