"""Parity pinned to the reference itself — when its output is on file.

``tests/golden/make_golden.jl`` runs inside an AutoGP.jl environment (Julia is not installed in this image, so the file
it writes cannot be produced here) and stores what the reference computes on its own test fixtures: Gram matrices of
the 114 kernels of ``test/test_GP.jl:24-33, 54-56``, ``Gen.logpdf(mvnormal, ...)`` values, the
``experiment_hmc.jl:180-184`` benchmarks, ``logpdf_grad`` and the predictive distribution of ``src/GP.jl:731-758``.
With ``tests/golden/julia_golden.json`` present these tests compare the CPU oracle (not gpu) and the CUDA path (gpu)
with it, at the tolerances of tests/test_gpu_parity.py; without it they are skipped, and parity stays "unpinned"
(DESIGN.md §5)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "julia_golden.json")
EPS = np.finfo(np.float64).eps

needs_file = pytest.mark.skipif(
    not os.path.exists(PATH),
    reason="tests/golden/julia_golden.json not on file: run `julia --project=<AutoGP.jl> tests/golden/make_golden.jl` "
           "(parity stays pinned to the oracle only)")


def load():
    with open(PATH) as f:
        return json.load(f)


def oracle_kernels(o):
    base = [o.WhiteNoise(1), o.Constant(0.5), o.Linear(0.1, 1.3, 0.7), o.SquaredExponential(0.47, 0.13),
            o.GammaExponential(0.42, 0.58, 3.2), o.Periodic(0.96, 0.21, 1.1)]
    out = list(base)
    for b1 in base:
        for b2 in base:
            out += [o.Plus(b1, b2), o.Times(b1, b2), o.ChangePoint(b1, b2, 0.5, 0.95)]
    return out


def oracle_benchmarks(o):
    return [(o.SquaredExponential(2.0), 0.01), (o.Plus(o.Linear(0.5), o.Periodic(2.0, 1.0)), 0.05),
            (o.ChangePoint(o.Linear(0.5), o.Linear(1.5), 1.0, 0.001), 0.001)]


def gram_close(K, K_ref):
    """|dK| <= 64 eps |K|, with the absolute floor next to a saturating ChangePoint (tests/test_gpu_parity.py)."""
    tol = 64 * EPS * np.abs(K_ref) + 8 * EPS * np.max(np.abs(K_ref))
    return np.all(np.abs(K - K_ref) <= tol)


@needs_file
def test_oracle_matches_the_reference_values():
    import autogp_oracle as o

    g = load()
    ts, xs, noise = np.array(g["ts"]), np.array(g["xs"]), g["noise"]
    for j, k in enumerate(oracle_kernels(o)):
        assert gram_close(o.eval_cov(k, ts), np.array(g["gram"][j])), f"eval_cov, kernel {j}"
        assert gram_close(o.compute_cov_matrix_vectorized(k, noise, ts), np.array(g["gram_noise"][j])), f"vectorized, kernel {j}"
        assert gram_close(o.compute_cov_matrix(k, noise, ts), np.array(g["gram_scalar"][j])), f"scalar, kernel {j}"
        ref = g["lml"][j]
        assert abs(o.log_marginal_likelihood(k, noise, ts, xs) - ref) <= 1e-8 * abs(ref), f"lml, kernel {j}"
    ts_h, xs_h = np.array(g["hmc_ts"]), np.array(g["hmc_xs"])
    for j, (k, nz) in enumerate(oracle_benchmarks(o)):
        ref = g["hmc_lml"][j]
        assert abs(o.log_marginal_likelihood(k, nz + o.JITTER, ts_h, xs_h) - ref) <= 1e-8 * abs(ref)
        _, gn = o.lml_grad_dense_fd(k, nz + o.JITTER, ts_h, xs_h)
        assert abs(gn - g["hmc_grad_noise"][j]) <= 2e-6 * abs(g["hmc_grad_noise"][j])
    k2, nz2 = oracle_benchmarks(o)[1]
    mu, cov = o.predictive_mvn(k2, nz2 + o.JITTER, ts_h[:30], xs_h[:30], ts_h[30:])
    assert np.max(np.abs(mu - np.array(g["pred_mean"]))) <= 1e-8
    assert np.max(np.abs(cov - np.array(g["pred_cov"]))) <= 1e-8
    assert abs(g["pred_logpdf"] - g["pred_logpdf_bayes"]) <= 1.5e-8 * abs(g["pred_logpdf"])  # the reference's own identity


@needs_file
@pytest.mark.gpu
def test_cuda_path_matches_the_reference_values(engine):
    import autogp.jl_b200 as agp
    import autogp_oracle as o
    from helpers import to_agp

    g = load()
    ts, xs, noise = np.array(g["ts"]), np.array(g["xs"]), g["noise"]
    kernels = [to_agp(k) for k in oracle_kernels(o)]
    for j, k in enumerate(kernels):
        assert gram_close(engine.gram(k, 0.0, ts), np.array(g["gram"][j])), f"eval_cov, kernel {j}"
        assert gram_close(engine.gram(k, noise, ts), np.array(g["gram_noise"][j])), f"vectorized, kernel {j}"
        assert gram_close(engine.gram(k, noise, ts, form=agp.gp.FORM_SCALAR), np.array(g["gram_scalar"][j])), f"scalar, kernel {j}"
    lml, info = engine.lml_batch(kernels, [noise] * len(kernels), ts, xs)
    assert np.all(info == 0)
    ref = np.array(g["lml"])
    assert np.all(np.abs(lml - ref) <= 1e-8 * np.abs(ref))
    ts_h, xs_h = np.array(g["hmc_ts"]), np.array(g["hmc_xs"])
    bm = oracle_benchmarks(o)
    nodes, noises = [to_agp(k) for k, _ in bm], [nz + o.JITTER for _, nz in bm]
    lml, info = engine.lml_batch(nodes, noises, ts_h, xs_h)
    ref = np.array(g["hmc_lml"])
    assert np.all(info == 0) and np.all(np.abs(lml - ref) <= 1e-8 * np.abs(ref))
    _, _, gnoise, ginfo = engine.lml_grad_batch(nodes, noises, ts_h, xs_h)
    gref = np.array(g["hmc_grad_noise"])
    assert np.all(ginfo == 0) and np.all(np.abs(gnoise - gref) <= 1e-8 * np.abs(gref))
    mean, cov, pinfo = engine.predict_batch(nodes[1:2], noises[1:2], ts_h[:30], xs_h[:30], ts_h[30:])
    assert pinfo[0] == 0
    assert np.max(np.abs(mean[0] - np.array(g["pred_mean"]))) <= 1e-8
    assert np.max(np.abs(cov[0] - np.array(g["pred_cov"]))) <= 1e-8
