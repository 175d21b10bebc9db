"""CPU suite, part 1: pin the oracle.

The reference holds no golden values for this path, so the oracle is pinned by
(a) the committed fixtures generated from the reference's own test kernels/grids,
(b) three independent routes agreeing (vectorised NumPy, scalar C, mpmath 50 digits),
(c) the identities the reference's tests assert (test/test_GP.jl:35-106,
    test/experiment_hmc.jl:111-132).
"""
import math

import numpy as np
import pytest

import autogp_oracle as o
import c_oracle
import helpers as H


def test_golden_gram_matches_oracle():
    ts, grams, _ = H.golden()
    for k, G in zip(H.fixture_kernels(), grams):
        K = o.compute_cov_matrix_vectorized(k, 0.0, ts)
        assert np.array_equal(K, G), k


def test_golden_lml_mpmath_vs_f64_routes():
    ts, _, lml = H.golden()
    xs = H.fixture_xs(ts)
    for i, k in enumerate(H.fixture_kernels()):
        a = o.log_marginal_likelihood(k, lml["noise"], ts, xs)
        b, info = c_oracle.lml(o.encode_program(k), ts, xs, lml["noise"])
        c = o.log_marginal_likelihood_lu(k, lml["noise"], ts, xs)
        assert info == 0
        truth = lml["fixture_mp"][i]
        # n = 15, well conditioned: FP64 routes agree with the 50-digit value to ~1e-13
        for v in (a, b, c):
            assert abs(v - truth) <= 1e-12 * max(1.0, abs(truth)), (k, v, truth)


def test_scalar_and_vectorised_gram_agree():
    _, ds = H.fixture_grid()
    ts = ds[::3]
    for k in H.fixture_kernels():
        Kv = o.compute_cov_matrix_vectorized(k, 0.3, ts)
        Ks = o.compute_cov_matrix(k, 0.3, ts)
        Kc0 = c_oracle.gram(o.encode_program(k), ts, 0.3, form=0)
        Kc1 = c_oracle.gram(o.encode_program(k), ts, 0.3, form=1)
        assert np.allclose(Kv, Ks, rtol=1e-14, atol=1e-300)
        # C (glibc scalar libm) vs NumPy (SIMD libm): identical formulas, <= 2 ulp transcendentals
        assert np.allclose(Kv, Kc0, rtol=4e-15, atol=1e-300)
        assert np.allclose(Ks, Kc1, rtol=4e-15, atol=1e-300)
        assert np.array_equal(Kv, Kv.T)  # Matrix(Symmetric(K))


def test_reparameterize_identity():
    """test/test_GP.jl:35-68: eval_cov(k, ds) ≈ eval_cov(reparameterize(k, T), ds_raw)."""
    ds_raw, ds = H.fixture_grid()
    slope = 1.0 / 20.0
    intercept = 0.5
    for k in H.fixture_kernels():
        M1 = o.eval_cov(k, ds)
        M2 = o.eval_cov(H.reparameterize(k, slope, intercept), ds_raw)
        # Julia's default isapprox: rtol = sqrt(eps)
        assert np.all(np.abs(M1 - M2) <= math.sqrt(np.finfo(float).eps) * np.maximum(np.abs(M1), np.abs(M2))), k


def test_rescale_identity():
    """test/test_GP.jl:70-106: eval_cov(rescale(k, T^-1)) ≈ unapply_var(T, eval_cov(k)), atol 1e-8."""
    ds = np.linspace(-10, 10, 50)
    slope = 2.0 / 20.0  # LinearTransform(ys_raw, -1, 1)
    inv_slope = 1.0 / slope
    for k in H.fixture_kernels():
        M1 = o.eval_cov(H.rescale(k, inv_slope), ds)
        M2 = (1 / slope ** 2) * o.eval_cov(k, ds)
        assert np.all(np.abs(M1 - M2) <= 1e-8), k


def test_predictive_likelihood_identity():
    """test/experiment_hmc.jl:111-132: logpdf(predictive MVN, xs_test) ≈ LML(joint) − LML(obs)."""
    ts = np.linspace(0.0, 10.0, 1000)
    rng = np.random.default_rng(7)
    idx = rng.permutation(1000)
    obs, test = np.sort(idx[:200]), np.sort(idx[200:260])
    for k, nz in H.hmc_benchmarks():
        noise = nz + o.JITTER
        xs = H.fixture_xs(ts / 10.0) + 0.05 * rng.standard_normal(1000)
        lml_obs = o.log_marginal_likelihood(k, noise, ts[obs], xs[obs])
        both = np.concatenate([obs, test])
        lml_joint = o.log_marginal_likelihood(k, noise, ts[both], xs[both])
        mu, cov = o.predictive_mvn(k, noise, ts[obs], xs[obs], ts[test])
        lp = o.mvn_logpdf(xs[test], mu, cov)
        assert lp == pytest.approx(lml_joint - lml_obs, rel=2e-7), k


def test_model_constants_and_transforms():
    assert o.JITTER == 1e-5  # Model.jl:22
    assert o.transform_param("noise", 0.0) == math.exp(-1.5)
    assert o.transform_param("period", 1.0) == math.exp(-0.5)
    assert o.transform_param("gamma", 0.0) == 1.0
    assert 0 < o.transform_param("gamma", 5.0) <= 2


def test_posdef_exception_info():
    ts = np.linspace(0, 1, 10)
    K = o.compute_cov_matrix_vectorized(o.Constant(1.0), -2.0, ts)
    with pytest.raises(o.PosDefException) as e:
        o.mvnormal_logpdf(np.zeros(10), K)
    assert e.value.info == 1
    _, info = c_oracle.lml(o.encode_program(o.Constant(1.0)), ts, np.zeros(10), -2.0)
    assert info == 1


def test_empty_mvnormal_scores_zero():
    assert o.mvnormal_logpdf(np.zeros(0), np.zeros((0, 0))) == 0.0


def test_linear_schedule_matches_reference_examples():
    """src/Schedule.jl:24-39."""
    assert o.linear_schedule(2048, 0.10) == [205, 410, 615, 820, 1025, 1230, 1435, 1640, 1845, 2048]
    assert o.linear_schedule(10, 0.5) == [5, 10]
    assert o.linear_schedule(100, 0.3)[-1] == 100


def test_weights_and_ess():
    lw = np.array([0.0, 0.0, 0.0, 0.0])
    lt, lnw = o.normalize_weights(lw)
    assert lt == pytest.approx(math.log(4))
    assert o.effective_sample_size(lnw) == pytest.approx(4.0)
    lw = np.array([0.0, -1e9, -1e9])
    assert o.effective_sample_size(o.normalize_weights(lw)[1]) == pytest.approx(1.0)


def test_gradient_routes_agree_and_match_mpmath_value():
    """The two CPU gradient routes (neither differentiates a kernel analytically) agree, and reproduce
    a 40-digit value computed once with mpmath for the most curved parameter (the period of a
    Periodic with lengthscale 0.066: second-order differences are off by 4e-6 there)."""
    nd, nz = o.synthetic_particle(20, "se*per+lin")
    ts, xs = o.synthetic_series(128)
    g_fd, gn_fd = o.lml_grad_fd(nd, nz, ts, xs)
    g_dn, gn_dn = o.lml_grad_dense_fd(nd, nz, ts, xs)
    assert np.allclose(g_fd, g_dn, rtol=1e-7, atol=1e-7)
    assert gn_fd == pytest.approx(gn_dn, rel=1e-7)
    assert g_dn[3] == pytest.approx(2120.759417696520798, rel=2e-9)
    # with_params round-trips the wire order
    assert o.with_params(nd, o.encode_program(nd)[2]) == nd
    k = o.ChangePoint(o.Linear(0.1, 1.3, 0.7), o.Periodic(0.96, 0.21, 1.1), 0.5, 0.95)
    assert o.with_params(k, o.encode_program(k)[2]) == k


def test_transform_param_grad_is_the_derivative():
    import autogp.jl_b200 as agp

    for f in ("noise", "period", "gamma", "lengthscale"):
        for z in (-1.3, 0.0, 0.8):
            hstep = 1e-6
            fd = (agp.transform_param(f, z + hstep) - agp.transform_param(f, z - hstep)) / (2 * hstep)
            assert agp.transform_param_grad(f, z) == pytest.approx(fd, rel=1e-8)


def test_infer_gp_sum_restatement_is_consistent_with_the_predictive_mvn():
    """src/GP.jl:904-993 restated: the X(T*) block is the predictive MVN of the Plus kernel (+ JITTER), the summand
    means add up to the noiseless prediction, a single summand reproduces X* up to the noise, and summing the latent
    blocks (with their cross-covariances) gives the covariance of the noiseless X*."""
    ts, xs = o.synthetic_series(40)
    tp = np.linspace(0.9, 1.3, 7)
    nodes = [o.Linear(0.2, 0.4, 0.9), o.Periodic(0.5, 0.3, 1.2), o.GammaExponential(0.3, 1.4, 0.8)]
    noise, npred = 0.07, 0.02
    mu, cov, idx = o.infer_gp_sum(nodes, noise, ts, xs, tp, noise_pred=npred)
    m = len(tp)
    assert mu.shape == (4 * m,) and cov.shape == (4 * m, 4 * m) and list(idx["X"]) == list(range(3 * m, 4 * m))
    mu_x, cov_x = o.predictive_mvn(o.Plus(o.Plus(nodes[0], nodes[1]), nodes[2]), noise, ts, xs, tp, noise_pred=npred)
    X = list(idx["X"])
    np.testing.assert_allclose(mu[X], mu_x, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(cov[np.ix_(X, X)], cov_x + o.GP_JITTER * np.eye(m), rtol=1e-9, atol=1e-12)
    F = [list(r) for r in idx["F"]]
    np.testing.assert_allclose(sum(mu[f] for f in F), mu_x, rtol=1e-9, atol=1e-12)
    tot = sum(cov[np.ix_(fa, fb)] for fa in F for fb in F)
    np.testing.assert_allclose(tot, cov_x - npred * np.eye(m) + 3 * o.GP_JITTER * np.eye(m), rtol=1e-8, atol=1e-10)
    # Cov[F_i*, X*] = sum_j Cov[F_i*, F_j*]  (X* = sum F* + independent noise)
    for fa in F:
        np.testing.assert_allclose(cov[np.ix_(fa, X)], sum(cov[np.ix_(fa, fb)] for fb in F) - o.GP_JITTER * np.eye(m), rtol=1e-8, atol=1e-10)
    mu1, cov1, idx1 = o.infer_gp_sum(nodes[:1], noise, ts, xs, tp)
    np.testing.assert_allclose(mu1[list(idx1["F"][0])], mu1[list(idx1["X"])], rtol=1e-12)
    np.testing.assert_allclose(cov1[:m, :m] + noise * np.eye(m), cov1[m:, m:], rtol=1e-10, atol=1e-12)


def test_sum_fixture_pins_the_restatements_added_for_the_next_rows():
    """tests/golden/sum_golden.json (make_golden.py sum): infer_gp_sum, predictive marginals and the noise gradient of the
    oracle on the Plus benchmark of experiment_hmc.jl:181 — a regression pin of the checker itself."""
    import json
    import os

    import make_golden

    with open(os.path.join(os.path.dirname(make_golden.__file__), "sum_golden.json")) as f:
        gold = json.load(f)
    nodes, noise, ts, xs, tp = make_golden.sum_fixture_inputs()
    mu, cov, _ = o.infer_gp_sum(nodes, noise, ts, xs, tp, noise_pred=0.0)
    np.testing.assert_allclose(mu, gold["infer_gp_sum_mean"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(cov, gold["infer_gp_sum_cov"], rtol=1e-9, atol=1e-11)
    whole = o.Plus(nodes[0], nodes[1])
    mu_p, cov_p = o.predictive_mvn(whole, noise, ts, xs, tp)
    np.testing.assert_allclose(mu_p, gold["predictive_mean"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(np.diag(cov_p), gold["predictive_var"], rtol=1e-9, atol=1e-11)
    assert abs(o.lml_grad_dense_fd(whole, noise, ts, xs)[1] - gold["lml_grad_noise"]) <= 1e-9 * abs(gold["lml_grad_noise"])
    assert abs(o.log_marginal_likelihood(whole, noise, ts, xs) - gold["lml"]) <= 1e-11 * abs(gold["lml"])
    # the X* block of the decomposition (noise_pred = 0) is the noiseless prediction of the whole kernel
    X = slice(2 * len(tp), 3 * len(tp))
    np.testing.assert_allclose(np.asarray(gold["infer_gp_sum_mean"])[X], gold["predictive_mean"], rtol=1e-8, atol=1e-10)
