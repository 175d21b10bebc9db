"""Host-side mirror of the likelihood site of ``Model.model`` (src/Model.jl:130-138).

    noise      = transform_param(:noise, z) + JITTER                 Model.jl:134
    cov_matrix = GP.compute_cov_matrix_vectorized(node, noise, ts)   Model.jl:135
    xs ~ mvnormal(zeros(n), cov_matrix)                              Model.jl:136

``mvnormal_logpdf`` stands in for Gen's ``mvnormal`` distribution scored at ``xs`` with the
kernel tree as its argument (the custom ``Gen.Distribution`` sketched in INTEGRATION.md), and
``log_marginal_likelihoods`` is its batched form over the particles of one SMC round.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

from . import gp

JITTER = 1e-5  # Model.jl:22

# GPConfig.prior defaults (src/GP.jl:1133-1137)
PRIOR = {
    "gamma": dict(scale=2.0, mu=0.0, sigma=1.0),
    "period": dict(mu=-1.5, sigma=1.0),
    "wildcard": dict(mu=-1.5, sigma=1.0),
}


class PosDefException(Exception):
    """``LinearAlgebra.PosDefException(info)``: what PDMats' ``cholesky`` throws inside
    ``mvnormal`` when the covariance is not positive definite."""

    def __init__(self, info: int, particle: Optional[int] = None):
        where = "" if particle is None else f" (particle {particle})"
        super().__init__(f"matrix is not positive definite; Cholesky factorization failed{where}, info={info}")
        self.info = int(info)
        self.particle = particle


def transform_log_normal(z: float, mu: float, sigma: float) -> float:  # Model.jl:24
    return math.exp(mu + sigma * z)


def untransform_log_normal(param: float, mu: float, sigma: float) -> float:  # Model.jl:25
    return (math.log(param) - mu) / sigma


def transform_logit_normal(z: float, scale: float, mu: float, sigma: float) -> float:  # Model.jl:27-29
    return scale * 1 / (1 + math.exp(-(mu + sigma * z)))


def untransform_logit_normal(param: float, scale: float, mu: float, sigma: float) -> float:  # Model.jl:31-33
    return (math.log(param / (scale - param)) - mu) / sigma


def transform_param(field: str, z: float, prior=PRIOR) -> float:
    """``Model.transform_param(field, z, config)`` (Model.jl:35-48)."""
    if field == "gamma":
        p = prior["gamma"]
        return transform_logit_normal(z, p["scale"], p["mu"], p["sigma"])
    p = prior["period"] if field == "period" else prior["wildcard"]
    return transform_log_normal(z, p["mu"], p["sigma"])


def untransform_param(field: str, param: float, prior=PRIOR) -> float:
    """``Model.untransform_param`` (Model.jl:50-63)."""
    if field == "gamma":
        p = prior["gamma"]
        return untransform_logit_normal(param, p["scale"], p["mu"], p["sigma"])
    p = prior["period"] if field == "period" else prior["wildcard"]
    return untransform_log_normal(param, p["mu"], p["sigma"])


def transform_param_grad(field: str, z: float, prior=PRIOR) -> float:
    """d transform_param(field, z) / dz — the chain-rule factor that takes the gradient with respect to
    a kernel parameter (agp_lml_grad_batch) to the gradient with respect to the latent z that
    ``Gen.hmc`` / ``Gen.map_optimize`` move (src/inference_utils.jl:63-67, src/Greedy.jl:95)."""
    if field == "gamma":
        p = prior["gamma"]
        s = 1.0 / (1.0 + math.exp(-(p["mu"] + p["sigma"] * z)))
        return p["scale"] * s * (1.0 - s) * p["sigma"]
    p = prior["period"] if field == "period" else prior["wildcard"]
    return transform_log_normal(z, p["mu"], p["sigma"]) * p["sigma"]


def log_marginal_likelihood_grads(nodes: Sequence[gp.Node], noises: Sequence[float], ts, xs, *,
                                  engine: Optional[gp.Engine] = None, check: bool = True):
    """Scores and gradients of every particle in one fused GPU call: (lml[P], grads, grad_noise[P]) with
    ``grads[p]`` in the order of ``gp.encode_program(nodes[p])[2]`` — the ``logpdf_grad`` of the custom Gen
    distribution sketched in INTEGRATION.md (``has_argument_grads`` true for params and noise)."""
    lml, grads, gnoise, info = (engine or gp.default_engine()).lml_grad_batch(nodes, noises, ts, xs)
    if check:
        bad = np.nonzero(info)[0]
        if bad.size:
            raise PosDefException(int(info[bad[0]]), int(bad[0]))
    return lml, grads, gnoise


def log_marginal_likelihoods(nodes: Sequence[gp.Node], noises: Sequence[float], ts, xs, *,
                             engine: Optional[gp.Engine] = None, check: bool = True) -> np.ndarray:
    """Scores ``xs ~ mvnormal(0, K_p + noise_p I)`` for every particle p in one fused GPU call.

    With ``check=True`` a non-positive-definite covariance raises :class:`PosDefException`
    (first failing particle), as the reference's scoring would; with ``check=False`` the
    failing entries are NaN and the caller reads ``info`` via :func:`log_marginal_likelihoods_info`.
    """
    lml, info = (engine or gp.default_engine()).lml_batch(nodes, noises, ts, xs)
    if check:
        bad = np.nonzero(info)[0]
        if bad.size:
            raise PosDefException(int(info[bad[0]]), int(bad[0]))
    return lml


def log_marginal_likelihoods_info(nodes, noises, ts, xs, *, engine: Optional[gp.Engine] = None):
    return (engine or gp.default_engine()).lml_batch(nodes, noises, ts, xs)


def mvnormal_logpdf(node: gp.Node, noise: float, ts, xs, *, engine: Optional[gp.Engine] = None) -> float:
    """One particle: ``logpdf(mvnormal, xs, zeros(n), compute_cov_matrix_vectorized(node, noise, ts))``."""
    return float(log_marginal_likelihoods([node], [noise], ts, xs, engine=engine)[0])


def predictive_logpdfs(nodes: Sequence[gp.Node], noises: Sequence[float], ts, xs, ts_new, xs_new, *,
                       engine: Optional[gp.Engine] = None) -> np.ndarray:
    """``logpdf(MvNormal(node, noise, ts, xs, ts_new), xs_new)`` for every particle — the per-particle ``logp`` of
    ``predict_proba`` (src/api.jl:686-699; default ``noise_pred = noise``; in the model's transformed units, the
    reference's linear ``y_transform`` adds the constant ``-m log|slope|``).  Computed through the identity the reference's own
    test asserts (test/experiment_hmc.jl:111-132), ``log p(x_new | x) = LML(ts ∪ ts_new) − LML(ts)``, with the second
    factorisation CONTINUED from the first (``agp_lml_run_append``): no m x m predictive covariance is ever formed."""
    eng = engine or gp.default_engine()
    ts, xs = np.asarray(ts, dtype=np.float64), np.asarray(xs, dtype=np.float64)
    ts_new, xs_new = np.asarray(ts_new, dtype=np.float64), np.asarray(xs_new, dtype=np.float64)
    if ts.shape != xs.shape or ts_new.shape != xs_new.shape:
        raise ValueError("time points and values must have equal length")
    n = ts.shape[0]
    eng.upload(nodes, noises, np.concatenate([ts, ts_new]), np.concatenate([xs, xs_new]))
    if n > 0:
        eng.set_prefix(n)
        eng.run()
        before, info = eng.fetch()
        if np.any(info != 0):
            bad = int(np.nonzero(info)[0][0])
            raise PosDefException(int(info[bad]), bad)
        eng.set_prefix(n + ts_new.shape[0])
        eng.run_append()
    else:   # nothing observed yet: the empty mvnormal scores 0 and there is no factor to continue
        before = np.zeros(len(nodes))
        eng.run()
    after, info = eng.fetch()
    if np.any(info != 0):
        bad = int(np.nonzero(info)[0][0])
        raise PosDefException(int(info[bad]), bad)
    return after - before
