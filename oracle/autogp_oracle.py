"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see below).

A NumPy/SciPy FP64 restatement of the AutoGP.jl GP log-marginal-likelihood hot path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this
module, and only as the *checker* / reported CPU baseline.  The product path
(``autogp.jl_b200``) never imports it and has no CPU fallback.

PARITY UNPINNED: Julia is not installed in this image and the reference ships no golden
vectors for the Gram matrix or the LML, and the LML arithmetic lives in
un-vendored third-party packages:
  * Gen.jl        (compat 0.4.5, Project.toml:27)  ``mvnormal`` logpdf
  * Distributions (0.25.79,     Project.toml:24)  ``logpdf(::MvNormal)``
  * PDMats.jl     (transitive, unpinned)          ``cholesky`` / ``invquad`` / ``logdet``
  * LinearAlgebra / OpenBLAS (Julia stdlib)       ``dpotrf('U')``, ``dtrsv``
Their published algorithm is restated in :func:`mvnormal_logpdf`.  The oracle is therefore
anchored by (i) following the reference's Gram code operation-by-operation (citations on
every function), (ii) three independent routes that must agree (vectorised NumPy, the scalar
C restatement in ``oracle/oracle_c.c``, and an mpmath 50-digit evaluation), and (iii) the
identities the reference's own tests assert (test/test_GP.jl:35-106,
test/experiment_hmc.jl:111-132), frozen as fixtures under ``tests/golden/``.

All file:line citations are into /root/reference (probsys/AutoGP.jl @ 2ad372d).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence, Tuple, Union

import numpy as np

# ----------------------------------------------------------------------------------------
# Kernel tree (src/GP.jl:39-51, 131-133, 157-159, 185-192, 228-234, 269-277, 315-322,
# 358-369, 404-415, 466-479).  Field order == Julia ``fieldnames`` order.
# ----------------------------------------------------------------------------------------


@dataclass(frozen=True)
class WhiteNoise:  # GP.jl:131-133
    value: float


@dataclass(frozen=True)
class Constant:  # GP.jl:157-159
    value: float


@dataclass(frozen=True)
class Linear:  # GP.jl:185-192 (bias=1, amplitude=1 defaults)
    intercept: float
    bias: float = 1.0
    amplitude: float = 1.0


@dataclass(frozen=True)
class SquaredExponential:  # GP.jl:228-234
    lengthscale: float
    amplitude: float = 1.0


@dataclass(frozen=True)
class GammaExponential:  # GP.jl:269-277 (asserts 0 < gamma <= 2)
    lengthscale: float
    gamma: float
    amplitude: float = 1.0

    def __post_init__(self):
        assert 0 < self.gamma <= 2


@dataclass(frozen=True)
class Periodic:  # GP.jl:315-322
    lengthscale: float
    period: float
    amplitude: float = 1.0


@dataclass(frozen=True)
class Plus:  # GP.jl:358-369
    left: "Node"
    right: "Node"


@dataclass(frozen=True)
class Times:  # GP.jl:404-415
    left: "Node"
    right: "Node"


@dataclass(frozen=True)
class ChangePoint:  # GP.jl:466-479
    left: "Node"
    right: "Node"
    location: float
    scale: float


Node = Union[WhiteNoise, Constant, Linear, SquaredExponential, GammaExponential, Periodic,
             Plus, Times, ChangePoint]
LEAVES = (WhiteNoise, Constant, Linear, SquaredExponential, GammaExponential, Periodic)
BINARY = (Plus, Times, ChangePoint)


def size(node: Node) -> int:  # GP.jl:93-95, 366, 412, 476
    return 1 if isinstance(node, LEAVES) else 1 + size(node.left) + size(node.right)


def depth(node: Node) -> int:  # GP.jl:103-104, 367
    return 1 if isinstance(node, LEAVES) else 1 + max(depth(node.left), depth(node.right))


def unroll(node: Node) -> List[Node]:  # GP.jl:111-113 (postfix order)
    if isinstance(node, LEAVES):
        return [node]
    return unroll(node.left) + unroll(node.right) + [node]


# ----------------------------------------------------------------------------------------
# Vectorised Gram build: eval_cov(node, ts)  — one n×n temporary per op, like the reference.
# Operation order mirrors the Julia broadcast expressions exactly (no FMA contraction: NumPy
# materialises each ufunc result, as Julia does between non-fused statements).
# ----------------------------------------------------------------------------------------


def sigma_cp(x, location, scale):  # GP.jl:481-483
    return 0.5 * (1.0 + np.tanh((location - x) / scale))


def eval_cov(node: Node, ts: np.ndarray) -> np.ndarray:
    ts = np.asarray(ts, dtype=np.float64)
    n = ts.shape[0]
    if isinstance(node, WhiteNoise):  # GP.jl:137-140: (ts .== ts') * value
        return (ts[:, None] == ts[None, :]).astype(np.float64) * float(node.value)
    if isinstance(node, Constant):  # GP.jl:163-166
        return float(node.value) * np.ones((n, n))
    if isinstance(node, Linear):  # GP.jl:199-203
        tm = ts - node.intercept
        C = tm[:, None] * tm[None, :]
        return node.bias + node.amplitude * C
    if isinstance(node, SquaredExponential):  # GP.jl:241-245
        dx = ts[:, None] - ts[None, :]
        l2 = node.lengthscale * node.lengthscale  # `^2` literal power == x*x in Julia
        C = np.exp(((-0.5 * dx) * dx) / l2)
        return node.amplitude * C
    if isinstance(node, GammaExponential):  # GP.jl:285-289
        dt = np.abs(ts[:, None] - ts[None, :])
        C = np.exp(-np.power(dt / node.lengthscale, node.gamma))
        return node.amplitude * C
    if isinstance(node, Periodic):  # GP.jl:331-336
        freq = math.pi / node.period
        dx = np.abs(ts[:, None] - ts[None, :])
        s = np.sin(freq * dx)
        coef = -2.0 / (node.lengthscale * node.lengthscale)
        C = np.exp(coef * (s * s))
        return node.amplitude * C
    if isinstance(node, Plus):  # GP.jl:375-377
        return eval_cov(node.left, ts) + eval_cov(node.right, ts)
    if isinstance(node, Times):  # GP.jl:421-423
        return eval_cov(node.left, ts) * eval_cov(node.right, ts)
    if isinstance(node, ChangePoint):  # GP.jl:493-503
        cx = sigma_cp(ts, node.location, node.scale)
        sig_1 = cx[:, None] * cx[None, :]
        sig_2 = (1.0 - cx)[:, None] * (1.0 - cx)[None, :]
        k_1 = eval_cov(node.left, ts)
        k_2 = eval_cov(node.right, ts)
        K = sig_1 * k_1 + sig_2 * k_2
        # Matrix(Symmetric(K)): upper triangle mirrored onto the lower (GP.jl:501-502)
        return np.triu(K) + np.triu(K, 1).T
    raise TypeError(f"unknown node {node!r}")


def compute_cov_matrix_vectorized(node: Node, noise: float, ts) -> np.ndarray:
    """GP.jl:666-668: eval_cov(node, ts) + noise*I (diagonal only)."""
    K = eval_cov(node, np.asarray(ts, dtype=np.float64)).copy()
    idx = np.arange(K.shape[0])
    K[idx, idx] = K[idx, idx] + noise
    return K


# ----------------------------------------------------------------------------------------
# Scalar Gram build: eval_cov(node, t1, t2) / compute_cov_matrix (GP.jl:135-491, 674-684).
# Pure-Python loops: small cases only.
# ----------------------------------------------------------------------------------------


def eval_cov_scalar(node: Node, t1: float, t2: float) -> float:
    if isinstance(node, WhiteNoise):  # GP.jl:135
        return float(t1 == t2) * node.value
    if isinstance(node, Constant):  # GP.jl:161
        return node.value
    if isinstance(node, Linear):  # GP.jl:194-197
        c = (t1 - node.intercept) * (t2 - node.intercept)
        return node.bias + node.amplitude * c
    if isinstance(node, SquaredExponential):  # GP.jl:236-239
        c = math.exp(-0.5 * (t1 - t2) * (t1 - t2) / (node.lengthscale * node.lengthscale))
        return node.amplitude * c
    if isinstance(node, GammaExponential):  # GP.jl:279-283
        dt = abs(t1 - t2)
        c = math.exp(-((dt / node.lengthscale) ** node.gamma))
        return node.amplitude * c
    if isinstance(node, Periodic):  # GP.jl:324-329
        freq = math.pi / node.period
        dx = abs(t1 - t2)
        s = math.sin(freq * dx)
        c = math.exp((-2.0 / (node.lengthscale * node.lengthscale)) * (s * s))
        return node.amplitude * c
    if isinstance(node, Plus):  # GP.jl:371-373
        return eval_cov_scalar(node.left, t1, t2) + eval_cov_scalar(node.right, t1, t2)
    if isinstance(node, Times):  # GP.jl:417-419
        return eval_cov_scalar(node.left, t1, t2) * eval_cov_scalar(node.right, t1, t2)
    if isinstance(node, ChangePoint):  # GP.jl:485-491
        s1 = 0.5 * (1.0 + math.tanh((node.location - t1) / node.scale))
        s2 = 0.5 * (1.0 + math.tanh((node.location - t2) / node.scale))
        k_left = s1 * eval_cov_scalar(node.left, t1, t2) * s2
        k_right = (1 - s1) * eval_cov_scalar(node.right, t1, t2) * (1 - s2)
        return k_left + k_right
    raise TypeError(f"unknown node {node!r}")


def compute_cov_matrix(node: Node, noise: float, ts) -> np.ndarray:
    """GP.jl:674-684 (scalar double loop, noise added on [i,i])."""
    ts = np.asarray(ts, dtype=np.float64)
    n = ts.shape[0]
    K = np.empty((n, n))
    for i in range(n):
        for j in range(n):
            K[i, j] = eval_cov_scalar(node, float(ts[i]), float(ts[j]))
        K[i, i] += noise
    return K


# ----------------------------------------------------------------------------------------
# Parameter transforms and the likelihood site (src/Model.jl:22-48, 130-138).
# ----------------------------------------------------------------------------------------

JITTER = 1e-5  # Model.jl:22
GP_JITTER = 1e-8  # the GP module's own constant (src/GP.jl:760), used by infer_gp_sum
PRIOR = {  # GP.jl:1133-1137
    "gamma": dict(scale=2.0, mu=0.0, sigma=1.0),
    "period": dict(mu=-1.5, sigma=1.0),
    "wildcard": dict(mu=-1.5, sigma=1.0),
}


def transform_log_normal(z, mu, sigma):  # Model.jl:24
    return math.exp(mu + sigma * z)


def transform_logit_normal(z, scale, mu, sigma):  # Model.jl:27-29
    return scale * 1 / (1 + math.exp(-(mu + sigma * z)))


def transform_param(field: str, z: float) -> float:  # Model.jl:35-48
    if field == "gamma":
        p = PRIOR["gamma"]
        return transform_logit_normal(z, p["scale"], p["mu"], p["sigma"])
    p = PRIOR["period"] if field == "period" else PRIOR["wildcard"]
    return transform_log_normal(z, p["mu"], p["sigma"])


class PosDefException(Exception):
    """Mirrors LinearAlgebra.PosDefException(info) raised by PDMats' cholesky."""

    def __init__(self, info: int):
        super().__init__(f"matrix is not positive definite; Cholesky factorization failed (info={info})")
        self.info = info


def cholesky_upper(K: np.ndarray) -> np.ndarray:
    """LAPACK dpotrf('U') on Symmetric(K) — only the upper triangle is read.  Unblocked
    restatement (row-by-row Cholesky–Crout); raises PosDefException(k) with LAPACK's info."""
    import scipy.linalg.lapack as lapack

    U, info = lapack.dpotrf(np.asfortranarray(K), lower=0, clean=1, overwrite_a=0)
    if info > 0:
        raise PosDefException(int(info))
    if info < 0:
        raise ValueError(f"dpotrf illegal argument {-info}")
    return U


def mvnormal_logpdf(xs: np.ndarray, K: np.ndarray) -> float:
    """Gen ``mvnormal`` logpdf with mu = 0 (Model.jl:136) →
    Distributions ``logpdf(MvNormal(0, PDMat(Symmetric(K))), x)``:
        -(n*log(2π) + logdet(K))/2 - invquad(K, x)/2,
    logdet = 2 Σ log U_ii,  invquad = ‖U⁻ᵀ x‖²  (PDMats, upper Cholesky K = UᵀU)."""
    import scipy.linalg

    xs = np.asarray(xs, dtype=np.float64)
    n = xs.shape[0]
    if n == 0:
        return 0.0
    U = cholesky_upper(K)
    logdet = 2.0 * float(np.sum(np.log(np.diag(U))))
    z = scipy.linalg.solve_triangular(U, xs, trans="T", lower=False)
    return -0.5 * (n * math.log(2.0 * math.pi) + logdet) - 0.5 * float(z @ z)


def log_marginal_likelihood(node: Node, noise: float, ts, xs) -> float:
    """Model.jl:134-136 with `noise` already transformed (+JITTER applied by the caller)."""
    K = compute_cov_matrix_vectorized(node, noise, ts)
    return mvnormal_logpdf(np.asarray(xs, dtype=np.float64), K)


def log_marginal_likelihood_lu(node: Node, noise: float, ts, xs) -> float:
    """Third route (no Cholesky): slogdet + LU solve.  Guards against a shared bug."""
    K = compute_cov_matrix_vectorized(node, noise, ts)
    xs = np.asarray(xs, dtype=np.float64)
    sign, logdet = np.linalg.slogdet(K)
    assert sign > 0
    alpha = np.linalg.solve(K, xs)
    return -0.5 * (len(xs) * math.log(2 * math.pi) + logdet) - 0.5 * float(xs @ alpha)


def log_marginal_likelihood_mp(node: Node, noise: float, ts, xs, dps: int = 50) -> float:
    """mpmath `dps`-digit LML on the *double-precision* inputs (n ≲ 64): the ground truth the
    FP64 routes are scored against."""
    import mpmath as mp

    with mp.workdps(dps):
        tsm = [mp.mpf(float(t)) for t in ts]
        n = len(tsm)

        def k(nd, a, b):
            if isinstance(nd, WhiteNoise):
                return mp.mpf(nd.value) if a == b else mp.mpf(0)
            if isinstance(nd, Constant):
                return mp.mpf(nd.value)
            if isinstance(nd, Linear):
                return mp.mpf(nd.bias) + mp.mpf(nd.amplitude) * (a - mp.mpf(nd.intercept)) * (b - mp.mpf(nd.intercept))
            if isinstance(nd, SquaredExponential):
                return mp.mpf(nd.amplitude) * mp.exp(-(a - b) ** 2 / (2 * mp.mpf(nd.lengthscale) ** 2))
            if isinstance(nd, GammaExponential):
                d = abs(a - b)
                if d == 0:
                    return mp.mpf(nd.amplitude)
                return mp.mpf(nd.amplitude) * mp.exp(-((d / mp.mpf(nd.lengthscale)) ** mp.mpf(nd.gamma)))
            if isinstance(nd, Periodic):
                s = mp.sin(mp.pi / mp.mpf(nd.period) * abs(a - b))
                return mp.mpf(nd.amplitude) * mp.exp(-2 / mp.mpf(nd.lengthscale) ** 2 * s * s)
            if isinstance(nd, Plus):
                return k(nd.left, a, b) + k(nd.right, a, b)
            if isinstance(nd, Times):
                return k(nd.left, a, b) * k(nd.right, a, b)
            if isinstance(nd, ChangePoint):
                s1 = (1 + mp.tanh((mp.mpf(nd.location) - a) / mp.mpf(nd.scale))) / 2
                s2 = (1 + mp.tanh((mp.mpf(nd.location) - b) / mp.mpf(nd.scale))) / 2
                return s1 * k(nd.left, a, b) * s2 + (1 - s1) * k(nd.right, a, b) * (1 - s2)
            raise TypeError(nd)

        K = mp.matrix(n, n)
        for i in range(n):
            for j in range(n):
                K[i, j] = k(node, tsm[i], tsm[j])
            K[i, i] += mp.mpf(float(noise))
        L = mp.cholesky(K)
        y = mp.matrix([mp.mpf(float(v)) for v in xs])
        z = mp.lu_solve(L, y)  # L is triangular; exact enough at 50 digits
        logdet = 2 * sum(mp.log(L[i, i]) for i in range(n))
        q = sum(z[i] * z[i] for i in range(n))
        return float(-(n * mp.log(2 * mp.pi) + logdet) / 2 - q / 2)


# ----------------------------------------------------------------------------------------
# Predictive conditional (src/GP.jl:731-758) — used for the experiment_hmc.jl:111-132 identity.
# ----------------------------------------------------------------------------------------


def predictive_mvn(node: Node, noise: float, ts, xs, ts_pred, noise_pred=None) -> Tuple[np.ndarray, np.ndarray]:
    ts = np.asarray(ts, dtype=np.float64)
    xs = np.asarray(xs, dtype=np.float64)
    ts_pred = np.asarray(ts_pred, dtype=np.float64)
    noise_pred = noise if noise_pred is None else noise_pred
    n_prev, n_new = len(ts), len(ts_pred)
    cov = compute_cov_matrix_vectorized(node, 0.0, np.concatenate([ts, ts_pred]))
    c11 = cov[:n_prev, :n_prev] + noise * np.eye(n_prev)
    c22 = cov[n_prev:, n_prev:]
    c12 = cov[:n_prev, n_prev:]
    c21 = cov[n_prev:, :n_prev]
    mu = c21 @ np.linalg.solve(c11, xs)
    cc = c22 - c21 @ np.linalg.solve(c11, c12)
    cc = 0.5 * cc + 0.5 * cc.T
    cc = cc + noise_pred * np.eye(n_new)
    return mu, cc


def infer_gp_sum(nodes: List[Node], noise: float, ts, xs, ts_pred, noise_pred=None):
    """src/GP.jl:904-993, block by block: joint prior over Z = [F_1(T*); ...; F_m(T*); X(T*); X(T)], conditioned on
    X(T) = xs.  Returns (mean, cov incl. GP.JITTER, index ranges F / X)."""
    ts = np.asarray(ts, dtype=np.float64)
    xs = np.asarray(xs, dtype=np.float64)
    ts_pred = np.asarray(ts_pred, dtype=np.float64)
    m, n, p = len(nodes), len(ts), len(ts_pred)
    noise_pred = noise if noise_pred is None else noise_pred            # :915
    z = np.concatenate([ts, ts_pred])                                   # :921
    Ktt, Ktp, Kpp = [], [], []
    for nd in nodes:                                                    # :922-930
        Ki = compute_cov_matrix_vectorized(nd, 0.0, z)
        a, b, c = Ki[:n, :n], Ki[:n, n:], Ki[n:, n:]
        Ktt.append(0.5 * (a + a.T))
        Ktp.append(b)
        Kpp.append(0.5 * (c + c.T))
    S_tt, S_tp, S_pp = Ktt[0].copy(), Ktp[0].copy(), Kpp[0].copy()      # reduce(+, ...), :933-935
    for i in range(1, m):
        S_tt, S_tp, S_pp = S_tt + Ktt[i], S_tp + Ktp[i], S_pp + Kpp[i]
    d_lat, d_all = m * p, m * p + p + n
    Sigma = np.zeros((d_all, d_all))
    xP = np.arange(d_lat, d_lat + p)
    xT = np.arange(d_lat + p, d_all)
    for i in range(m):                                                  # :947-961
        lP = np.arange(i * p, (i + 1) * p)
        Sigma[np.ix_(lP, lP)] = Kpp[i]
        Sigma[np.ix_(lP, xP)] = Kpp[i]
        Sigma[np.ix_(xP, lP)] = Kpp[i].T
        Sigma[np.ix_(lP, xT)] = Ktp[i].T
        Sigma[np.ix_(xT, lP)] = Ktp[i]
    Sigma[np.ix_(xT, xT)] = S_tt + noise * np.eye(n)                    # :964-967
    Sigma[np.ix_(xT, xP)] = S_tp
    Sigma[np.ix_(xP, xT)] = S_tp.T
    Sigma[np.ix_(xP, xP)] = S_pp + noise_pred * np.eye(p)
    Sigma = 0.5 * (Sigma + Sigma.T)                                     # :970
    keep = np.concatenate([np.arange(d_lat), xP])                       # :973-974
    S_aa, S_ab, S_bb = Sigma[np.ix_(keep, keep)], Sigma[np.ix_(keep, xT)], Sigma[np.ix_(xT, xT)]
    import scipy.linalg

    U = cholesky_upper(S_bb)                                            # :982
    solve = lambda B: scipy.linalg.cho_solve((U, False), B)
    mu = S_ab @ solve(xs)                                               # :983
    cov = S_aa - S_ab @ solve(S_ab.T)                                   # :984
    cov = 0.5 * (cov + cov.T) + GP_JITTER * np.eye(len(keep))           # :985-986 (GP.JITTER, :760)
    return mu, cov, {"F": [range(i * p, (i + 1) * p) for i in range(m)], "X": range(d_lat, d_lat + p)}


# ----------------------------------------------------------------------------------------
# Gradient of the LML (what Gen.hmc / map_optimize obtain from mvnormal.logpdf_grad + ReverseDiff
# through eval_cov: src/inference_utils.jl:63-67, src/Greedy.jl:95, 370).  Two independent routes,
# neither of which differentiates a kernel analytically:
#   lml_grad_fd        central differences of log_marginal_likelihood itself
#   lml_grad_dense_fd  1/2 tr((alpha alpha' - K^-1) dK/dtheta) with dK/dtheta by central differences
# ----------------------------------------------------------------------------------------


def with_params(node: Node, params: np.ndarray) -> Node:
    """Rebuild `node` with the parameter vector of encode_program(node)[2] replaced by `params`
    (wire order: nodes in unroll order, Julia fieldnames order per node)."""
    it = iter(np.asarray(params, dtype=np.float64).tolist())

    def build(nd):
        if isinstance(nd, LEAVES):
            vals = {f: next(it) for f in nd.__dataclass_fields__}
            return type(nd)(**vals)
        left, right = build(nd.left), build(nd.right)
        if isinstance(nd, ChangePoint):
            return ChangePoint(left, right, next(it), next(it))
        return type(nd)(left, right)

    return build(node)


def _stencil5(f, x0: float, hstep: float):
    """Fourth-order central difference (-f(2h) + 8 f(h) - 8 f(-h) + f(-2h)) / 12h."""
    return (-f(x0 + 2 * hstep) + 8.0 * f(x0 + hstep) - 8.0 * f(x0 - hstep) + f(x0 - 2 * hstep)) / (12.0 * hstep)


def lml_grad_fd(node: Node, noise: float, ts, xs, rel_step: float = 1e-5) -> Tuple[np.ndarray, float]:
    params = encode_program(node)[2]
    g = np.zeros_like(params)
    for j in range(len(params)):
        def f(v, j=j):
            q = params.copy()
            q[j] = v
            return log_marginal_likelihood(with_params(node, q), noise, ts, xs)
        g[j] = _stencil5(f, params[j], rel_step * max(abs(params[j]), 1e-3))
    gn = _stencil5(lambda v: log_marginal_likelihood(node, v, ts, xs), noise, rel_step * max(abs(noise), 1e-3))
    return g, gn


def lml_grad_dense_fd(node: Node, noise: float, ts, xs, rel_step: float = 1e-5) -> Tuple[np.ndarray, float]:
    ts = np.asarray(ts, dtype=np.float64)
    xs = np.asarray(xs, dtype=np.float64)
    K = compute_cov_matrix_vectorized(node, noise, ts)
    Kinv = np.linalg.inv(K)
    alpha = Kinv @ xs
    A = np.outer(alpha, alpha) - Kinv
    params = encode_program(node)[2]
    g = np.zeros_like(params)
    for j in range(len(params)):
        def f(v, j=j):
            q = params.copy()
            q[j] = v
            return eval_cov(with_params(node, q), ts)
        dK = _stencil5(f, params[j], rel_step * max(abs(params[j]), 1e-3))
        g[j] = 0.5 * float(np.sum(A * dK))
    return g, 0.5 * float(np.trace(A))


def mvn_logpdf(x: np.ndarray, mu: np.ndarray, cov: np.ndarray) -> float:
    return mvnormal_logpdf(np.asarray(x) - np.asarray(mu), cov)


# ----------------------------------------------------------------------------------------
# SMC weight consumers (src/inference_smc_anneal_data.jl:22-31, 127-141; Gen particle filter).
# ----------------------------------------------------------------------------------------


def logsumexp(v: np.ndarray) -> float:
    v = np.asarray(v, dtype=np.float64)
    m = float(np.max(v))
    if not math.isfinite(m):
        return m
    return m + math.log(float(np.sum(np.exp(v - m))))


def normalize_weights(log_weights: np.ndarray) -> Tuple[float, np.ndarray]:
    """Gen.normalize_weights: (log_total_weight, log_normalized_weights)."""
    lt = logsumexp(log_weights)
    return lt, np.asarray(log_weights, dtype=np.float64) - lt


def effective_sample_size(log_normalized_weights: np.ndarray) -> float:
    """Gen.effective_sample_size: exp(-logsumexp(2*lnw)) = 1/Σw²."""
    return math.exp(-logsumexp(2.0 * np.asarray(log_normalized_weights)))


def linear_schedule(n: int, percent: float) -> List[int]:  # src/Schedule.jl:24-39
    assert 0 < n and 0 < percent < 1
    step = int(round(percent * n))  # Julia round = half-to-even, like Python's
    cps = list(range(step, n + 1, step))
    remaining = n - cps[-1]
    assert 0 <= remaining < step
    if remaining == 0:
        return cps
    if remaining < step / 2:
        cps[-1] = n
        return cps
    return cps + [n]


# ----------------------------------------------------------------------------------------
# Wire format shared with include/agp_b200.h (opcodes = GPConfig codes, GP.jl:1101-1108, +9).
# The encoder here is an independent restatement used to cross-check the product's encoder.
# ----------------------------------------------------------------------------------------

OP_CONSTANT, OP_LINEAR, OP_SE, OP_GE, OP_PERIODIC, OP_PLUS, OP_TIMES, OP_CP, OP_WN = 1, 2, 3, 4, 5, 6, 7, 8, 9


def encode_program(node: Node) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    ops, offs, params = [], [], []
    for nd in unroll(node):
        offs.append(len(params))
        if isinstance(nd, WhiteNoise):
            ops.append(OP_WN); params += [nd.value]
        elif isinstance(nd, Constant):
            ops.append(OP_CONSTANT); params += [nd.value]
        elif isinstance(nd, Linear):
            ops.append(OP_LINEAR); params += [nd.intercept, nd.bias, nd.amplitude]
        elif isinstance(nd, SquaredExponential):
            ops.append(OP_SE); params += [nd.lengthscale, nd.amplitude]
        elif isinstance(nd, GammaExponential):
            ops.append(OP_GE); params += [nd.lengthscale, nd.gamma, nd.amplitude]
        elif isinstance(nd, Periodic):
            ops.append(OP_PERIODIC); params += [nd.lengthscale, nd.period, nd.amplitude]
        elif isinstance(nd, Plus):
            ops.append(OP_PLUS)
        elif isinstance(nd, Times):
            ops.append(OP_TIMES)
        elif isinstance(nd, ChangePoint):
            ops.append(OP_CP); params += [nd.location, nd.scale]
    return (np.asarray(ops, dtype=np.int32), np.asarray(offs, dtype=np.int32),
            np.asarray(params, dtype=np.float64))


# ----------------------------------------------------------------------------------------
# Synthetic workload of SURVEY.md §8(d) (shared by tests and bench so both see one definition).
# ----------------------------------------------------------------------------------------


def synthetic_series(n: int) -> Tuple[np.ndarray, np.ndarray]:
    t = np.arange(n, dtype=np.float64) / max(n - 1, 1)
    perm = np.random.default_rng(0).permutation(n)
    eps = np.random.default_rng(1).standard_normal(n)
    x = 0.3 * np.sin(2 * np.pi * 4 * t) + 0.5 * (t - 0.5) + 0.05 * eps
    # LinearTransform(data, width=1): mean 0, range 1 (Transforms.jl:71-81)
    a = float(x.max() - x.min())
    if a == 0.0:  # n == 1: the reference's LinearTransform refuses (<2 values); keep the raw value
        return t[perm].copy(), x[perm].copy()
    x = (1.0 / a) * x + (-(1.0 * float(x.mean())) / a)
    return t[perm].copy(), x[perm].copy()


def synthetic_particle(p: int, tree: str = "se*per+lin") -> Tuple[Node, float]:
    """Per-particle hyper-parameters drawn from the reference priors (GP.jl:1133-1137)."""
    rng = np.random.default_rng(1000 + p)

    def pos():
        return transform_param("wildcard", float(rng.standard_normal()))

    def per():
        return transform_param("period", float(rng.standard_normal()))

    def gam():
        return transform_param("gamma", float(rng.standard_normal()))

    if tree == "se*per+lin":
        node = Plus(Times(SquaredExponential(pos(), pos()), Periodic(pos(), per(), pos())),
                    Linear(pos(), pos(), pos()))
    elif tree == "se+wn":
        node = Plus(SquaredExponential(pos(), pos()), WhiteNoise(pos()))
    elif tree == "ge+per*lin":
        node = Plus(GammaExponential(pos(), gam(), pos()), Times(Periodic(pos(), per(), pos()), Linear(pos(), pos(), pos())))
    elif tree == "cp(lin,se)":
        node = ChangePoint(Linear(pos(), pos(), pos()), SquaredExponential(pos(), pos()), pos(), 0.001)
    else:
        raise ValueError(tree)
    noise = transform_param("noise", float(rng.standard_normal())) + JITTER
    return node, noise
