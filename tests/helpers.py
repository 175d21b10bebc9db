"""Shared test helpers: oracle <-> product node conversion and the reference's test fixtures."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import autogp_oracle as o  # noqa: E402
import make_golden  # noqa: E402  (fixture definitions shared with the generator)


def to_agp(nd):
    """oracle node -> autogp.jl_b200 node (same type names and field order)."""
    import autogp.jl_b200 as agp

    cls = getattr(agp, type(nd).__name__)
    if isinstance(nd, o.LEAVES):
        return cls(**nd.__dict__)
    if isinstance(nd, o.ChangePoint):
        return cls(to_agp(nd.left), to_agp(nd.right), nd.location, nd.scale)
    return cls(to_agp(nd.left), to_agp(nd.right))


def from_agp(nd):
    """autogp.jl_b200 node -> oracle node."""
    import autogp.jl_b200 as agp

    cls = getattr(o, type(nd).__name__)
    if isinstance(nd, agp.LeafNode):
        return cls(**nd.__dict__)
    if isinstance(nd, agp.ChangePoint):
        return cls(from_agp(nd.left), from_agp(nd.right), nd.location, nd.scale)
    return cls(from_agp(nd.left), from_agp(nd.right))


class OracleEngine:
    """Checker-side stand-in with the two Engine methods the loops call."""

    def __init__(self, fail=None):
        self.batches = []
        self.fail = fail            # callable(node, noise) -> bool: report "not positive definite"

    def lml_batch(self, nodes, noises, ts, xs):
        self.batches.append(("lml", len(nodes)))
        lml, info = np.full(len(nodes), np.nan), np.zeros(len(nodes), dtype=np.int32)
        for i, (nd, nz) in enumerate(zip(nodes, noises)):
            try:
                lml[i] = o.log_marginal_likelihood(from_agp(nd), nz, ts, xs)
            except o.PosDefException as e:        # what the C-ABI reports through info
                info[i] = max(int(getattr(e, "info", 1)), 1)
        return lml, info

    def lml_grad_batch(self, nodes, noises, ts, xs):
        self.batches.append(("grad", len(nodes)))
        lml, grads, gn, info = [], [], [], []
        for nd, nz in zip(nodes, noises):
            ond = from_agp(nd)
            if self.fail is not None and self.fail(nd, nz):
                lml.append(np.nan); grads.append(np.full(len(o.encode_program(ond)[2]), np.nan)); gn.append(np.nan); info.append(3)
                continue
            try:
                g, g_noise = o.lml_grad_dense_fd(ond, nz, ts, xs)
                val = o.log_marginal_likelihood(ond, nz, ts, xs)
            except (np.linalg.LinAlgError, ValueError, FloatingPointError, o.PosDefException):
                g, g_noise, val = np.full(len(o.encode_program(ond)[2]), np.nan), np.nan, np.nan
            lml.append(val); grads.append(g); gn.append(g_noise); info.append(0 if np.isfinite(val) else 1)
        return np.array(lml), grads, np.array(gn), np.array(info, dtype=np.int32)


class OracleEngineWithNoiseCall(OracleEngine):
    """... plus the noise-gradient-only call of the real engine."""

    def lml_grad_noise_batch(self, nodes, noises, ts, xs):
        lml, _, gn, info = OracleEngine.lml_grad_batch(self, nodes, noises, ts, xs)
        self.batches[-1] = ("noise", len(nodes))
        return lml, gn, info


def reparameterize(nd, slope, intercept):
    """GP.reparameterize(node, LinearTransform(slope, intercept)) — src/GP.jl:142,168,205-209,
    247-250,291-294,338-341,382-386,425-429,505-511 (used only to replay test_GP.jl identities)."""
    if isinstance(nd, (o.WhiteNoise, o.Constant)):
        return nd
    if isinstance(nd, o.Linear):
        return o.Linear((nd.intercept - intercept) / slope, nd.bias, slope ** 2 * nd.amplitude)
    if isinstance(nd, o.SquaredExponential):
        return o.SquaredExponential(nd.lengthscale / abs(slope), nd.amplitude)
    if isinstance(nd, o.GammaExponential):
        return o.GammaExponential(nd.lengthscale / abs(slope), nd.gamma, nd.amplitude)
    if isinstance(nd, o.Periodic):
        return o.Periodic(nd.lengthscale, nd.period / abs(slope), nd.amplitude)
    if isinstance(nd, o.Plus):
        return o.Plus(reparameterize(nd.left, slope, intercept), reparameterize(nd.right, slope, intercept))
    if isinstance(nd, o.Times):
        return o.Times(reparameterize(nd.left, slope, intercept), reparameterize(nd.right, slope, intercept))
    return o.ChangePoint(reparameterize(nd.left, slope, intercept), reparameterize(nd.right, slope, intercept),
                         (nd.location - intercept) / slope, nd.scale / slope)


def rescale(nd, slope):
    """GP.rescale(node, LinearTransform(slope, .)) — src/GP.jl:143,169,211-215,252-255,296-299,
    343-346,388-392,431-436,513-517."""
    s2 = slope ** 2
    if isinstance(nd, o.WhiteNoise):
        return o.WhiteNoise(s2 * nd.value)
    if isinstance(nd, o.Constant):
        return o.Constant(s2 * nd.value)
    if isinstance(nd, o.Linear):
        return o.Linear(nd.intercept, s2 * nd.bias, s2 * nd.amplitude)
    if isinstance(nd, o.SquaredExponential):
        return o.SquaredExponential(nd.lengthscale, s2 * nd.amplitude)
    if isinstance(nd, o.GammaExponential):
        return o.GammaExponential(nd.lengthscale, nd.gamma, s2 * nd.amplitude)
    if isinstance(nd, o.Periodic):
        return o.Periodic(nd.lengthscale, nd.period, s2 * nd.amplitude)
    if isinstance(nd, o.Plus):
        return o.Plus(rescale(nd.left, slope), rescale(nd.right, slope))
    if isinstance(nd, o.Times):
        return o.Times(rescale(nd.left, slope), nd.right)
    return o.ChangePoint(rescale(nd.left, slope), rescale(nd.right, slope), nd.location, nd.scale)


def golden():
    import json

    g = np.load(os.path.join(ROOT, "tests", "golden", "gram_golden.npz"))
    with open(os.path.join(ROOT, "tests", "golden", "lml_golden.json")) as f:
        lml = json.load(f)
    return g["ts"], g["grams"], lml


fixture_kernels = make_golden.fixture_kernels
base_kernels = make_golden.base_kernels
fixture_grid = make_golden.fixture_grid
fixture_xs = make_golden.fixture_xs
hmc_benchmarks = make_golden.hmc_benchmarks


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
