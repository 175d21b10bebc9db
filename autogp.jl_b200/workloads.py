"""Synthetic workloads of SURVEY.md §8(d): the series and per-particle hyper-parameters every benchmark, probe and
stress tool of this repository runs on.  Product-side definition (the oracle keeps its own, independently written
copy for the parity tests; ``tests/test_abi_host.py`` checks that the two agree bit for bit).

Data conventions follow the reference: time points in [0, 1] and shuffled (``src/api.jl:98-102``), observations
mapped to mean 0 / range 1 (``LinearTransform``, ``src/Transforms.jl:71-81``), hyper-parameters drawn from the
reference priors through ``transform_param`` (``src/Model.jl:35-48``, ``src/GP.jl:1133-1137``).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import gp
from .model import JITTER, transform_param

TREES = ("se*per+lin", "se+wn", "ge+per*lin", "cp(lin,se)")


def synthetic_series(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """``t_i = i/(n-1)`` shuffled by ``default_rng(0)``; ``x = 0.3 sin(8 pi t) + 0.5 (t - 0.5) + 0.05 eps``, rescaled."""
    t = np.arange(n, dtype=np.float64) / max(n - 1, 1)
    perm = np.random.default_rng(0).permutation(n)
    eps = np.random.default_rng(1).standard_normal(n)
    x = 0.3 * np.sin(2 * np.pi * 4 * t) + 0.5 * (t - 0.5) + 0.05 * eps
    width = float(x.max() - x.min())
    if width == 0.0:  # n == 1: LinearTransform needs two distinct values; keep the raw value
        return t[perm].copy(), x[perm].copy()
    x = (1.0 / width) * x + (-(1.0 * float(x.mean())) / width)
    return t[perm].copy(), x[perm].copy()


def synthetic_particle(p: int, tree: str = "se*per+lin") -> Tuple[gp.Node, float]:
    """Particle ``p``: one N(0,1) latent per parameter from ``default_rng(1000 + p)``, pushed through the prior transforms."""
    rng = np.random.default_rng(1000 + p)

    def draw(field: str) -> float:
        return transform_param(field, float(rng.standard_normal()))

    def pos() -> float:
        return draw("wildcard")

    if tree == "se*per+lin":
        node = gp.Plus(gp.Times(gp.SquaredExponential(pos(), pos()), gp.Periodic(pos(), draw("period"), pos())),
                       gp.Linear(pos(), pos(), pos()))
    elif tree == "se+wn":
        node = gp.Plus(gp.SquaredExponential(pos(), pos()), gp.WhiteNoise(pos()))
    elif tree == "ge+per*lin":
        node = gp.Plus(gp.GammaExponential(pos(), draw("gamma"), pos()),
                       gp.Times(gp.Periodic(pos(), draw("period"), pos()), gp.Linear(pos(), pos(), pos())))
    elif tree == "cp(lin,se)":
        node = gp.ChangePoint(gp.Linear(pos(), pos(), pos()), gp.SquaredExponential(pos(), pos()), pos(), 0.001)
    else:
        raise ValueError(f"unknown tree {tree!r}; one of {TREES}")
    noise = draw("noise") + JITTER
    return node, noise


def synthetic_batch(P: int, tree: str = "se*per+lin", first: int = 0) -> Tuple[List[gp.Node], List[float]]:
    """Particles ``first .. first + P - 1`` (a rank of a sharded run passes its own offset)."""
    parts = [synthetic_particle(first + p, tree) for p in range(P)]
    return [nd for nd, _ in parts], [nz for _, nz in parts]
