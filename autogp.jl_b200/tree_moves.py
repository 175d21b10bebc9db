"""The reference's structure proposals behind the ``rejuvenate.Proposer`` interface.

``Gen.metropolis_hastings(trace, tree_rejuvenation_proposal, (biased,), tree_rejuvenation_involution)``
(src/inference_smc_anneal_data.jl:90-96) accepts with probability

    min(1,  p(tree', xs) / p(tree, xs)  *  q(reverse choices | tree') / q(forward choices | tree)),

the model ratio being the weight of ``Gen.update`` (tree prior x likelihood).  ``rejuvenate.mh_structure_lockstep``
scores the likelihood ratio of all particles in ONE batched ``agp_lml_batch`` call; a proposer supplies the proposed
tree and the log of every other factor.  This module restates those other factors as plain densities instead of
Gen choice maps:

* the tree prior ``covariance_prior`` (src/Model.jl:66-127) with the node distributions of ``GPConfig``
  (src/GP.jl:1119-1131) — sampling and log-density, both over heap indices (children of ``idx`` are ``2 idx``,
  ``2 idx + 1``; depth ``1 + floor(log2 idx)``, GP.jl:1141);
* ``pick_random_node`` / ``generate_random_path`` (src/inference_utils.jl:16-89);
* SUBTREE-REPLACE (src/inference_rejuv_tree_sr.jl:17-88): a random node's subtree is redrawn from the prior;
* DETACH-ATTACH (src/inference_rejuv_tree_da.jl:17-281): a subtree b inside a subtree a takes a's place (detach), or
  a fresh auxiliary tree with a hole is put in a's place and a goes into the hole (attach);
* the mixture ``tree_rejuvenation_proposal`` (src/inference_rejuv_tree.jl:23-33): detach-attach with probability 1/2
  unless ``max_depth == 1``.

Trees carry PARAMETERS (``gp.Node`` fields); the prior is over the LATENTS ``z ~ normal(0, 1)`` of
``transform_param`` (Model.jl:35-48), so densities go through ``rejuvenate.latents``.  A ChangePoint's scale is the
constant .001 (Model.jl:121).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import gp, model

# integer codes of GPConfig (GP.jl:1101-1108), 1-based like the reference; index c - 1 into the distributions
CONSTANT, LINEAR, SQUARED_EXPONENTIAL, GAMMA_EXPONENTIAL, PERIODIC, PLUS, TIMES, CHANGEPOINT = range(1, 9)
_CODE_TO_TYPE = {CONSTANT: gp.Constant, LINEAR: gp.Linear, SQUARED_EXPONENTIAL: gp.SquaredExponential,
                 GAMMA_EXPONENTIAL: gp.GammaExponential, PERIODIC: gp.Periodic, PLUS: gp.Plus, TIMES: gp.Times,
                 CHANGEPOINT: gp.ChangePoint}
_TYPE_TO_CODE = {t: c for c, t in _CODE_TO_TYPE.items()}
CHANGEPOINT_SCALE = 0.001   # Model.jl:121
_LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


def _normalize(v: Sequence[float]) -> Tuple[float, ...]:
    s = float(sum(v))
    return tuple(float(x) / s for x in v)


@dataclass(frozen=True)
class GPConfig:
    """The fields of ``GP.GPConfig`` (GP.jl:1099-1138) the tree prior and the proposals read."""
    node_dist_leaf: Tuple[float, ...] = _normalize([0., 1, 0, 1, 1])
    node_dist_nocp: Tuple[float, ...] = _normalize([0., 6, 0, 6, 6, 5, 5])
    node_dist_cp: Tuple[float, ...] = _normalize([0., 6, 0, 6, 6, 4, 4, 2])
    max_branch: int = 2
    max_depth: int = -1
    changepoints: bool = True
    noise: Optional[float] = None

    def __post_init__(self):
        if self.max_branch != 2:
            raise ValueError("kernel trees are binary (GP.jl:1126)")
        if not (self.max_depth == -1 or self.max_depth >= 1):
            raise ValueError("max_depth is -1 (unbounded) or >= 1")


def idx_to_depth(idx: int) -> int:   # GP.jl:1141
    return idx.bit_length()


def get_node_dist(idx: int, config: GPConfig) -> Tuple[float, ...]:
    """``Model.get_node_dist`` (Model.jl:66-76)."""
    d = idx_to_depth(idx)
    if not (config.max_depth == -1 or 1 <= d <= config.max_depth):
        raise ValueError(f"node index {idx} (depth {d}) is below max_depth = {config.max_depth}")
    if d == config.max_depth:
        return config.node_dist_leaf
    return config.node_dist_cp if config.changepoints else config.node_dist_nocp


def _log_categorical(code: int, dist: Sequence[float]) -> float:
    """``Gen.logpdf(categorical, code, dist)``: -Inf outside the support (a ChangePoint where the model's
    distribution has seven entries scores zero probability, Model.jl:110-113)."""
    if 1 <= code <= len(dist) and dist[code - 1] > 0.0:
        return math.log(dist[code - 1])
    return -math.inf


def _draw_categorical(dist: Sequence[float], rng: np.random.Generator) -> int:
    return 1 + int(rng.choice(len(dist), p=np.asarray(dist)))


def _log_std_normal(z: float) -> float:
    return -0.5 * z * z - _LOG_SQRT_2PI


def _node_latents(node: gp.Node) -> List[float]:
    """Latents of ONE node (not its subtree), in ``fieldnames`` order (Model.jl:90-94, 115-116)."""
    t = type(node)
    if t is gp.ChangePoint:
        return [model.untransform_param("location", node.location)]
    if t in (gp.Plus, gp.Times):
        return []
    if t not in _TYPE_TO_CODE:
        raise ValueError(f"{t.__name__} is not generated by covariance_prior")
    return [model.untransform_param(f, getattr(node, f)) for f in gp._LEAF_FIELDS[t][1]]


# ------------------------------------------------------------------------------------------------
# covariance_prior: sample and score
# ------------------------------------------------------------------------------------------------
def sample_tree_prior(idx: int, config: GPConfig, rng: np.random.Generator) -> gp.Node:
    """``covariance_prior(idx, config)`` (Model.jl:78-127).  Iterative construction would buy nothing: the prior's
    expected size is small (branch mass 10/28), and Python's recursion limit is far above any tree that scores."""
    code = _draw_categorical(get_node_dist(idx, config), rng)
    t = _CODE_TO_TYPE[code]
    if code <= PERIODIC:
        return t(*[model.transform_param(f, float(rng.standard_normal())) for f in gp._LEAF_FIELDS[t][1]])
    if code in (PLUS, TIMES):
        below = replace(config, changepoints=False)                       # Model.jl:101
        left = sample_tree_prior(2 * idx, below, rng)
        return t(left, sample_tree_prior(2 * idx + 1, below, rng))
    location = model.transform_param("location", float(rng.standard_normal()))
    left = sample_tree_prior(2 * idx, config, rng)
    return gp.ChangePoint(left, sample_tree_prior(2 * idx + 1, config, rng), location, CHANGEPOINT_SCALE)


def log_prior_tree(node: gp.Node, idx: int, config: GPConfig) -> float:
    """log density of the choices under ``covariance_prior(idx, config)`` that produce ``node``: node types and
    latents of the whole subtree.  -inf for a tree the prior cannot generate (too deep, ChangePoint under an
    operator, a type of zero mass)."""
    d = idx_to_depth(idx)
    if config.max_depth != -1 and d > config.max_depth:
        return -math.inf
    t = type(node)
    if t not in _TYPE_TO_CODE:
        return -math.inf
    lp = _log_categorical(_TYPE_TO_CODE[t], get_node_dist(idx, config))
    if lp == -math.inf:
        return lp
    lp += sum(_log_std_normal(z) for z in _node_latents(node))
    if isinstance(node, gp.LeafNode):
        return lp
    below = config if t is gp.ChangePoint else replace(config, changepoints=False)
    lp += log_prior_tree(node.left, 2 * idx, below)
    if lp == -math.inf:
        return lp
    return lp + log_prior_tree(node.right, 2 * idx + 1, below)


# ------------------------------------------------------------------------------------------------
# heap-index helpers
# ------------------------------------------------------------------------------------------------
def _path_bits(idx_from: int, idx_to: int) -> List[int]:
    """Child choices (0 = left, 1 = right) leading from heap index ``idx_from`` down to ``idx_to``."""
    k = idx_to.bit_length() - idx_from.bit_length()
    if k < 0 or (idx_to >> k) != idx_from:
        raise ValueError(f"index {idx_to} is not below {idx_from}")
    return [(idx_to >> (k - 1 - j)) & 1 for j in range(k)]


def subtree_at(root: gp.Node, idx: int) -> gp.Node:
    nd = root
    for b in _path_bits(1, idx):
        nd = nd.right if b else nd.left
    return nd


def replace_at(root: gp.Node, idx: int, new: gp.Node) -> gp.Node:
    """``replace_subtree_choices`` (inference_utils.jl:147-169) on trees: ``root`` with the subtree at ``idx`` replaced."""
    bits = _path_bits(1, idx)

    def go(nd, j):
        if j == len(bits):
            return new
        left, right = (nd.left, go(nd.right, j + 1)) if bits[j] else (go(nd.left, j + 1), nd.right)
        if isinstance(nd, gp.ChangePoint):
            return gp.ChangePoint(left, right, nd.location, nd.scale)
        return type(nd)(left, right)

    return go(root, 0)


# ------------------------------------------------------------------------------------------------
# pick_random_node / generate_random_path
# ------------------------------------------------------------------------------------------------
def _p_done(node: gp.Node, biased: bool, leaf: bool, noroot: bool) -> float:   # inference_utils.jl:16-20
    if isinstance(node, gp.LeafNode):
        if noroot:
            raise ValueError("Impossible pick_random_node call.")
        return 1.0
    return 0.0 if (noroot or leaf) else (0.5 if biased else 1.0 / gp.size(node))


def _p_left(node: gp.Node, biased: bool) -> float:   # inference_utils.jl:22-23
    return 0.5 if biased else gp.size(node.left) / (gp.size(node) - 1)


def pick_random_node(node: gp.Node, idx: int, biased: bool, rng: np.random.Generator, *, leaf: bool = False,
                     noroot: bool = False) -> Tuple[gp.Node, int, float]:
    """``pick_random_node`` (inference_utils.jl:26-59): (picked node, its heap index, log probability of the pick)."""
    lp = 0.0
    while True:
        pd = _p_done(node, biased, leaf, noroot)
        if rng.random() < pd:
            return node, idx, lp + math.log(pd)
        lp += math.log1p(-pd)
        pl = _p_left(node, biased)
        if rng.random() < pl:
            node, idx, lp = node.left, 2 * idx, lp + math.log(pl)
        else:
            node, idx, lp = node.right, 2 * idx + 1, lp + math.log1p(-pl)
        noroot = False


def log_pick_random_node(node: gp.Node, idx: int, target: int, biased: bool, *, leaf: bool = False,
                         noroot: bool = False) -> float:
    """log probability that ``pick_random_node`` started at (``node``, ``idx``) ends at heap index ``target``."""
    lp = 0.0
    for b in _path_bits(idx, target):
        if isinstance(node, gp.LeafNode):
            return -math.inf
        pd = _p_done(node, biased, leaf, noroot)
        pl = _p_left(node, biased)
        step = (1.0 - pd) * ((1.0 - pl) if b else pl)
        if step <= 0.0:
            return -math.inf
        lp += math.log(step)
        node = node.right if b else node.left
        noroot = False
    if isinstance(node, gp.LeafNode) and noroot:
        return -math.inf
    pd = _p_done(node, biased, leaf, noroot)
    return lp + math.log(pd) if pd > 0.0 else -math.inf


def _path_p_done(depth: int, max_depth: int, noroot: bool) -> float:   # inference_utils.jl:74
    return 0.0 if noroot else (1.0 if depth == max_depth else 0.5)


def generate_random_path(idx: int, max_depth: int, rng: np.random.Generator, *, noroot: bool = False) -> Tuple[int, float]:
    """``generate_random_path`` (inference_utils.jl:61-89): (heap index of the hole, log probability)."""
    lp = 0.0
    while True:
        d = idx_to_depth(idx)
        if not (max_depth == -1 or 1 <= d <= max_depth):
            raise ValueError("generate_random_path below max_depth")
        pd = _path_p_done(d, max_depth, noroot)
        if rng.random() < pd:
            return idx, lp + math.log(pd)
        lp += math.log1p(-pd) + math.log(0.5)
        idx = 2 * idx + (0 if rng.random() < 0.5 else 1)
        noroot = False


def log_generate_random_path(idx: int, hole: int, max_depth: int, *, noroot: bool = False) -> float:
    lp = 0.0
    for b in _path_bits(idx, hole):
        d = idx_to_depth(idx)
        if max_depth != -1 and d > max_depth:
            return -math.inf
        pd = _path_p_done(d, max_depth, noroot)
        if pd >= 1.0:
            return -math.inf
        lp += math.log1p(-pd) + math.log(0.5)
        idx = 2 * idx + b
        noroot = False
    d = idx_to_depth(idx)
    if max_depth != -1 and d > max_depth:
        return -math.inf
    pd = _path_p_done(d, max_depth, noroot)
    return lp + math.log(pd) if pd > 0.0 else -math.inf


# ------------------------------------------------------------------------------------------------
# SUBTREE-REPLACE
# ------------------------------------------------------------------------------------------------
def _changepoints_allowed_at(root: gp.Node, idx: int, config: GPConfig) -> bool:
    """inference_rejuv_tree_sr.jl:27-39: at the root, or below a ChangePoint parent."""
    if not config.changepoints:
        return False
    if idx == 1:
        return True
    return isinstance(subtree_at(root, idx >> 1), gp.ChangePoint)


def subtree_replace_proposal(root: gp.Node, rng: np.random.Generator, config: GPConfig, biased: bool) -> Tuple[gp.Node, float]:
    """One subtree-replace move: (proposed tree, log[prior ratio x reverse/forward proposal ratio]).

    Forward: pick a node, draw a new subtree from ``covariance_prior(idx, config')``.  Reverse: pick the same index in
    the new tree, draw the old subtree.  The subtree densities under the proposal are evaluated with the proposal's
    own config' (ChangePoints iff root or ChangePoint parent), the model's with the config the ancestors imply; in a
    tree the prior can generate the two coincide and only the pick probabilities remain — but they are computed
    separately, as Gen does, so a trace outside the prior's support is weighed the same way."""
    _, idx, lp_pick_fwd = pick_random_node(root, 1, biased, rng)
    cfg = replace(config, changepoints=_changepoints_allowed_at(root, idx, config))
    old_sub = subtree_at(root, idx)
    new_sub = sample_tree_prior(idx, cfg, rng)
    proposed = replace_at(root, idx, new_sub)
    lq_fwd = lp_pick_fwd + log_prior_tree(new_sub, idx, cfg)
    lq_bwd = log_pick_random_node(proposed, 1, idx, biased) + log_prior_tree(old_sub, idx, cfg)
    lp_new, lp_old = log_prior_tree(proposed, 1, config), log_prior_tree(root, 1, config)
    if lp_new == -math.inf or lq_bwd == -math.inf:
        return proposed, -math.inf
    return proposed, (lp_new - lp_old) + (lq_bwd - lq_fwd)


# ------------------------------------------------------------------------------------------------
# DETACH-ATTACH
# ------------------------------------------------------------------------------------------------
def _aux_node_dist(idx: int, on_path: Dict[int, bool], force_cp: bool, config: GPConfig):
    """``get_node_dist_attach_detach`` (inference_rejuv_tree_da.jl:17-44)."""
    dist = list(get_node_dist(idx, config))
    if idx not in on_path:
        return dist
    if on_path[idx]:
        return None                                   # the hole: no choices
    if force_cp:
        if not config.changepoints:
            raise ValueError("force_cp without changepoints")
        return [1.0 if c == CHANGEPOINT else 0.0 for c in range(1, len(dist) + 1)]
    n_leaf = len(config.node_dist_leaf)
    dist = [0.0 if c < n_leaf else p for c, p in enumerate(dist)]
    s = sum(dist)
    return [p / s for p in dist] if s > 0 else dist


def _path_dict(idx_a: int, hole: int) -> Dict[int, bool]:
    path, idx = {}, idx_a
    for b in _path_bits(idx_a, hole):
        path[idx] = False
        idx = 2 * idx + b
    path[hole] = True
    return path


_HOLE = object()


def sample_aux_tree(idx: int, on_path: Dict[int, bool], force_cp: bool, config: GPConfig, rng: np.random.Generator):
    """``covariance_proposal_attach_detach`` (inference_rejuv_tree_da.jl:46-90): a tree with ``_HOLE`` at the hole and
    its log density."""
    dist = _aux_node_dist(idx, on_path, force_cp, config)
    if dist is None:
        return _HOLE, 0.0
    code = _draw_categorical(dist, rng)
    lp = math.log(dist[code - 1])
    t = _CODE_TO_TYPE[code]
    if code <= PERIODIC:
        zs = [float(rng.standard_normal()) for _ in gp._LEAF_FIELDS[t][1]]
        return (t(*[model.transform_param(f, z) for f, z in zip(gp._LEAF_FIELDS[t][1], zs)]),
                lp + sum(_log_std_normal(z) for z in zs))
    if code in (PLUS, TIMES):
        below = replace(config, changepoints=False)
        left, ll = sample_aux_tree(2 * idx, on_path, force_cp, below, rng)
        right, lr = sample_aux_tree(2 * idx + 1, on_path, force_cp, below, rng)
        return _Open(t, left, right, None), lp + ll + lr
    z = float(rng.standard_normal())
    left, ll = sample_aux_tree(2 * idx, on_path, force_cp, config, rng)
    right, lr = sample_aux_tree(2 * idx + 1, on_path, force_cp, config, rng)
    return _Open(gp.ChangePoint, left, right, model.transform_param("location", z)), lp + _log_std_normal(z) + ll + lr


@dataclass
class _Open:
    """A branch of the auxiliary tree before the hole is filled."""
    t: type
    left: object
    right: object
    location: Optional[float]


def _fill(aux, plug: gp.Node) -> gp.Node:
    if aux is _HOLE:
        return plug
    if isinstance(aux, _Open):
        left, right = _fill(aux.left, plug), _fill(aux.right, plug)
        if aux.t is gp.ChangePoint:
            return gp.ChangePoint(left, right, aux.location, CHANGEPOINT_SCALE)
        return aux.t(left, right)
    return aux


def log_aux_tree(node: gp.Node, idx: int, on_path: Dict[int, bool], force_cp: bool, config: GPConfig) -> float:
    """log density under ``covariance_proposal_attach_detach`` of the nodes of ``node`` (rooted at ``idx``) OUTSIDE the
    hole's subtree — what the reverse of a detach move has to propose."""
    if config.max_depth != -1 and idx_to_depth(idx) > config.max_depth:
        return -math.inf
    dist = _aux_node_dist(idx, on_path, force_cp, config)
    if dist is None:
        return 0.0
    t = type(node)
    if t not in _TYPE_TO_CODE:
        return -math.inf
    code = _TYPE_TO_CODE[t]
    lp = _log_categorical(code, dist)
    if lp == -math.inf:
        return lp
    lp += sum(_log_std_normal(z) for z in _node_latents(node))
    if isinstance(node, gp.LeafNode):
        return lp
    if t is gp.ChangePoint and not config.changepoints:
        return -math.inf                                               # the @assert of :79
    below = config if t is gp.ChangePoint else replace(config, changepoints=False)
    lp += log_aux_tree(node.left, 2 * idx, on_path, force_cp, below)
    if lp == -math.inf:
        return lp
    return lp + log_aux_tree(node.right, 2 * idx + 1, on_path, force_cp, below)


def _max_depth_aux(config: GPConfig, height_a: int) -> int:   # inference_rejuv_tree_da.jl:146-147
    return -1 if config.max_depth == -1 else config.max_depth - (height_a - 1)


def _p_detach(tree_size: int) -> float:   # inference_rejuv_tree_da.jl:104
    return 0.0 if tree_size == 1 else 0.5


def detach_attach_proposal(root: gp.Node, rng: np.random.Generator, config: GPConfig, biased: bool,
                           noroot: bool = False) -> Tuple[gp.Node, float]:
    """One detach-attach move (inference_rejuv_tree_da.jl:92-281): (proposed tree, log[prior ratio x reverse/forward
    proposal ratio])."""
    n = gp.size(root)
    if n == 1 and config.max_depth == 1:
        raise ValueError("Cannot apply ATTACH-DETACH with config.max_depth = 1.")
    pd = _p_detach(n)
    lp_old = log_prior_tree(root, 1, config)
    if rng.random() < pd:
        # ---- DETACH: subtree b (inside subtree a) takes a's place
        node_a, idx_a, lp_a = pick_random_node(root, 1, biased, rng)
        if noroot and isinstance(node_a, gp.LeafNode):
            return root, -math.inf        # the reference errors here; tree_rejuvenation_proposal never passes noroot
        node_b, idx_b, lp_b = pick_random_node(node_a, idx_a, biased, rng, noroot=noroot)
        proposed = replace_at(root, idx_a, node_b)
        lq_fwd = math.log(pd) + lp_a + lp_b
        # reverse = ATTACH on the proposed tree: pick a, draw the path to b, draw what was discarded
        p_attach_new = 1.0 - _p_detach(gp.size(proposed))
        force_cp = isinstance(node_b, gp.ChangePoint)                 # root_type of the CURRENT tree at a (:155-156)
        on_path = _path_dict(idx_a, idx_b)
        lq_bwd = (math.log(p_attach_new) + log_pick_random_node(proposed, 1, idx_a, biased)
                  + log_generate_random_path(idx_a, idx_b, _max_depth_aux(config, gp.depth(node_b)), noroot=noroot)
                  + log_aux_tree(node_a, idx_a, on_path, force_cp, config))
    else:
        # ---- ATTACH: an auxiliary tree with a hole takes a's place, a goes into the hole
        node_a, idx_a, lp_a = pick_random_node(root, 1, biased, rng)
        max_aux = _max_depth_aux(config, gp.depth(node_a))
        idx_b, lp_path = generate_random_path(idx_a, max_aux, rng, noroot=noroot)
        on_path = _path_dict(idx_a, idx_b)
        force_cp = isinstance(node_a, gp.ChangePoint)
        aux, lp_aux = sample_aux_tree(idx_a, on_path, force_cp, config, rng)
        new_a = _fill(aux, node_a)
        proposed = replace_at(root, idx_a, new_a)
        lq_fwd = math.log(1.0 - pd) + lp_a + lp_path + lp_aux
        # reverse = DETACH on the proposed tree: pick a, then b inside it
        pd_new = _p_detach(gp.size(proposed))
        if pd_new == 0.0:
            # a one-node tree attached into its own place (hole at a): the reverse attach has p_attach = 1
            lq_bwd = -math.inf
        else:
            lq_bwd = (math.log(pd_new) + log_pick_random_node(proposed, 1, idx_a, biased)
                      + log_pick_random_node(new_a, idx_a, idx_b, biased, noroot=noroot))
    lp_new = log_prior_tree(proposed, 1, config)
    if lp_new == -math.inf or lq_bwd == -math.inf:
        return proposed, -math.inf
    return proposed, (lp_new - lp_old) + (lq_bwd - lq_fwd)


# ------------------------------------------------------------------------------------------------
# tree_rejuvenation_proposal
# ------------------------------------------------------------------------------------------------
def tree_rejuvenation_proposer(config: GPConfig = GPConfig(), biased: bool = False):
    """``tree_rejuvenation_proposal`` + ``tree_rejuvenation_involution`` (inference_rejuv_tree.jl:22-55) as a
    ``rejuvenate.Proposer``: the move type is its own reverse, so its probability cancels."""
    p_detach_attach = 0.5 if config.max_depth != 1 else 0.0

    def propose(node: gp.Node, rng: np.random.Generator) -> Tuple[gp.Node, float]:
        if rng.random() < p_detach_attach:
            return detach_attach_proposal(node, rng, config, biased, False)
        return subtree_replace_proposal(node, rng, config, biased)

    return propose
