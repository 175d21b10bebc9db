"""Lock-step rejuvenation: every particle's proposal of one MCMC iteration scored by ONE batched GPU call.

The reference rejuvenates particle by particle inside ``Threads.@threads`` (src/inference_smc_anneal_data.jl:240):
``rejuvenate_particle_structure`` (:78-119) runs ``n_mcmc`` involutive-MH moves on the tree, each accepted move followed
by ``rejuvenate_particle_parameters`` (:33-76) = ``n_hmc`` rounds of ``Gen.hmc`` on the leaf parameters and on the
noise (L = 10 leapfrog steps, eps = 0.02, early exit after ``n_exit`` consecutive rejections).  Every leapfrog step is
one LML + gradient evaluation, every MH proposal one LML evaluation: the hot path of this repository.  Here the
loops are turned inside out — the same moves, but all particles advance together so that each step is a single
``agp_lml_grad_batch`` / ``agp_lml_batch`` call over the whole (shrinking) set of active particles.

What stays on the host is what the reference keeps in Gen traces: the latent ``z ~ normal(0, 1)`` of every leaf field
(src/Model.jl:90-97), of a ChangePoint's location (:115-116; its scale is the constant .001) and of the noise (:133),
the momenta, and the accept/reject bookkeeping.  Randomness is drawn from one generator per particle, so a particle's
chain does not depend on which other particles share its batch.

Deviation from the reference, stated once: a covariance that stops being positive definite inside a trajectory makes
Gen throw ``PosDefException`` out of ``fit_smc!``; a lock-step batch cannot unwind one particle, so that particle's
move is rejected (score −inf) and the others carry on.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import gp, model

_BINARY = (gp.Plus, gp.Times, gp.ChangePoint)


# ------------------------------------------------------------------------------------------------
# kernel tree <-> latent vector
# ------------------------------------------------------------------------------------------------
def parameter_fields(node: gp.Node) -> List[Tuple[str, bool]]:
    """(field name, has a latent?) for every slot of ``gp.encode_program(node)[2]``, in that order."""
    out: List[Tuple[str, bool]] = []
    for nd in gp.unroll(node):
        t = type(nd)
        if t is gp.ChangePoint:
            out += [("location", True), ("scale", False)]          # Model.jl:115-121: scale is fixed
        elif t is gp.WhiteNoise:
            out.append(("value", False))                           # never sampled by covariance_prior
        elif t in gp._LEAF_FIELDS:
            out += [(f, True) for f in gp._LEAF_FIELDS[t][1]]      # Model.jl:90-94: every field
    return out


def with_parameters(node: gp.Node, params: Sequence[float]) -> gp.Node:
    """The same tree with its parameters replaced (``encode_program`` order)."""
    it = iter(params)

    def build(nd):
        t = type(nd)
        if t in (gp.Plus, gp.Times):
            left = build(nd.left)
            return t(left, build(nd.right))
        if t is gp.ChangePoint:
            left = build(nd.left)
            right = build(nd.right)
            return gp.ChangePoint(left, right, float(next(it)), float(next(it)))
        return t(*[float(next(it)) for _ in gp._LEAF_FIELDS[t][1]])

    out = build(node)
    if next(it, None) is not None:
        raise ValueError("more parameters than the tree has slots")
    return out


def latents(node: gp.Node) -> np.ndarray:
    """z of every latent slot: ``untransform_param`` (Model.jl:50-63) of the tree's parameters."""
    params = gp.encode_program(node)[2]
    return np.array([model.untransform_param(f, v) for (f, lat), v in zip(parameter_fields(node), params) if lat])


def with_latents(node: gp.Node, z: Sequence[float], fields=None, params=None) -> gp.Node:
    """The tree with its latents replaced.  ``fields`` = ``parameter_fields(node)`` and ``params`` = the tree's parameter
    vector may be passed by a caller that moves the same structure many times (one HMC trajectory = L calls per particle)."""
    params = (gp.encode_program(node)[2] if params is None else params).copy()
    it = iter(z)
    for j, (f, lat) in enumerate(parameter_fields(node) if fields is None else fields):
        if lat:
            params[j] = model.transform_param(f, float(next(it)))
    return with_parameters(node, params)


def latent_gradient(node: gp.Node, z: Sequence[float], grad_params: np.ndarray, fields=None) -> np.ndarray:
    """dLML/dz from dLML/dparams (chain rule through ``transform_param``)."""
    it = iter(z)
    out = []
    for (f, lat), g in zip(parameter_fields(node) if fields is None else fields, grad_params):
        if lat:
            out.append(g * model.transform_param_grad(f, float(next(it))))
    return np.array(out)


def noise_of(z_noise: float) -> float:
    return model.transform_param("noise", z_noise) + model.JITTER   # Model.jl:134


# ------------------------------------------------------------------------------------------------
# state
# ------------------------------------------------------------------------------------------------
@dataclass
class Chains:
    """What the reference keeps in P Gen traces, for the moves of this module."""
    nodes: List[gp.Node]
    z_noise: np.ndarray                       # latent of the noise, [P]
    fixed_noises: Optional[np.ndarray] = None # infer_noise=False with given noises: the values used as they are ([P];
                                              # z_noise is then never read: a noise <= JITTER has no latent)
    lml: Optional[np.ndarray] = None          # cached score / gradients of the CURRENT state (None: unknown)
    grad_z: Optional[List[np.ndarray]] = None
    grad_zn: Optional[np.ndarray] = None
    n_calls: int = 0                          # batched GPU calls issued so far
    n_evals: int = 0                          # particle evaluations inside them
    n_noise_only_calls: int = 0               # of n_calls: the cheaper noise-gradient-only calls
    stats: dict = field(default_factory=lambda: {"mh": 0, "mh_trials": 0, "hmc": 0, "hmc_trials": 0, "not_pd": 0})

    @property
    def P(self) -> int:
        return len(self.nodes)

    def noises(self) -> np.ndarray:
        return self.noises_at(np.arange(self.P), self.z_noise)

    def noises_at(self, who, z_noise) -> np.ndarray:
        """Noise of particles ``who`` at the candidate latents ``z_noise`` (or their fixed noises)."""
        if self.fixed_noises is not None:
            return np.asarray(self.fixed_noises, dtype=np.float64)[np.asarray(who, dtype=np.int64)]
        return np.array([noise_of(z) for z in z_noise])


def _evaluate(ch: Chains, who, nodes, z_list, z_noise, ts, xs, engine, noise_only: bool = False, fields_list=None):
    """LML and latent-space gradients of the given candidate states: one batched call.  ``noise_only``: the cheaper
    ``agp_lml_grad_noise_batch`` (no K^-1, no kernel-tree walk); the parameter gradients come back as None."""
    noises = list(ch.noises_at(who, z_noise))
    ch.n_calls += 1
    ch.n_evals += len(nodes)
    if noise_only:
        lml, gnoise, info = engine.lml_grad_noise_batch(nodes, noises, ts, xs)
        ch.n_noise_only_calls += 1
        gz = [None] * len(nodes)
    else:
        lml, gparams, gnoise, info = engine.lml_grad_batch(nodes, noises, ts, xs)
        gz = [latent_gradient(nd, z, g, None if fields_list is None else fields_list[a]) for a, (nd, z, g) in enumerate(zip(nodes, z_list, gparams))]
    gzn = np.array([g * model.transform_param_grad("noise", z) for g, z in zip(gnoise, z_noise)])
    ok = np.asarray(info) == 0
    ok &= np.isfinite(lml)
    return np.asarray(lml, dtype=np.float64), gz, gzn, ok


def refresh(ch: Chains, ts, xs, engine) -> None:
    """Score and gradient of the current state of every particle (the first ``choice_gradients`` of ``Gen.hmc``)."""
    zs = [latents(nd) for nd in ch.nodes]
    lml, gz, gzn, ok = _evaluate(ch, np.arange(ch.P), ch.nodes, zs, ch.z_noise, ts, xs, engine)
    if not ok.all():
        p = int(np.nonzero(~ok)[0][0])
        raise model.PosDefException(1, p)   # the CURRENT state must be scoreable, as in the reference
    ch.lml, ch.grad_z, ch.grad_zn = lml, gz, gzn


_LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


def _logpdf_std_normal(v: np.ndarray) -> float:
    return float(-0.5 * np.dot(v, v) - v.size * _LOG_SQRT_2PI)


# ------------------------------------------------------------------------------------------------
# Gen.hmc, all particles in lock step
# ------------------------------------------------------------------------------------------------
def hmc_lockstep(ch: Chains, active: np.ndarray, ts, xs, *, select: str, L: int, eps: float,
                 rngs: Sequence[np.random.Generator], engine) -> np.ndarray:
    """One ``Gen.hmc(trace, selection; L, eps)`` move for every particle in ``active`` (indices), the selection being
    the leaf/changepoint latents (``select="params"``) or the noise latent (``select="noise"``): momenta ~ N(0, I),
    L leapfrog steps ``p += eps/2·∇; z += eps·p; ∇ = ∇score(z); p += eps/2·∇``, accept with probability
    ``exp(score' − score + logN(p') − logN(p))`` where ``score = LML + Σ logN(z)`` over the selection (everything else
    in the trace is unchanged and cancels).  Returns the accepted mask over ``active``.  L batched calls (for the noise
    move L − 1 of them are the cheaper ``agp_lml_grad_noise_batch``)."""
    assert select in ("params", "noise")
    if ch.lml is None:
        refresh(ch, ts, xs, engine)
    active = np.asarray(active, dtype=np.int64)
    A = len(active)
    if A == 0:
        return np.zeros(0, dtype=bool)
    nodes = [ch.nodes[p] for p in active]
    z0 = [latents(nd) for nd in nodes]
    zn0 = ch.z_noise[active].copy()
    sel0 = z0 if select == "params" else [np.array([v]) for v in zn0]

    def grad_score(gz, gzn, sel):   # gradient of LML + log prior of the selected latents
        g = gz if select == "params" else [np.array([v]) for v in gzn]
        return [gi - si for gi, si in zip(g, sel)]

    sel = [s.copy() for s in sel0]
    grad = grad_score([ch.grad_z[p] for p in active], ch.grad_zn[active], sel)
    mom0 = [rngs[p].standard_normal(s.size) for p, s in zip(active, sel)]
    mom = [m.copy() for m in mom0]
    alive = np.ones(A, dtype=bool)              # trajectory still scoreable
    lml = ch.lml[active].copy()
    gz = [ch.grad_z[p] for p in active]
    gzn = ch.grad_zn[active].copy()
    cand_nodes = list(nodes)
    cheap_noise = select == "noise" and hasattr(engine, "lml_grad_noise_batch")
    # the structure of a particle's tree is fixed along the trajectory: its field list and parameter vector are derived once
    fields = [parameter_fields(nd) for nd in nodes]
    params0 = [gp.encode_program(nd)[2] for nd in nodes] if select == "params" else None
    for step in range(L):
        for a in range(A):
            if alive[a]:
                mom[a] = mom[a] + (eps / 2) * grad[a]
                sel[a] = sel[a] + eps * mom[a]
        # candidates (a dead trajectory keeps presenting its start state so that the batch stays scoreable)
        if select == "params":
            zs = [sel[a] if alive[a] else z0[a] for a in range(A)]
            try_nodes = []
            for a in range(A):
                try:
                    try_nodes.append(with_latents(nodes[a], zs[a], fields[a], params0[a]))
                except (AssertionError, OverflowError, ValueError):   # e.g. gamma rounded to 0 or out of (0, 2]
                    alive[a] = False
                    zs[a] = z0[a]
                    try_nodes.append(nodes[a])
            cand_nodes, zns = try_nodes, zn0
        else:
            zs = z0
            for a in range(A):
                if alive[a] and not (abs(sel[a][0]) < 700.0):      # exp(-1.5 + z) would overflow / NaN
                    alive[a] = False
            zns = np.array([sel[a][0] if alive[a] else zn0[a] for a in range(A)])
        # a noise trajectory needs dLML/dnoise only; its LAST evaluation is a full one so that the cached parameter
        # gradients of an accepted state are current for the parameter move that follows
        lml, gz, gzn, ok = _evaluate(ch, active, cand_nodes, zs, zns, ts, xs, engine, noise_only=cheap_noise and step < L - 1, fields_list=fields)
        newly_dead = alive & ~ok
        ch.stats["not_pd"] += int(newly_dead.sum())
        alive &= ok
        grad = grad_score(gz, gzn, sel)
        for a in range(A):
            if alive[a]:
                mom[a] = mom[a] + (eps / 2) * grad[a]
    accepted = np.zeros(A, dtype=bool)
    for a, p in enumerate(active):
        u = rngs[p].random()                    # drawn whether or not the move can be accepted: keeps the streams aligned
        if not alive[a]:
            continue
        new_score = lml[a] + _logpdf_std_normal(sel[a])
        old_score = ch.lml[p] + _logpdf_std_normal(sel0[a])
        alpha = new_score - old_score + _logpdf_std_normal(mom[a]) - _logpdf_std_normal(mom0[a])
        if math.log(u) < alpha:
            accepted[a] = True
            if select == "params":
                ch.nodes[p] = cand_nodes[a]
            else:
                ch.z_noise[p] = sel[a][0]
            ch.lml[p], ch.grad_z[p], ch.grad_zn[p] = lml[a], gz[a], gzn[a]
    return accepted


def rejuvenate_parameters_lockstep(ch: Chains, particles, n_hmc: int, ts, xs, *, rngs, engine,
                                   hmc_config: Optional[dict] = None, infer_noise: bool = True):
    """``rejuvenate_particle_parameters`` (inference_smc_anneal_data.jl:33-76) for the given particles at once.
    A particle leaves the loop after ``n_exit`` consecutive rejections of its parameter move (:69-71); the batch shrinks
    accordingly.  Returns (n_accept, n_trial) per particle."""
    cfg = hmc_config or {}
    L_param, eps_param = cfg.get("L_param", 10), cfg.get("eps_param", 0.02)
    L_noise, eps_noise = cfg.get("L_noise", 10), cfg.get("eps_noise", 0.02)
    n_exit = cfg.get("n_exit", n_hmc)
    particles = np.asarray(particles, dtype=np.int64)
    for p in particles:
        if not any(lat for _, lat in parameter_fields(ch.nodes[p])):
            raise AssertionError("length(leaf_addrs) > 0")           # :60
    n_accept = {int(p): 0 for p in particles}
    n_trial = {int(p): 0 for p in particles}
    n_reject = {int(p): 0 for p in particles}
    active = particles.copy()
    for _ in range(n_hmc):
        if len(active) == 0:
            break
        acc = hmc_lockstep(ch, active, ts, xs, select="params", L=L_param, eps=eps_param, rngs=rngs, engine=engine)
        if infer_noise:
            hmc_lockstep(ch, active, ts, xs, select="noise", L=L_noise, eps=eps_noise, rngs=rngs, engine=engine)
        keep = []
        for a, p in enumerate(active.tolist()):
            n_trial[p] += 1
            n_accept[p] += int(acc[a])
            n_reject[p] = 0 if acc[a] else n_reject[p] + 1
            if n_reject[p] != n_exit:
                keep.append(p)
        active = np.array(keep, dtype=np.int64)
    ch.stats["hmc"] += sum(n_accept.values())
    ch.stats["hmc_trials"] += sum(n_trial.values())
    return n_accept, n_trial


# ------------------------------------------------------------------------------------------------
# Gen.map_optimize, all traces in lock step (greedy search: src/Greedy.jl:93-101, 366-374)
# ------------------------------------------------------------------------------------------------
def _log_prior(z: np.ndarray, zn: float, infer_noise: bool) -> float:
    return _logpdf_std_normal(z) + (_logpdf_std_normal(np.array([zn])) if infer_noise else 0.0)


def map_optimize_lockstep(ch: Chains, particles, ts, xs, *, engine, max_opt: int = 500, max_step_size: float = 0.1,
                          tau: float = 0.5, min_step_size: float = 1e-16, infer_noise: bool = True) -> dict:
    """The parameter optimisation of the greedy search for many candidate traces at once:

        for i = 1:MAX_OPT;  trace = Gen.map_optimize(trace, select(noise, leaf latents...));  stop when the score stops changing

    (src/Greedy.jl:93-101, 366-374), where ``Gen.map_optimize`` is one gradient step with backtracking line search
    (step = max_step_size, halved by ``tau`` until the score does not get worse, given up below ``min_step_size``).
    score = LML + log N(z; 0, I) over the selected latents.  Every backtracking trial of all still-searching traces is one
    ``agp_lml_batch`` call, every accepted step one ``agp_lml_grad_batch`` call for the traces that moved.
    Returns {particle: (iterations, final score)}."""
    if ch.lml is None:
        refresh(ch, ts, xs, engine)
    active = [int(p) for p in particles]
    iters = {p: 0 for p in active}
    z = {p: latents(ch.nodes[p]) for p in active}
    score = {p: ch.lml[p] + _log_prior(z[p], ch.z_noise[p], infer_noise) for p in active}
    for _ in range(max_opt):
        if not active:
            break
        # ---- one Gen.map_optimize per active trace: backtracking line search along the gradient of the score
        grad = {p: ch.grad_z[p] - z[p] for p in active}
        gradn = {p: (ch.grad_zn[p] - ch.z_noise[p]) if infer_noise else 0.0 for p in active}
        step = {p: max_step_size for p in active}
        searching = list(active)
        new_state = {}
        while searching:
            cands, zs, zns, who = [], [], [], []
            for p in searching:
                zc, znc = z[p] + grad[p] * step[p], ch.z_noise[p] + gradn[p] * step[p]
                try:
                    if not abs(znc) < 700.0:
                        raise OverflowError
                    cands.append(with_latents(ch.nodes[p], zc))
                except (AssertionError, OverflowError, ValueError):
                    cands.append(None)
                zs.append(zc)
                zns.append(znc)
                who.append(p)
            ok_idx = [i for i, c in enumerate(cands) if c is not None]
            lml = np.full(len(cands), -np.inf)
            if ok_idx:
                l2, info = engine.lml_batch([cands[i] for i in ok_idx], list(ch.noises_at([who[i] for i in ok_idx], [zns[i] for i in ok_idx])), ts, xs)
                ch.n_calls += 1
                ch.n_evals += len(ok_idx)
                for j, i in enumerate(ok_idx):
                    if info[j] == 0 and np.isfinite(l2[j]):
                        lml[i] = l2[j]
            still = []
            for i, p in enumerate(who):
                new_score = lml[i] + _log_prior(zs[i], zns[i], infer_noise)
                if new_score - score[p] >= 0.0:                 # it got better (or stayed): take it
                    new_state[p] = (cands[i], zs[i], zns[i], new_score)
                elif step[p] < min_step_size:                   # it got worse and the step is exhausted: keep the trace
                    pass
                else:
                    step[p] *= tau
                    still.append(p)
            searching = still
        # ---- commit, and stop the traces whose score did not change (Greedy.jl:97-100)
        nxt = []
        for p in active:
            iters[p] += 1
            if p in new_state:
                nd, zc, znc, sc = new_state[p]
                changed = sc != score[p]
                ch.nodes[p], z[p], score[p] = nd, zc, sc
                ch.z_noise[p] = znc
                if changed:
                    nxt.append(p)
        active = nxt
        if active:   # gradients at the new points: one batched call
            nodes = [ch.nodes[p] for p in active]
            l2, gz, gzn, ok = _evaluate(ch, active, nodes, [z[p] for p in active], ch.z_noise[active], ts, xs, engine)
            for a, p in enumerate(active):
                ch.lml[p], ch.grad_z[p], ch.grad_zn[p] = l2[a], gz[a], gzn[a]
    # traces that finished on an accepted step of equal score keep a cache from before that step: refresh lazily
    stale = [p for p in iters if abs((ch.lml[p] + _log_prior(z[p], ch.z_noise[p], infer_noise)) - score[p]) > 0.0]
    if stale:
        nodes = [ch.nodes[p] for p in stale]
        l2, gz, gzn, ok = _evaluate(ch, stale, nodes, [z[p] for p in stale], ch.z_noise[stale], ts, xs, engine)
        for a, p in enumerate(stale):
            ch.lml[p], ch.grad_z[p], ch.grad_zn[p] = l2[a], gz[a], gzn[a]
    return {p: (iters[p], score[p]) for p in iters}


# ------------------------------------------------------------------------------------------------
# structure moves: one batched score per MH iteration
# ------------------------------------------------------------------------------------------------
# A proposer returns, for one particle, the proposed tree and the log of every factor of the MH ratio EXCEPT the
# likelihood ratio: prior ratio x backward/forward proposal ratio (x Jacobian).  That is where the reference's
# tree_rejuvenation_proposal / involution (src/inference_rejuv_tree_sr.jl) plug in; the likelihood ratio — the
# expensive part — is what this module batches.
Proposer = Callable[[gp.Node, np.random.Generator], Tuple[gp.Node, float]]


def mh_structure_lockstep(ch: Chains, particles, propose: Proposer, ts, xs, *, rngs, engine) -> np.ndarray:
    """One ``Gen.metropolis_hastings(trace, proposal, involution)`` (inference_smc_anneal_data.jl:90-96) per particle;
    all proposed trees are scored by one ``agp_lml_batch`` call.  Returns the accepted mask over ``particles``."""
    if ch.lml is None:
        refresh(ch, ts, xs, engine)
    particles = np.asarray(particles, dtype=np.int64)
    if len(particles) == 0:
        return np.zeros(0, dtype=bool)
    props = [propose(ch.nodes[p], rngs[p]) for p in particles]
    noises = list(ch.noises_at(particles, ch.z_noise[particles]))
    lml, info = engine.lml_batch([nd for nd, _ in props], noises, ts, xs)
    ch.n_calls += 1
    ch.n_evals += len(props)
    accepted = np.zeros(len(particles), dtype=bool)
    for a, p in enumerate(particles):
        u = rngs[p].random()
        if info[a] != 0 or not np.isfinite(lml[a]):
            ch.stats["not_pd"] += 1
            continue
        if math.log(u) < lml[a] - ch.lml[p] + props[a][1]:
            accepted[a] = True
            ch.nodes[p] = props[a][0]
    ch.stats["mh"] += int(accepted.sum())
    ch.stats["mh_trials"] += len(particles)
    if accepted.any():
        # gradients of the new trees are needed by the parameter moves that follow: one batched call for the
        # accepted particles only
        idx = particles[accepted]
        nodes = [ch.nodes[p] for p in idx]
        l2, gz, gzn, ok = _evaluate(ch, idx, nodes, [latents(nd) for nd in nodes], ch.z_noise[idx], ts, xs, engine)
        for a, p in enumerate(idx):
            ch.lml[p], ch.grad_z[p], ch.grad_zn[p] = l2[a], gz[a], gzn[a]
    return accepted


def rejuvenate_structure_lockstep(ch: Chains, n_mcmc: int, n_hmc: int, propose: Proposer, ts, xs, *, seed: int,
                                  engine: Optional[gp.Engine] = None, hmc_config: Optional[dict] = None,
                                  infer_noise: bool = True, rngs=None) -> dict:
    """``rejuvenate_particle_structure`` (inference_smc_anneal_data.jl:78-119) for all particles in lock step: per
    iteration one batched score of the P proposals, then the parameter moves of the particles that accepted."""
    engine = engine or gp.default_engine()
    rngs = rngs or particle_rngs(seed, ch.P)
    everyone = np.arange(ch.P)
    for _ in range(n_mcmc):
        acc = mh_structure_lockstep(ch, everyone, propose, ts, xs, rngs=rngs, engine=engine)
        if acc.any() and n_hmc > 0:
            rejuvenate_parameters_lockstep(ch, everyone[acc], n_hmc, ts, xs, rngs=rngs, engine=engine,
                                           hmc_config=hmc_config, infer_noise=infer_noise)
    return ch.stats


def particle_rngs(seed: int, P: int) -> List[np.random.Generator]:
    return [np.random.default_rng([int(seed), p]) for p in range(P)]


# ------------------------------------------------------------------------------------------------
# a minimal structure proposer (leaf swap): enough to drive the loop; the reference's subtree-replace / detach-attach
# involutions plug into the same interface from the Julia side
# ------------------------------------------------------------------------------------------------
_LEAF_DIST = {gp.Linear: 1 / 3, gp.GammaExponential: 1 / 3, gp.Periodic: 1 / 3}   # node_dist_leaf, GP.jl:1121


def sample_leaf_from_prior(rng: np.random.Generator) -> gp.Node:
    types = list(_LEAF_DIST)
    t = types[int(rng.choice(len(types), p=list(_LEAF_DIST.values())))]
    fields = gp._LEAF_FIELDS[t][1]
    return t(*[model.transform_param(f, float(rng.standard_normal())) for f in fields])


def leaf_swap_proposal(node: gp.Node, rng: np.random.Generator) -> Tuple[gp.Node, float]:
    """Pick uniformly one of the leaves whose type the prior can generate and redraw it (type and latents) from the
    prior.  The number of such leaves does not change, the three types have equal mass under every node distribution
    of GPConfig (GP.jl:1121-1123) and the new latents come from their prior, so prior ratio x backward/forward proposal
    ratio = 1: the returned log factor is 0 and the move is accepted on the likelihood ratio alone."""
    order = gp.unroll(node)
    leaves = [i for i, nd in enumerate(order) if type(nd) in _LEAF_DIST]
    if not leaves:
        return node, -math.inf
    target = leaves[int(rng.integers(len(leaves)))]
    new_leaf = sample_leaf_from_prior(rng)
    counter = iter(range(len(order)))

    def build(nd):   # walks the tree in unroll (postfix) order
        if isinstance(nd, _BINARY):
            left = build(nd.left)
            right = build(nd.right)
            next(counter)
            if isinstance(nd, gp.ChangePoint):
                return gp.ChangePoint(left, right, nd.location, nd.scale)
            return type(nd)(left, right)
        return new_leaf if next(counter) == target else nd

    return build(node), 0.0
