"""Host-side mirror of the SMC reweight / resample step around the hot path.

Reference: ``smc_step!`` (src/inference_smc_anneal_data.jl:127-141), the weight consumers
``compute_particle_weights`` / ``effective_sample_size`` (:22-31) and Gen's
``maybe_resample!`` as called at :229-234.  Particles are independent between resampling
barriers, so they shard across GPUs (one process per GPU); the only exchange is ONE
all-gather of the per-particle log-weights, after which every rank resamples identically from
a shared seed.  No other collective exists on this path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import gp


def logsumexp(v: np.ndarray) -> float:
    v = np.asarray(v, dtype=np.float64)
    if v.size == 0:
        return -math.inf
    m = float(np.max(v))
    if not math.isfinite(m):
        return m
    return m + math.log(float(np.sum(np.exp(v - m))))


def normalize_weights(log_weights: np.ndarray):
    """``Gen.normalize_weights``: (log_total_weight, log_normalized_weights)."""
    lt = logsumexp(log_weights)
    return lt, np.asarray(log_weights, dtype=np.float64) - lt


def compute_particle_weights(log_weights: np.ndarray) -> np.ndarray:  # inference_smc_anneal_data.jl:22-25
    return np.exp(normalize_weights(log_weights)[1])


def effective_sample_size(log_weights: np.ndarray) -> float:  # inference_smc_anneal_data.jl:28-31
    lnw = normalize_weights(log_weights)[1]
    return math.exp(-logsumexp(2.0 * lnw))


def linear_schedule(n: int, percent: float) -> List[int]:
    """``Schedule.linear_schedule(n, percent)`` (src/Schedule.jl:24-39): data prefixes growing by about ``n * percent``."""
    if not (0 < n and 0 < percent < 1):
        raise ValueError("linear_schedule: need 0 < n and 0 < percent < 1")
    step = int(round(percent * n))  # Julia's round and Python's both round half to even
    checkpoints = list(range(step, n + 1, step))
    remaining = n - checkpoints[-1]
    if remaining == 0:
        return checkpoints
    if remaining < step / 2:
        checkpoints[-1] = n
        return checkpoints
    return checkpoints + [n]


def _dist():
    """``torch.distributed`` when this process has a process group, else None.  A process that never imported torch has
    none, and a single-GPU fit does not pay the import (about 2 s) for finding that out."""
    import sys

    if "torch" not in sys.modules:
        return None
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def shard_range(P: int, rank: int, world: int):
    """Contiguous block of particles owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class ParticleState:
    """The slice of ``Gen.ParticleFilterState`` this path needs: kernels, noises, weights."""
    nodes: List[gp.Node]
    noises: List[float]
    log_weights: np.ndarray = None
    scores: np.ndarray = None  # current LML of each particle (the trace score's likelihood part)
    log_ml_est: float = 0.0

    def __post_init__(self):
        P = len(self.nodes)
        if self.log_weights is None:
            self.log_weights = np.zeros(P)
        if self.scores is None:
            self.scores = np.zeros(P)  # n = 0: empty mvnormal scores 0 (:185-189)


def all_gather_log_weights(local: "np.ndarray | object", P: int, group=None, device_tensor=None) -> np.ndarray:
    """The single collective of the path: gather every rank's log-weights.

    With ``torch.distributed`` initialised this is one ``all_gather`` (NCCL over NVLink when the
    tensors live on the GPU, gloo in the CPU tests); single-process it is the identity.
    Shards may be ragged (P not divisible by the world size): shards are padded to the largest.
    """
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        if device_tensor is not None:
            return device_tensor.detach().cpu().numpy().copy()
        return np.asarray(local, dtype=np.float64).copy()
    import torch

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    width = -(-P // world)
    if device_tensor is None and _device_only_backend(dist, group):
        # an NCCL-only process group cannot move host tensors: stage the shard on this rank's GPU (the caller has set it
        # with torch.cuda.set_device, as every NCCL program must)
        device_tensor = torch.as_tensor(np.asarray(local, dtype=np.float64), device=f"cuda:{torch.cuda.current_device()}")
    if device_tensor is not None:
        src = device_tensor
        buf = torch.full((width,), float("nan"), dtype=torch.float64, device=src.device)
        buf[: src.numel()] = src
    else:
        buf = torch.full((width,), float("nan"), dtype=torch.float64)
        loc = torch.as_tensor(np.asarray(local, dtype=np.float64))
        buf[: loc.numel()] = loc
    out = torch.empty((world * width,), dtype=torch.float64, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy()
    pieces = []
    for r in range(world):
        lo, hi = shard_range(P, r, world)
        pieces.append(out[r * width: r * width + (hi - lo)])
    return np.concatenate(pieces)


def _device_only_backend(dist, group) -> bool:
    try:
        return "gloo" not in str(dist.get_backend(group)).lower() and "nccl" in str(dist.get_backend(group)).lower()
    except Exception:
        return False


def resample_indices(log_weights: np.ndarray, seed: int) -> np.ndarray:
    """Multinomial parent indices from normalised weights (Gen.maybe_resample! default),
    drawn from a seed shared by all ranks so the result is replicated without communication."""
    w = compute_particle_weights(log_weights)
    rng = np.random.default_rng(seed)
    return rng.choice(len(w), size=len(w), replace=True, p=w / w.sum())


def maybe_resample(state: ParticleState, ess_threshold: float, seed: int) -> bool:
    """``Gen.maybe_resample!(state, ess_threshold=...)`` on the replicated weight vector."""
    P = len(state.nodes)
    lt, lnw = normalize_weights(state.log_weights)
    ess = math.exp(-logsumexp(2.0 * lnw))
    if not (ess < ess_threshold):
        return False
    parents = resample_indices(state.log_weights, seed)
    state.log_ml_est += lt - math.log(P)
    state.nodes = [state.nodes[i] for i in parents]
    state.noises = [state.noises[i] for i in parents]
    state.scores = state.scores[parents].copy()
    state.log_weights = np.zeros(P)
    return True


def smc_step(state: ParticleState, ts, xs, *, engine: Optional[gp.Engine] = None, group=None) -> np.ndarray:
    """``smc_step!`` (:127-141): re-score every particle on the grown data prefix and add the
    incremental weight ``LML_new - score_old``.  Each rank scores its shard on its own GPU;
    the all-gather replicates the new scores so every rank holds the full weight vector."""
    dist = _dist()
    P = len(state.nodes)
    dist_on = dist is not None
    world = dist.get_world_size(group) if dist_on else 1
    rank = dist.get_rank(group) if dist_on else 0
    lo, hi = shard_range(P, rank, world)
    eng = engine or gp.default_engine()
    local, info = eng.lml_batch(state.nodes[lo:hi], state.noises[lo:hi], ts, xs)
    if np.any(info != 0):
        from .model import PosDefException
        bad = int(np.nonzero(info)[0][0])
        raise PosDefException(int(info[bad]), lo + bad)
    new_scores = all_gather_log_weights(local, P, group=group)
    state.log_weights = state.log_weights + (new_scores - state.scores)
    state.scores = new_scores
    return new_scores


def rejuvenate(state: ParticleState, ts, xs, *, n_mcmc: int, n_hmc: int, propose, seed: int,
               engine: Optional[gp.Engine] = None, group=None, hmc_config: Optional[dict] = None,
               infer_noise: bool = True, round_index: int = 0) -> dict:
    """Step 4 of ``run_smc_anneal_data`` (src/inference_smc_anneal_data.jl:236-250): ``rejuvenate_particle_structure``
    on every particle.  Each rank rejuvenates ITS shard in lock step on its own GPU (``rejuvenate.py``: one batched call
    per MH proposal / leapfrog step of the whole shard); particle p draws from the random stream (seed, p) whatever
    rank it lives on, so the result does not depend on the number of GPUs; ``round_index`` (the SMC round) enters the
    stream's seed, so a caller that passes the same ``seed`` every round still gets fresh draws.  The rejuvenated kernels are a few hundred
    bytes per particle: one all-gather of them replicates the state for the next resampling step.  MCMC moves leave
    the log-weights alone and replace the trace scores."""
    from . import model, rejuvenate as rj

    dist = _dist()
    P = len(state.nodes)
    dist_on = dist is not None
    world = dist.get_world_size(group) if dist_on else 1
    rank = dist.get_rank(group) if dist_on else 0
    lo, hi = shard_range(P, rank, world)
    eng = engine or gp.default_engine()
    if infer_noise:
        bad = [lo + a for a, nz in enumerate(state.noises[lo:hi]) if not nz > model.JITTER]
        if bad:
            raise ValueError(f"rejuvenate(infer_noise=True): noise of particle {bad[0]} is not above JITTER = {model.JITTER} "
                             "(the latent of src/Model.jl:133-134 does not exist); pass infer_noise=False for fixed noises")
        z_noise = np.array([model.untransform_param("noise", nz - model.JITTER) for nz in state.noises[lo:hi]])
    else:
        z_noise = np.zeros(hi - lo)  # never read: the chain's noise is carried through unchanged (below)
    ch = rj.Chains(list(state.nodes[lo:hi]), z_noise.copy(), fixed_noises=None if infer_noise else np.array(state.noises[lo:hi], dtype=np.float64))
    rngs = [np.random.default_rng([int(seed), int(round_index), p]) for p in range(lo, hi)]
    stats = rj.rejuvenate_structure_lockstep(ch, n_mcmc, n_hmc, propose, ts, xs, seed=seed, engine=eng, rngs=rngs,
                                             hmc_config=hmc_config, infer_noise=infer_noise) if hi > lo else dict(ch.stats)
    z0 = z_noise
    # a noise the chain did not move keeps its bits (exp(log(x)) is not always x)
    new_noises = [state.noises[lo + a] if z == z0[a] else rj.noise_of(z) for a, z in enumerate(ch.z_noise)]
    mine = (lo, ch.nodes, new_noises, None if ch.lml is None else ch.lml.tolist(), stats)
    if world > 1:
        shards = [None] * world
        dist.all_gather_object(shards, mine, group=group)
    else:
        shards = [mine]
    total = {}
    for lo_r, nodes, noises, lml, st in shards:
        for a, nd in enumerate(nodes):
            state.nodes[lo_r + a] = nd
            state.noises[lo_r + a] = noises[a]
            if lml is not None:
                state.scores[lo_r + a] = lml[a]
        for k, v in st.items():
            total[k] = total.get(k, 0) + v
    return total


def initialize_particles(n_particles: int, config, seed: int) -> ParticleState:
    """``Gen.initialize_particle_filter(model, (Float64[], config), observations, n_particles)``
    (src/inference_smc_anneal_data.jl:183-190): kernels from ``covariance_prior``, noise latents from ``normal(0, 1)``
    unless ``config.noise`` fixes the noise — then the constrained latent's prior density is every particle's initial
    log-weight, as ``Gen.generate`` returns it.  No data yet: the empty ``mvnormal`` scores 0.  Particle p draws from
    the stream (seed, p), whatever rank asks."""
    from . import model, tree_moves

    nodes, noises = [], []
    for p in range(n_particles):
        rng = np.random.default_rng([int(seed), 0x5EED, p])
        nodes.append(tree_moves.sample_tree_prior(1, config, rng))
        z = float(rng.standard_normal())
        noises.append(float(config.noise) if config.noise is not None
                      else model.transform_param("noise", z) + model.JITTER)
    state = ParticleState(nodes, noises)
    if config.noise is not None:
        z_fixed = model.untransform_param("noise", float(config.noise))     # :186 (no JITTER subtracted there either)
        state.log_weights[:] = -0.5 * z_fixed * z_fixed - 0.5 * math.log(2.0 * math.pi)
    return state


def run_smc_anneal_data(ts, xs, *, config=None, biased: bool = False, n_particles: int = 4, n_mcmc=10, n_hmc=10,
                        hmc_config: Optional[dict] = None, permutation: Optional[Sequence[int]] = None,
                        schedule: Optional[Sequence[int]] = None, adaptive_resampling: bool = True,
                        adaptive_rejuvenation: bool = False, seed: int = 0, engine: Optional[gp.Engine] = None,
                        group=None, callback_fn=None) -> ParticleState:
    """``run_smc_anneal_data`` (src/inference_smc_anneal_data.jl:143-273): SMC over growing data prefixes with the
    reference's structure proposals (``tree_moves``), every scoring step a batched GPU call — reweight
    (``smc_step``), resample on the replicated weights (``maybe_resample``; never after the last prefix), rejuvenate
    in lock step (``rejuvenate``).  ``permutation`` is 0-based.  With ``torch.distributed`` initialised every rank
    calls this with the same arguments and holds the same state afterwards."""
    from . import tree_moves

    config = config or tree_moves.GPConfig()
    ts, xs = np.asarray(ts, dtype=np.float64), np.asarray(xs, dtype=np.float64)
    n = ts.shape[0]
    if xs.shape != ts.shape:
        raise ValueError("ts and xs must have equal length")
    permutation = np.arange(n) if permutation is None else np.asarray(permutation, dtype=np.int64)
    if sorted(permutation.tolist()) != list(range(n)):
        raise ValueError("permutation must be a permutation of 0..n-1")
    ts, xs = ts[permutation], xs[permutation]
    schedule = list(range(1, n + 1)) if schedule is None else [int(s) for s in schedule]
    if not (schedule and 1 <= schedule[0] and schedule[-1] == n and all(b > a for a, b in zip(schedule, schedule[1:]))):
        raise ValueError("schedule must increase strictly from >= 1 to len(ts)")
    n_mcmc = [int(n_mcmc)] * len(schedule) if np.isscalar(n_mcmc) else [int(v) for v in n_mcmc]
    n_hmc = [int(n_hmc)] * len(schedule) if np.isscalar(n_hmc) else [int(v) for v in n_hmc]
    if len(n_mcmc) != len(schedule) or len(n_hmc) != len(schedule):
        raise ValueError("n_mcmc / n_hmc: one value, or one per schedule entry")
    propose = tree_moves.tree_rejuvenation_proposer(config, biased)
    engine = engine or gp.default_engine()
    if hasattr(engine, "reserve"):
        # the series grows round by round: size the factor (and the gradient calls' augmented one) once
        dist = _dist()
        world = dist.get_world_size(group) if dist is not None else 1
        try:
            engine.reserve(n, -(-n_particles // world), gradient=any(h > 0 for h in n_hmc))
        except Exception:   # no room for everything at once: the calls size what they need as they go
            pass
    state = initialize_particles(n_particles, config, seed)
    if callback_fn:
        callback_fn(state=state, ts=ts, xs=xs, step=0, rejuvenated=False, resampled=False, stats=None)
    for i, step in enumerate(schedule):
        ts_obs, xs_obs = ts[:step], xs[:step]
        smc_step(state, ts_obs, xs_obs, engine=engine, group=group)
        resampled = False
        if step < schedule[-1]:
            ess_threshold = n_particles / 2 if adaptive_resampling else n_particles
            resampled = maybe_resample(state, ess_threshold, seed=int(np.random.default_rng([int(seed), 0xE55, i]).integers(2 ** 62)))
        rejuvenated, stats = False, None
        if not adaptive_rejuvenation or resampled:
            rejuvenated = True
            stats = rejuvenate(state, ts_obs, xs_obs, n_mcmc=n_mcmc[i], n_hmc=n_hmc[i], propose=propose, seed=seed,
                               engine=engine, group=group, hmc_config=hmc_config, infer_noise=config.noise is None,
                               round_index=i + 1)
        if callback_fn:
            callback_fn(state=state, ts=ts, xs=xs, step=step, rejuvenated=rejuvenated, resampled=resampled, stats=stats)
    return state
