"""CPU suite, part 3: the N>1 path (particle sharding + the single log-weight all-gather) on
world_size-2 gloo.  The GPU engine is replaced by a test double that scores a shard with the
oracle — only the host-side sharding / collective / replicated-resampling logic is under test."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import autogp_oracle as o
import helpers as H


class OracleEngine:
    """Test double with the Engine.lml_batch signature."""

    def __init__(self, lookup):
        self.lookup = lookup
        self.calls = []

    def lml_batch(self, nodes, noises, ts, xs):
        self.calls.append(len(nodes))
        lml = np.array([o.log_marginal_likelihood(self.lookup[id(nd)], nz, ts, xs) for nd, nz in zip(nodes, noises)])
        return lml, np.zeros(len(nodes), dtype=np.int32)


def _make_state(P):
    from autogp.jl_b200 import smc

    parts = [o.synthetic_particle(p, "se+wn") for p in range(P)]
    nodes = [H.to_agp(nd) for nd, _ in parts]
    lookup = {id(a): nd for a, (nd, _) in zip(nodes, parts)}
    return smc.ParticleState(nodes=nodes, noises=[nz for _, nz in parts]), lookup


def _run_rounds(P, group=None):
    from autogp.jl_b200 import smc

    state, lookup = _make_state(P)
    eng = OracleEngine(lookup)
    ts, xs = o.synthetic_series(48)
    out = []
    for step, seed in ((16, 5), (32, 6), (48, 7)):
        scores = smc.smc_step(state, ts[:step], xs[:step], engine=eng, group=group)
        ess = smc.effective_sample_size(state.log_weights)
        resampled = smc.maybe_resample(state, ess_threshold=P / 2, seed=seed)
        # resampling moves agp nodes around: keep the lookup in sync (ids are preserved)
        out.append((scores.copy(), state.log_weights.copy(), ess, resampled, state.log_ml_est,
                    [lookup[id(nd)] for nd in state.nodes]))
    return out, eng.calls


def _worker(rank, world, port, P, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out, calls = _run_rounds(P)
        q.put((rank, [(s.tolist(), w.tolist(), e, r, m, [repr(n) for n in nodes]) for s, w, e, r, m, nodes in out], calls))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("P", [6, 7])
def test_two_rank_smc_matches_single_process(P):
    ref, ref_calls = _run_rounds(P)
    assert ref_calls == [P, P, P]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    # each rank scored only its shard
    assert sorted(results[0][2] + results[1][2]) == sorted([P // 2] * 3 + [P - P // 2] * 3)
    for rank, rounds, _ in results:
        for got, want in zip(rounds, ref):
            s, w, e, r, m, nodes = got
            assert np.allclose(s, want[0], rtol=0, atol=0)
            assert np.allclose(w, want[1], rtol=0, atol=0)
            assert e == want[2] and r == want[3] and m == want[4]
            assert nodes == [repr(n) for n in want[5]]


# ---- a full SMC round with rejuvenation: reweight -> resample -> lock-step rejuvenation of the shard -> all-gather ----
def _run_rounds_with_rejuvenation(P, group=None):
    from autogp.jl_b200 import rejuvenate as rj, smc

    parts = [o.synthetic_particle(p, "ge+per*lin") for p in range(P)]
    state = smc.ParticleState(nodes=[H.to_agp(nd) for nd, _ in parts], noises=[nz for _, nz in parts])
    eng = H.OracleEngine()
    ts, xs = o.synthetic_series(36)
    cfg = {"L_param": 2, "L_noise": 2, "eps_param": 0.03, "eps_noise": 0.03, "n_exit": 1}
    out = []
    for step, seed in ((18, 5), (36, 6)):
        smc.smc_step(state, ts[:step], xs[:step], engine=eng, group=group)
        smc.maybe_resample(state, ess_threshold=P / 2, seed=seed)
        stats = smc.rejuvenate(state, ts[:step], xs[:step], n_mcmc=2, n_hmc=1, propose=rj.leaf_swap_proposal, seed=100 + seed,
                               engine=eng, group=group, hmc_config=cfg)
        out.append(([repr(nd) for nd in state.nodes], list(state.noises), state.scores.tolist(), state.log_weights.tolist(), stats))
    return out, sum(b for _, b in eng.batches)


def _worker_rejuv(rank, world, port, P, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank,) + _run_rounds_with_rejuvenation(P))
    finally:
        dist.destroy_process_group()


def test_two_rank_smc_round_with_rejuvenation_matches_single_process():
    P = 5
    ref, ref_evals = _run_rounds_with_rejuvenation(P)
    assert any(r[4]["mh"] > 0 for r in ref) and any(r[4]["hmc_trials"] > 0 for r in ref)
    # the scores the state carries after rejuvenation are the LMLs of the rejuvenated particles
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_rejuv, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rounds, evals in results:
        assert evals < ref_evals                       # each rank evaluated only its shard
        for got, want in zip(rounds, ref):
            assert got[0] == want[0] and got[1] == want[1]          # same kernels, same noises, bitwise
            assert got[2] == want[2] and got[3] == want[3]
            assert got[4] == want[4]                                # same accept / reject counts, summed over ranks
    assert sum(r[2] for r in results) == ref_evals


# ---- the whole loop (run_smc_anneal_data) with the reference's structure proposals, two ranks against one ----
def _run_whole_loop(P):
    from autogp.jl_b200 import smc, tree_moves as tm

    ts, xs = o.synthetic_series(30)
    state = smc.run_smc_anneal_data(ts, xs, config=tm.GPConfig(max_depth=2), n_particles=P, n_mcmc=2, n_hmc=1,
                                    schedule=[10, 20, 30], seed=9, engine=H.OracleEngine(),
                                    hmc_config={"L_param": 2, "L_noise": 2, "n_exit": 1})
    return [repr(nd) for nd in state.nodes], list(state.noises), state.scores.tolist(), state.log_weights.tolist(), state.log_ml_est


def _worker_loop(rank, world, port, P, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank,) + tuple(_run_whole_loop(P)))
    finally:
        dist.destroy_process_group()


def test_two_rank_run_smc_anneal_data_matches_single_process():
    P = 5
    ref = _run_whole_loop(P)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_loop, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for res in results:
        assert tuple(res[1:]) == tuple(ref)          # prior draws, proposals, decisions: identical on 1 and 2 ranks, bitwise
