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
