"""Lock-step rejuvenation loops (autogp.jl_b200/rejuvenate.py) against a per-particle replay.

The CPU tests drive the loops with a stand-in engine built on the oracle (LML by SciPy Cholesky, gradient by the
dense finite-difference route), the GPU test runs the same loops through the C-ABI and compares the chains."""
import math

import numpy as np
import pytest

import autogp_oracle as o
import autogp.jl_b200 as agp
from autogp.jl_b200 import rejuvenate as rj
from helpers import OracleEngine, OracleEngineWithNoiseCall, from_agp, to_agp


def start_state(P, seed=5):
    rng = np.random.default_rng(seed)
    trees = [
        lambda: agp.Plus(agp.Times(rj.sample_leaf_from_prior(rng), rj.sample_leaf_from_prior(rng)), rj.sample_leaf_from_prior(rng)),
        lambda: rj.sample_leaf_from_prior(rng),
        lambda: agp.ChangePoint(rj.sample_leaf_from_prior(rng), rj.sample_leaf_from_prior(rng),
                                agp.transform_param("location", float(rng.standard_normal())), 0.001),
    ]
    nodes = [trees[p % len(trees)]() for p in range(P)]
    return nodes, rng.standard_normal(P) * 0.5


def series(n):
    ts, xs = o.synthetic_series(n)
    return ts, xs


def test_latent_vector_round_trip_and_slot_order():
    nd = to_agp(o.ChangePoint(o.Plus(o.Linear(0.1, 1.3, 0.7), o.GammaExponential(0.42, 0.58, 3.2)),
                              o.Times(o.Periodic(0.96, 0.21, 1.1), o.WhiteNoise(0.3)), 0.5, 0.001))
    fields = rj.parameter_fields(nd)
    params = agp.encode_program(nd)[2]
    assert len(fields) == len(params)
    assert [f for f, _ in fields] == ["intercept", "bias", "amplitude", "lengthscale", "gamma", "amplitude",
                                      "lengthscale", "period", "amplitude", "value", "location", "scale"]
    assert [lat for _, lat in fields] == [True] * 9 + [False, True, False]
    z = rj.latents(nd)
    assert z.shape == (10,)
    back = agp.encode_program(rj.with_latents(nd, z))[2]
    np.testing.assert_allclose(back, params, rtol=1e-14)
    # latents are the N(0,1) choices of the reference's prior: transform_param of them gives the parameters
    assert math.isclose(agp.transform_param("gamma", z[4]), 0.58, rel_tol=1e-14)
    shifted = rj.with_latents(nd, z + 0.1)
    assert agp.encode_program(shifted)[2][9] == 0.3 and agp.encode_program(shifted)[2][11] == 0.001   # fixed slots untouched
    with pytest.raises(ValueError):
        rj.with_parameters(nd, list(params) + [1.0])


def test_latent_gradient_is_the_chain_rule():
    nd = agp.Plus(agp.Periodic(0.5, 0.2, 1.5), agp.GammaExponential(0.3, 1.2, 0.8))
    ts, xs = series(20)
    z = rj.latents(nd)
    g, _ = o.lml_grad_dense_fd(from_agp(nd), 0.1, ts, xs)
    gz = rj.latent_gradient(nd, z, g)
    for j in range(len(z)):
        f = lambda v: o.log_marginal_likelihood(from_agp(rj.with_latents(nd, np.r_[z[:j], v, z[j + 1:]])), 0.1, ts, xs)
        fd = (f(z[j] + 1e-5) - f(z[j] - 1e-5)) / 2e-5
        assert abs(fd - gz[j]) <= 1e-5 * max(1.0, abs(fd))


def run_joint(P, n_hmc, cfg, seed, n=20, engine=None, particles=None):
    nodes, zn = start_state(P)
    ts, xs = series(n)
    ch = rj.Chains(list(nodes), zn.copy())
    eng = engine or OracleEngine()
    rngs = rj.particle_rngs(seed, P)
    idx = np.arange(P) if particles is None else np.asarray(particles)
    acc = rj.rejuvenate_parameters_lockstep(ch, idx, n_hmc, ts, xs, rngs=rngs, engine=eng, hmc_config=cfg)
    return ch, acc, eng


def test_lockstep_hmc_equals_particle_by_particle_chains():
    P, n_hmc, cfg, seed = 5, 2, {"L_param": 3, "L_noise": 2, "eps_param": 0.05, "eps_noise": 0.05}, 11
    ch, (n_acc, n_trial), _ = run_joint(P, n_hmc, cfg, seed)
    nodes0, zn0 = start_state(P)
    ts, xs = series(20)
    moved = 0
    for p in range(P):   # the reference's loop: one particle at a time (batch of one), same random stream
        one = rj.Chains([nodes0[p]], zn0[p:p + 1].copy())
        rngs = {0: rj.particle_rngs(seed, P)[p]}
        a1, t1 = rj.rejuvenate_parameters_lockstep(one, [0], n_hmc, ts, xs, rngs=rngs, engine=OracleEngine(), hmc_config=cfg)
        assert one.nodes[0] == ch.nodes[p]
        assert one.z_noise[0] == ch.z_noise[p]
        assert (a1[0], t1[0]) == (n_acc[p], n_trial[p])
        moved += nodes0[p] != ch.nodes[p]
    assert moved >= 3   # the default step size accepts most trajectories


def test_small_steps_conserve_the_hamiltonian():
    P = 3
    nodes, zn = start_state(P)
    ts, xs = series(20)
    ch = rj.Chains(list(nodes), zn.copy())
    eng = OracleEngine()
    acc = rj.hmc_lockstep(ch, np.arange(P), ts, xs, select="params", L=4, eps=1e-4, rngs=rj.particle_rngs(3, P), engine=eng)
    assert acc.all()
    assert [b for b in eng.batches] == [("grad", P)] * 5      # 1 refresh + L leapfrog evaluations, each ONE batch
    # cached score of the accepted state is the score of the state
    for p in range(P):
        assert math.isclose(ch.lml[p], o.log_marginal_likelihood(from_agp(ch.nodes[p]), rj.noise_of(ch.z_noise[p]), ts, xs), rel_tol=1e-12)
    acc = rj.hmc_lockstep(ch, np.arange(P), ts, xs, select="noise", L=2, eps=1e-4, rngs=rj.particle_rngs(4, P), engine=eng)
    assert acc.all() and len(eng.batches) == 7                 # no second refresh: the cache is current


def test_noise_moves_use_the_cheap_call_except_for_their_last_step():
    P, cfg, seed = 4, {"L_param": 2, "L_noise": 4, "eps_param": 0.03, "eps_noise": 0.03}, 6
    plain, acc_plain, _ = run_joint(P, 2, cfg, seed)
    eng = OracleEngineWithNoiseCall()
    cheap, acc_cheap, _ = run_joint(P, 2, cfg, seed, engine=eng)
    assert cheap.nodes == plain.nodes and np.array_equal(cheap.z_noise, plain.z_noise) and acc_cheap == acc_plain
    kinds = [k for k, _ in eng.batches]
    assert kinds == ["grad"] + (["grad"] * 2 + ["noise"] * 3 + ["grad"]) * 2
    assert cheap.n_noise_only_calls == 6 and cheap.n_calls == len(kinds)
    for p in range(P):   # the cache after a noise move holds the parameter gradients of the accepted state
        g, _ = o.lml_grad_dense_fd(from_agp(cheap.nodes[p]), rj.noise_of(cheap.z_noise[p]), *series(20))
        np.testing.assert_array_equal(cheap.grad_z[p], rj.latent_gradient(cheap.nodes[p], rj.latents(cheap.nodes[p]), g))


def test_consecutive_rejections_shrink_the_batch():
    # absurd step size: every trajectory is rejected, n_exit = 1 ends each particle after its first trial
    cfg = {"L_param": 1, "L_noise": 1, "eps_param": 50.0, "eps_noise": 1e-3, "n_exit": 1}
    ch, (n_acc, n_trial), eng = run_joint(4, 5, cfg, seed=2)
    assert all(v == 0 for v in n_acc.values()) and all(v == 1 for v in n_trial.values())
    assert eng.batches == [("grad", 4)] * 3                   # refresh + params step + noise step, then nobody is left
    nodes0, _ = start_state(4)
    assert ch.nodes == nodes0
    assert ch.stats["hmc_trials"] == 4


def test_not_positive_definite_trajectory_is_rejected_without_disturbing_the_others():
    P, cfg = 4, {"L_param": 2, "L_noise": 2, "eps_param": 0.05, "eps_noise": 0.05}
    clean, _, _ = run_joint(P, 1, cfg, seed=9)
    nodes0, zn0 = start_state(P)
    victim = nodes0[2]
    # particle 2 (the only ChangePoint of the start state) "fails" whenever it is evaluated away from its start state
    eng = OracleEngine(fail=lambda nd, nz: type(nd) is agp.ChangePoint and nd != victim)
    ch, (n_acc, _), _ = run_joint(P, 1, cfg, seed=9, engine=eng)
    assert ch.nodes[2] == nodes0[2] and n_acc[2] == 0 and ch.stats["not_pd"] >= 1
    for p in (0, 1, 3):
        assert ch.nodes[p] == clean.nodes[p] and ch.z_noise[p] == clean.z_noise[p]


def test_leaf_swap_proposal_keeps_the_tree_shape():
    rng = np.random.default_rng(0)
    nd = agp.ChangePoint(agp.Plus(agp.Linear(0.1), agp.Periodic(0.5, 0.2)), agp.GammaExponential(0.3, 1.0), 0.4, 0.001)
    for _ in range(50):
        new, logf = rj.leaf_swap_proposal(nd, rng)
        assert logf == 0.0
        old_order, new_order = agp.unroll(nd), agp.unroll(new)
        assert len(old_order) == len(new_order)
        changed = [i for i, (a, b) in enumerate(zip(old_order, new_order)) if isinstance(a, agp.LeafNode) and a != b]
        assert len(changed) == 1 and type(new_order[changed[0]]) in (agp.Linear, agp.GammaExponential, agp.Periodic)
        assert [type(a) for a in old_order if not isinstance(a, agp.LeafNode)] == [type(a) for a in new_order if not isinstance(a, agp.LeafNode)]
        assert new.location == 0.4 and new.scale == 0.001
    only_se = agp.SquaredExponential(0.3)
    assert rj.leaf_swap_proposal(only_se, rng) == (only_se, -math.inf)


def test_structure_loop_scores_all_proposals_in_one_call_per_iteration():
    P, seed = 6, 21
    nodes, zn = start_state(P)
    ts, xs = series(20)
    ch = rj.Chains(list(nodes), zn.copy())
    eng = OracleEngine()
    stats = rj.rejuvenate_structure_lockstep(ch, 3, 0, rj.leaf_swap_proposal, ts, xs, seed=seed, engine=eng)
    assert [b for b in eng.batches if b[0] == "lml"] == [("lml", P)] * 3
    assert stats["mh_trials"] == 3 * P and 0 < stats["mh"] < 3 * P
    # replay particle by particle: same proposals, same uniforms, same decisions
    for p in range(P):
        rng = rj.particle_rngs(seed, P)[p]
        cur, cur_lml = nodes[p], o.log_marginal_likelihood(from_agp(nodes[p]), rj.noise_of(zn[p]), ts, xs)
        for _ in range(3):
            prop, logf = rj.leaf_swap_proposal(cur, rng)
            l2 = o.log_marginal_likelihood(from_agp(prop), rj.noise_of(zn[p]), ts, xs)
            if math.log(rng.random()) < l2 - cur_lml + logf:
                cur, cur_lml = prop, l2
        assert cur == ch.nodes[p]
        assert math.isclose(cur_lml, ch.lml[p], rel_tol=1e-12)


def test_structure_loop_with_parameter_moves_runs_only_the_accepted_particles():
    P = 5
    nodes, zn = start_state(P)
    ts, xs = series(16)
    ch = rj.Chains(list(nodes), zn.copy())
    eng = OracleEngine()
    cfg = {"L_param": 2, "L_noise": 1, "eps_param": 0.02, "eps_noise": 0.02, "n_exit": 1}
    stats = rj.rejuvenate_structure_lockstep(ch, 2, 1, rj.leaf_swap_proposal, ts, xs, seed=4, engine=eng, hmc_config=cfg)
    assert stats["hmc_trials"] == stats["mh"]          # n_hmc = 1 trial per accepted structure move
    assert ch.n_calls == len(eng.batches) and ch.n_evals == sum(b for _, b in eng.batches)
    for p in range(P):   # cached scores are the scores of the final states
        assert math.isclose(ch.lml[p], o.log_marginal_likelihood(from_agp(ch.nodes[p]), rj.noise_of(ch.z_noise[p]), ts, xs), rel_tol=1e-12)


def _map_optimize_one(node, zn, ts, xs, max_opt, max_step_size=0.1, tau=0.5, min_step_size=1e-16):
    """Gen.map_optimize + the loop of Greedy.jl:93-101 for ONE trace, written out directly on the oracle."""
    def score_of(nd, z, zn_):
        return (o.log_marginal_likelihood(from_agp(nd), rj.noise_of(zn_), ts, xs)
                - 0.5 * np.dot(z, z) - z.size * 0.5 * math.log(2 * math.pi) - 0.5 * zn_ * zn_ - 0.5 * math.log(2 * math.pi))

    def grad_of(nd, z, zn_):
        g, gn = o.lml_grad_dense_fd(from_agp(nd), rj.noise_of(zn_), ts, xs)
        return rj.latent_gradient(nd, z, g) - z, gn * agp.transform_param_grad("noise", zn_) - zn_

    iters, z = 0, rj.latents(node)
    score = score_of(node, z, zn)
    g, gn = grad_of(node, z, zn)
    for _ in range(max_opt):
        iters += 1
        step, new = max_step_size, None
        while True:
            cz, czn = z + g * step, zn + gn * step
            cand = rj.with_latents(node, cz)
            new_score = score_of(cand, cz, czn)
            if new_score - score >= 0:
                new = (cand, cz, czn, new_score)
                break
            if step < min_step_size:
                break
            step *= tau
        if new is None:
            break
        unchanged = new[3] == score
        node, z, zn, score = new
        if unchanged:
            break
        g, gn = grad_of(node, z, zn)
    return node, zn, score, iters


def test_lockstep_map_optimize_equals_trace_by_trace_optimisation():
    P = 5
    nodes, zn = start_state(P, seed=8)
    ts, xs = series(24)
    ch = rj.Chains(list(nodes), zn.copy())
    eng = OracleEngine()
    out = rj.map_optimize_lockstep(ch, np.arange(P), ts, xs, engine=eng, max_opt=6)
    for p in range(P):
        nd, z1, sc, iters = _map_optimize_one(nodes[p], zn[p], ts, xs, 6)
        assert out[p][0] == iters
        assert ch.nodes[p] == nd and ch.z_noise[p] == z1
        assert abs(out[p][1] - sc) <= 1e-12 * abs(sc)
        start = o.log_marginal_likelihood(from_agp(nodes[p]), rj.noise_of(zn[p]), ts, xs) + rj._log_prior(rj.latents(nodes[p]), zn[p], True)
        assert out[p][1] >= start                                    # the score never gets worse
        # the cache describes the final state
        assert math.isclose(ch.lml[p], o.log_marginal_likelihood(from_agp(ch.nodes[p]), rj.noise_of(ch.z_noise[p]), ts, xs), rel_tol=1e-12)
    # every backtracking trial of all searching traces was ONE batched call
    assert all(k in ("lml", "grad") for k, _ in eng.batches) and max(b for _, b in eng.batches) <= P
    assert sum(1 for k, _ in eng.batches if k == "lml") >= max(v[0] for v in out.values())


@pytest.mark.gpu
def test_gpu_chains_follow_the_oracle_chains():
    """The same lock-step loops through the C-ABI and through the oracle stand-in: identical decisions, latents equal
    to 1e-6 (tolerance = the oracle's finite-difference gradient error accumulated over the trajectories)."""
    P, seed, n = 6, 13, 96
    nodes, zn = start_state(P)
    ts, xs = series(n)
    cfg = {"L_param": 5, "L_noise": 5, "eps_param": 0.02, "eps_noise": 0.02}
    out = []
    for eng in (agp.Engine(0), OracleEngine()):
        ch = rj.Chains(list(nodes), zn.copy())
        rj.rejuvenate_structure_lockstep(ch, 3, 2, rj.leaf_swap_proposal, ts, xs, seed=seed, engine=eng, hmc_config=cfg)
        out.append(ch)
    g, c = out
    assert g.stats == c.stats and g.n_calls == c.n_calls
    assert g.stats["hmc"] > 0 and g.stats["mh"] > 0
    for p in range(P):
        assert [type(a) for a in agp.unroll(g.nodes[p])] == [type(a) for a in agp.unroll(c.nodes[p])]
        np.testing.assert_allclose(rj.latents(g.nodes[p]), rj.latents(c.nodes[p]), atol=1e-6, rtol=0)
        assert abs(g.z_noise[p] - c.z_noise[p]) <= 1e-6
        assert abs(g.lml[p] - c.lml[p]) <= 1e-8 * abs(c.lml[p]) + 1e-6


@pytest.mark.gpu
def test_gpu_map_optimize_climbs_like_the_oracle_run():
    """Greedy search's parameter optimisation through the C-ABI.  One Gen.map_optimize step (gradient + backtracking)
    lands where the same step on the oracle stand-in lands (1e-5: the oracle's finite-difference gradient error times
    the step); over many steps the score never decreases and the reported score is the oracle's score of the final
    state.  (Whole trajectories are not compared: a halving decision at a near-tie may legitimately differ.)"""
    P, n = 5, 96
    nodes, zn = start_state(P, seed=8)
    ts, xs = series(n)

    def score_of(nd, z_noise):
        return o.log_marginal_likelihood(from_agp(nd), rj.noise_of(z_noise), ts, xs) + rj._log_prior(rj.latents(nd), z_noise, True)

    gpu = agp.Engine(0)
    one = []
    for eng in (gpu, OracleEngine()):
        ch = rj.Chains(list(nodes), zn.copy())
        one.append((ch, rj.map_optimize_lockstep(ch, np.arange(P), ts, xs, engine=eng, max_opt=1)))
    (g, og), (c, oc) = one
    for p in range(P):
        np.testing.assert_allclose(rj.latents(g.nodes[p]), rj.latents(c.nodes[p]), rtol=0, atol=1e-5)
        assert abs(g.z_noise[p] - c.z_noise[p]) <= 1e-5 and abs(og[p][1] - oc[p][1]) <= 1e-6 * abs(oc[p][1]) + 1e-6
    ch = rj.Chains(list(nodes), zn.copy())
    out = rj.map_optimize_lockstep(ch, np.arange(P), ts, xs, engine=gpu, max_opt=12)
    climbed = 0
    for p in range(P):
        start, final = score_of(nodes[p], zn[p]), score_of(ch.nodes[p], ch.z_noise[p])
        assert out[p][1] >= start - 1e-9 * abs(start) and out[p][1] >= og[p][1] - 1e-9 * abs(og[p][1])
        assert abs(out[p][1] - final) <= 1e-9 * abs(final)
        assert abs(ch.lml[p] + rj._log_prior(rj.latents(ch.nodes[p]), ch.z_noise[p], True) - out[p][1]) <= 1e-9 * abs(final)   # cache is current
        climbed += out[p][1] > start + 1e-3
    assert climbed >= 3
