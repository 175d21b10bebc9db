"""CPU tests of the structure proposals (autogp.jl_b200/tree_moves.py): the reference's subtree-replace and detach-attach
involutions (src/inference_rejuv_tree_sr.jl, _da.jl) restated as (proposed tree, log ratio) pairs.

The decisive property: with a FLAT likelihood, Metropolis-Hastings with these proposals must leave the tree prior
``covariance_prior`` (src/Model.jl:78-127) invariant.  A wrong factor anywhere (pick probabilities, path
probabilities, the auxiliary tree's forced node types, the depth-dependent node distributions after a subtree changes
depth) shifts the chain's size / type statistics away from the prior's — which is what the chi-square tests look at.
"""
import math
from collections import Counter

import numpy as np
import pytest

from autogp.jl_b200 import gp, tree_moves as tm


def _all_indices(node, idx=1):
    out = [idx]
    if isinstance(node, gp.BinaryOpNode):
        out += _all_indices(node.left, 2 * idx) + _all_indices(node.right, 2 * idx + 1)
    return out


def _valid(node, config, idx=1, cp_ok=None):
    cp_ok = config.changepoints if cp_ok is None else cp_ok
    if config.max_depth != -1 and tm.idx_to_depth(idx) > config.max_depth:
        return False
    if isinstance(node, gp.ChangePoint):
        return cp_ok and _valid(node.left, config, 2 * idx, True) and _valid(node.right, config, 2 * idx + 1, True)
    if isinstance(node, (gp.Plus, gp.Times)):
        return _valid(node.left, config, 2 * idx, False) and _valid(node.right, config, 2 * idx + 1, False)
    return type(node) in (gp.Linear, gp.GammaExponential, gp.Periodic)


@pytest.mark.parametrize("config", [tm.GPConfig(), tm.GPConfig(max_depth=3), tm.GPConfig(changepoints=False, max_depth=4)])
def test_prior_samples_are_valid_and_score_finite(config):
    rng = np.random.default_rng(1)
    for _ in range(300):
        t = tm.sample_tree_prior(1, config, rng)
        assert _valid(t, config)
        assert math.isfinite(tm.log_prior_tree(t, 1, config))


def test_prior_density_of_a_known_tree():
    # Plus(Linear, Periodic) under the default config: root from node_dist_cp (Plus: 4/28), children from node_dist_nocp
    # (Linear 6/28, Periodic 6/28), six latents ~ normal(0, 1)
    from autogp.jl_b200 import model
    z = [0.3, -1.2, 0.5, 0.1, 0.9, -0.4]
    lin = gp.Linear(*[model.transform_param(f, v) for f, v in zip(("intercept", "bias", "amplitude"), z[:3])])
    per = gp.Periodic(*[model.transform_param(f, v) for f, v in zip(("lengthscale", "period", "amplitude"), z[3:])])
    want = math.log(4 / 28) + 2 * math.log(6 / 28) + sum(-0.5 * v * v - 0.5 * math.log(2 * math.pi) for v in z)
    assert tm.log_prior_tree(lin + per, 1, tm.GPConfig()) == pytest.approx(want, abs=1e-12)
    # a ChangePoint below a Plus has zero mass (Model.jl:101, 110-113)
    bad = gp.Plus(gp.ChangePoint(lin, per, 0.5, 0.001), per)
    assert tm.log_prior_tree(bad, 1, tm.GPConfig()) == -math.inf
    # at max_depth only the leaf distribution is left
    assert tm.log_prior_tree(lin + per, 1, tm.GPConfig(max_depth=1)) == -math.inf


@pytest.mark.parametrize("biased", [False, True])
def test_pick_probabilities_sum_to_one_and_match_the_sampler(biased):
    rng = np.random.default_rng(2)
    tree = None
    while tree is None or gp.size(tree) < 7:
        tree = tm.sample_tree_prior(1, tm.GPConfig(), rng)
    idxs = _all_indices(tree)
    lp = {i: tm.log_pick_random_node(tree, 1, i, biased) for i in idxs}
    assert sum(math.exp(v) for v in lp.values()) == pytest.approx(1.0, abs=1e-12)
    if not biased:   # uniform over nodes (inference_utils.jl:19-23)
        assert all(math.exp(v) == pytest.approx(1.0 / len(idxs), abs=1e-12) for v in lp.values())
    counts = Counter()
    for _ in range(20000):
        _, i, l = tm.pick_random_node(tree, 1, biased, rng)
        counts[i] += 1
        assert l == pytest.approx(lp[i], abs=1e-12)
    for i in idxs:
        p = math.exp(lp[i])
        assert abs(counts[i] / 20000 - p) < 5 * math.sqrt(p * (1 - p) / 20000) + 1e-3
    # noroot: the root is never picked, the rest still sums to one
    lpn = [tm.log_pick_random_node(tree, 1, i, biased, noroot=True) for i in idxs]
    assert lpn[0] == -math.inf
    assert sum(math.exp(v) for v in lpn[1:]) == pytest.approx(1.0, abs=1e-12)


def test_path_probabilities():
    # bounded: all holes down to max_depth sum to one
    total = 0.0
    for hole in range(2, 2 ** 4):
        if (hole >> (hole.bit_length() - 2)) == 2:     # below heap index 2
            total += math.exp(tm.log_generate_random_path(2, hole, 4))
    assert total == pytest.approx(1.0, abs=1e-12)
    assert tm.log_generate_random_path(2, 16, 4) == -math.inf
    rng = np.random.default_rng(3)
    for _ in range(200):
        hole, lp = tm.generate_random_path(3, -1, rng)
        assert lp == pytest.approx(tm.log_generate_random_path(3, hole, -1), abs=1e-12)
        hole, lp = tm.generate_random_path(1, 3, rng, noroot=True)
        assert hole > 1 and lp == pytest.approx(tm.log_generate_random_path(1, hole, 3, noroot=True), abs=1e-12)


def _stats(tree):
    return (min(gp.size(tree), 9), type(tree).__name__)


def _chi2(chain: Counter, prior: Counter, n_chain_eff: float):
    keys = [k for k in set(chain) | set(prior) if prior[k] >= 30]
    n_c, n_p = sum(chain[k] for k in keys), sum(prior[k] for k in keys)
    x = 0.0
    for k in keys:
        pc, pp = chain[k] / n_c, prior[k] / n_p
        var = pp * (1 - pp) * (1.0 / n_p + 1.0 / n_chain_eff)
        x += (pc - pp) ** 2 / var
    return x, len(keys)


@pytest.mark.parametrize("config,biased,move", [
    (tm.GPConfig(), False, "mixture"),
    (tm.GPConfig(), True, "mixture"),
    (tm.GPConfig(max_depth=3), False, "mixture"),
    (tm.GPConfig(max_depth=3), True, "detach_attach"),
    (tm.GPConfig(changepoints=False), False, "detach_attach"),
    (tm.GPConfig(max_depth=4), False, "subtree_replace"),
])
def test_flat_likelihood_chain_leaves_the_prior_invariant(config, biased, move):
    rng = np.random.default_rng(11)
    n_prior = 40000
    prior = Counter(_stats(tm.sample_tree_prior(1, config, rng)) for _ in range(n_prior))
    propose = {"mixture": tm.tree_rejuvenation_proposer(config, biased),
               "detach_attach": lambda nd, r: tm.detach_attach_proposal(nd, r, config, biased),
               "subtree_replace": lambda nd, r: tm.subtree_replace_proposal(nd, r, config, biased)}[move]
    # many short chains started FROM the prior: every state of every chain is then a prior draw iff the kernel is
    # invariant, and chains are independent of each other
    n_chains, n_steps, chain = 4000, 10, Counter()
    accepted = 0
    for _ in range(n_chains):
        t = tm.sample_tree_prior(1, config, rng)
        for _ in range(n_steps):
            new, logr = propose(t, rng)
            if math.log(rng.random()) < logr:
                assert _valid(new, config)
                accepted += new is not t
                t = new
        chain[_stats(t)] += 1
    assert accepted > 0.05 * n_chains * n_steps      # the moves do move
    x, k = _chi2(chain, prior, n_chains)
    # chi-square with k - 1 degrees of freedom: mean k - 1, sd sqrt(2 (k - 1)); 5 sd of slack keeps the test quiet
    assert x < (k - 1) + 5 * math.sqrt(2 * (k - 1)) + 5, (x, k, chain, prior)


def test_a_wrong_ratio_is_caught_by_the_invariance_test():
    """The test above has teeth: dropping the pick-probability ratio of subtree-replace (accepting on the prior alone,
    i.e. log ratio 0) visibly distorts the size distribution."""
    config, rng = tm.GPConfig(), np.random.default_rng(5)
    prior = Counter(_stats(tm.sample_tree_prior(1, config, rng)) for _ in range(40000))
    chain = Counter()
    for _ in range(4000):
        t = tm.sample_tree_prior(1, config, rng)
        for _ in range(10):
            new, logr = tm.subtree_replace_proposal(t, rng, config, False)
            if logr > -math.inf:     # wrong on purpose
                t = new
        chain[_stats(t)] += 1
    x, k = _chi2(chain, prior, 4000)
    assert x > (k - 1) + 5 * math.sqrt(2 * (k - 1)) + 5


def test_detach_then_attach_ratio_is_antisymmetric():
    """For one concrete pair of trees the forward and the reverse log ratios must cancel: detach b out of a, then the
    attach that rebuilds the old tree.  Evaluated through the density functions the proposal uses."""
    config = tm.GPConfig()
    rng = np.random.default_rng(7)
    lin = tm.sample_tree_prior(4, tm.GPConfig(max_depth=3), rng)
    per = tm.sample_tree_prior(5, tm.GPConfig(max_depth=3), rng)
    ge = tm.sample_tree_prior(3, tm.GPConfig(max_depth=2), rng)
    old = gp.Plus(gp.Times(lin, per), ge)        # a = index 2 (the Times), b = index 5 (per)
    new = tm.replace_at(old, 2, per)
    for biased in (False, True):
        on_path = tm._path_dict(2, 5)
        fwd = (math.log(0.5) + tm.log_pick_random_node(old, 1, 2, biased)
               + tm.log_pick_random_node(tm.subtree_at(old, 2), 2, 5, biased))
        bwd = (math.log(0.5) + tm.log_pick_random_node(new, 1, 2, biased)
               + tm.log_generate_random_path(2, 5, -1)
               + tm.log_aux_tree(tm.subtree_at(old, 2), 2, on_path, False, config))
        detach = tm.log_prior_tree(new, 1, config) - tm.log_prior_tree(old, 1, config) + bwd - fwd
        attach = tm.log_prior_tree(old, 1, config) - tm.log_prior_tree(new, 1, config) + fwd - bwd
        assert detach == pytest.approx(-attach, abs=1e-12)
        # the discarded nodes (the Times and lin) appear in the old prior and in the reverse proposal: what is left of
        # the ratio is structural — the Times has mass 5/28 under the prior (below a Plus) and 4/10 among the
        # branch types of the auxiliary proposal, which starts from the GLOBAL config (inference_rejuv_tree_da.jl:157-159)
        want = (-math.log(5 / 28) + math.log(0.4)                                   # Times: prior vs forced-branch dist
                + math.log(0.5 * 0.5 * 0.5)                                         # path: go on, right, stop
                + tm.log_pick_random_node(new, 1, 2, biased) - tm.log_pick_random_node(old, 1, 2, biased)
                - tm.log_pick_random_node(tm.subtree_at(old, 2), 2, 5, biased))
        assert detach == pytest.approx(want, abs=1e-9)


# ------------------------------------------------------------------------------------------------
# run_smc_anneal_data: the whole loop of src/inference_smc_anneal_data.jl:143-273 around the batched scoring call
# ------------------------------------------------------------------------------------------------
def _series(n, seed=4):
    rng = np.random.default_rng(seed)
    ts = rng.permutation(np.arange(n) / (n - 1))
    xs = 0.8 * np.sin(2 * np.pi * ts / 0.25) + 0.5 * ts + 0.05 * rng.standard_normal(n)
    return ts, xs


def test_run_smc_anneal_data_on_the_oracle_stand_in():
    """Host logic of the SMC loop with the oracle standing in for the GPU engine (CPU, small)."""
    from autogp.jl_b200 import smc
    from helpers import OracleEngineWithNoiseCall, from_agp
    import autogp_oracle as o

    ts, xs = _series(24)
    seen = []
    state = smc.run_smc_anneal_data(ts, xs, config=tm.GPConfig(max_depth=2), n_particles=4, n_mcmc=2, n_hmc=1,
                                    schedule=smc.linear_schedule(24, 0.5), seed=3, engine=OracleEngineWithNoiseCall(),
                                    hmc_config={"L_param": 2, "L_noise": 2},
                                    callback_fn=lambda **kw: seen.append((kw["step"], kw["rejuvenated"], kw["resampled"])))
    assert [s[0] for s in seen] == [0, 12, 24] and seen[1][1] and seen[2][1] and not seen[2][2]
    assert len(state.nodes) == 4 and np.all(np.isfinite(state.scores)) and np.all(np.isfinite(state.log_weights))
    for nd, nz, sc in zip(state.nodes, state.noises, state.scores):     # the stored scores are those of the final state
        assert _valid(nd, tm.GPConfig(max_depth=2))
        assert sc == pytest.approx(o.log_marginal_likelihood(from_agp(nd), nz, ts, xs), rel=1e-9)
    # same seed, same run; a fixed noise is carried through and weighs every particle by the latent's prior density
    again = smc.run_smc_anneal_data(ts, xs, config=tm.GPConfig(max_depth=2), n_particles=4, n_mcmc=2, n_hmc=1,
                                    schedule=smc.linear_schedule(24, 0.5), seed=3, engine=OracleEngineWithNoiseCall(),
                                    hmc_config={"L_param": 2, "L_noise": 2})
    assert [repr(a) for a in again.nodes] == [repr(a) for a in state.nodes] and again.noises == state.noises
    fixed = smc.initialize_particles(3, tm.GPConfig(noise=0.1), seed=1)
    assert fixed.noises == [0.1] * 3 and np.all(fixed.log_weights < 0) and len(set(fixed.log_weights)) == 1
    with pytest.raises(ValueError):
        smc.run_smc_anneal_data(ts, xs, schedule=[5, 5, 24], engine=OracleEngineWithNoiseCall())


@pytest.mark.gpu
def test_gpu_structure_learning_finds_the_periodic_structure():
    """The whole reference loop through the C-ABI: 32 particles, data-annealing schedule, the reference's own
    proposals.  On a clean periodic-plus-trend series the posterior must put its weight on kernels with a Periodic
    component, score far above the prior's particles, and the stored scores must be the engine's scores of the final
    particles (state consistency)."""
    import autogp.jl_b200 as agp
    from autogp.jl_b200 import smc

    n, P = 160, 32
    ts, xs = _series(n)
    eng = agp.Engine(0)
    cfg = tm.GPConfig(max_depth=3)
    prior = smc.initialize_particles(P, cfg, seed=21)
    prior_lml, prior_info = eng.lml_batch(prior.nodes, prior.noises, ts, xs)
    state = smc.run_smc_anneal_data(ts, xs, config=cfg, n_particles=P, n_mcmc=12, n_hmc=4,
                                    schedule=smc.linear_schedule(n, 0.25), seed=21, engine=eng)
    lml, info = eng.lml_batch(state.nodes, state.noises, ts, xs)
    assert np.all(info == 0)
    np.testing.assert_allclose(state.scores, lml, rtol=1e-9)
    w = smc.compute_particle_weights(state.log_weights)
    assert w.sum() == pytest.approx(1.0)
    has_per = np.array([any(isinstance(a, agp.Periodic) for a in agp.unroll(nd)) for nd in state.nodes])
    assert w[has_per].sum() > 0.8
    assert np.median(lml) > np.nanmedian(np.where(prior_info == 0, prior_lml, np.nan)) + 20
    assert math.isfinite(state.log_ml_est)


def test_move_edge_cases():
    rng = np.random.default_rng(0)
    # max_depth = 1: only leaves exist; detach-attach is impossible (the reference errors) and the mixture never picks it
    cfg1 = tm.GPConfig(max_depth=1)
    leaf = tm.sample_tree_prior(1, cfg1, rng)
    assert isinstance(leaf, gp.LeafNode)
    with pytest.raises(ValueError):
        tm.detach_attach_proposal(leaf, rng, cfg1, False)
    propose = tm.tree_rejuvenation_proposer(cfg1, False)
    for _ in range(50):
        new, logr = propose(leaf, rng)
        assert isinstance(new, gp.LeafNode) and logr == pytest.approx(0.0, abs=1e-12)   # leaf for leaf from the prior: ratio 1
    # a tree the prior cannot generate (ChangePoint below Plus) can be left but never entered
    lin, per = gp.Linear(0.3, 0.2, 0.1), gp.Periodic(0.5, 0.3, 0.2)
    bad = gp.Plus(gp.ChangePoint(lin, per, 0.5, tm.CHANGEPOINT_SCALE), per)
    assert tm.log_prior_tree(bad, 1, tm.GPConfig()) == -math.inf
    # no ChangePoints in the configuration: none is ever proposed
    cfg = tm.GPConfig(changepoints=False)
    t = tm.sample_tree_prior(1, cfg, rng)
    for _ in range(300):
        new, logr = tm.tree_rejuvenation_proposer(cfg, True)(t, rng)
        assert not any(isinstance(a, gp.ChangePoint) for a in gp.unroll(new))
        if math.log(rng.random()) < logr:
            t = new
    # heap-index helpers
    tree = gp.Plus(gp.Times(lin, per), per)
    assert tm.subtree_at(tree, 5) is per and tm.subtree_at(tree, 2).left is lin
    assert repr(tm.replace_at(tree, 4, per)) == repr(gp.Plus(gp.Times(per, per), per))
    with pytest.raises(ValueError):
        tm._path_bits(2, 7)


def test_run_smc_anneal_data_with_fixed_noise():
    """config.noise fixes the observation noise (src/inference_smc_anneal_data.jl:183-187): no noise moves, every
    particle keeps the value, and the constrained latent's prior density enters every initial log-weight."""
    from autogp.jl_b200 import smc
    from helpers import OracleEngineWithNoiseCall

    ts, xs = _series(20)
    eng = OracleEngineWithNoiseCall()
    state = smc.run_smc_anneal_data(ts, xs, config=tm.GPConfig(max_depth=2, noise=0.05), n_particles=3, n_mcmc=2, n_hmc=1,
                                    schedule=[10, 20], seed=2, engine=eng, hmc_config={"L_param": 2, "L_noise": 2})
    assert state.noises == [0.05] * 3
    assert not any(kind == "noise" for kind, _ in eng.batches)
    assert np.all(np.isfinite(state.log_weights)) and np.all(np.isfinite(state.scores))


def test_single_process_smc_does_not_import_torch():
    """A fit without a process group must not pay the torch import (about 2 s): smc.* consult torch.distributed only when
    torch is already in the process."""
    import subprocess
    import sys
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path[:0] = [%r, %r, %r]\n"
        "import numpy as np\n"
        "from autogp.jl_b200 import smc, tree_moves as tm\n"
        "from helpers import OracleEngineWithNoiseCall\n"
        "ts = np.linspace(0, 1, 12); xs = np.sin(7 * ts)\n"
        "st = smc.run_smc_anneal_data(ts, xs, config=tm.GPConfig(max_depth=2), n_particles=2, n_mcmc=1, n_hmc=1, schedule=[6, 12], seed=1,\n"
        "                             engine=OracleEngineWithNoiseCall(), hmc_config={'L_param': 1, 'L_noise': 1})\n"
        "assert len(st.nodes) == 2\n"
        "assert 'torch' not in sys.modules, 'torch was imported'\n"
        "print('ok')\n") % (root, os.path.join(root, "oracle"), os.path.join(root, "tests"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
