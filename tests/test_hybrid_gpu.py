"""GPU suite of the hybrid factorisation (csrc/agp_ozaki.cu): the long contractions of the blocked Cholesky as exact
int8 digit-plane products on tcgen05 (kind::i8, TMEM accumulators), the rest on the FP64 persistent kernel.

Stated tolerances.  The scheme rounds every entry of L once at 2^-55 of a per-row power-of-two bound of sqrt(K_ii) (seven
balanced base-256 digits) and drops digit products below 2^-62 of the two row scales, so the error of a contraction is
ABSOLUTE in units of sqrt(K_ii K_kk) (about sqrt(depth) 2^-57), where FP64 accumulation rounds relative to the running sum.
Observed on these cases: factor entries within 1e-14 of the FP64 schedule's (relative to max |L|), LML within 2e-15 of it
(n <= 2300, these trees), and over the 64 benchmark particles 1.7e-13 at n = 2048, 1.2e-13 at n = 4096, 6.4e-12 at n = 8192
(tools/hybrid_check.py) — against the north_star bound 1e-8.  Asserted here: 1e-11 against the FP64 schedule and the oracle
up to n = 2300, 1e-9 at n = 8192.
"""
import numpy as np
import pytest

import autogp_oracle as o
import c_oracle
import helpers as H

pytestmark = pytest.mark.gpu

TREES = ["se*per+lin", "ge+per*lin", "se+wn", "cp(lin,se)", "se*per+lin"]


def _batch(n, P):
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p, TREES[p % len(TREES)]) for p in range(P)]
    return ts, xs, parts, [H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]


@pytest.fixture()
def eng():
    import autogp.jl_b200 as agp

    e = agp.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("n,width", [(300, 1), (640, 2), (1100, 4), (1100, 3), (1153, 2), (2048, 4), (2300, 5)])
def test_hybrid_matches_the_fp64_schedule_and_the_oracle(eng, n, width):
    """Forced on (mode 1) at ragged sizes, odd numbers of tile rows (the second CTA of a pair idles on the last row),
    every width: LML and factor against the single-launch FP64 schedule and the oracle."""
    ts, xs, parts, nodes, noises = _batch(n, 5)
    eng.set_hybrid(0)
    lml0, info0 = eng.lml_batch(nodes, noises, ts, xs)
    assert not eng.hybrid_info()[0]
    L0 = eng.factor(1)
    eng.set_hybrid(1, width, 2)
    lml1, info1 = eng.lml_batch(nodes, noises, ts, xs)
    assert eng.hybrid_info()[0] and eng.hybrid_info()[1] == width
    L1 = eng.factor(1)
    assert np.all(info0 == 0) and np.all(info1 == 0)
    assert np.max(np.abs(lml1 - lml0) / np.abs(lml0)) <= 1e-11
    ref = np.array([o.log_marginal_likelihood(nd, nz, ts, xs) for nd, nz in parts[:2]])
    assert np.max(np.abs(lml1[:2] - ref) / np.abs(ref)) <= 1e-11
    assert np.max(np.abs(L1 - L0)) <= 1e-12 * np.max(np.abs(L0))
    # deterministic: integer sums are exact, the FP64 items run the same arithmetic per tile
    again, _ = eng.lml_batch(nodes, noises, ts, xs)
    assert np.array_equal(again, lml1)


def test_hybrid_is_chosen_by_size_and_can_be_switched_off(eng):
    """Default: plain LML runs from 12 block columns on (n >= 1409) — one block column later per 8 particles below 48 —, the
    gradient calls from 8 (n >= 897); continuations (set_prefix + run_append) and predictive batches never."""
    ts, xs, parts, nodes, noises = _batch(2048, 48)
    eng.upload(nodes[:3], noises[:3], ts, xs)       # 3 particles: the switch sits at 18 block columns
    assert not eng.hybrid_info()[0]
    eng.upload(nodes[:20], noises[:20], ts, xs)     # 20 particles: at 16
    assert eng.hybrid_info()[0]
    eng.set_prefix(1900)
    assert not eng.hybrid_info()[0]
    eng.upload(nodes, noises, ts, xs)
    assert eng.hybrid_info()[0]
    eng.run()
    lml_h, info = eng.fetch()
    assert np.all(info == 0)
    eng.set_prefix(1400)
    assert not eng.hybrid_info()[0]
    eng.set_prefix(2048)
    eng.set_hybrid(0)
    assert not eng.hybrid_info()[0]
    eng.run()
    lml_f, _ = eng.fetch()
    assert np.max(np.abs(lml_h - lml_f) / np.abs(lml_f)) <= 1e-10
    ts_g, xs_g = o.synthetic_series(1536)
    eng.set_hybrid(-1)
    _, _, _, ginfo = eng.lml_grad_batch(nodes[:2], noises[:2], ts_g, xs_g)
    assert eng.hybrid_info()[0] and np.all(ginfo == 0)
    eng.lml_grad_batch(nodes[:2], noises[:2], ts_g[:1000], xs_g[:1000])
    assert eng.hybrid_info()[0]
    eng.lml_grad_batch(nodes[:2], noises[:2], ts_g[:890], xs_g[:890])
    assert not eng.hybrid_info()[0]
    eng.upload(nodes, noises, ts, xs)
    # a data-annealing continuation of a hybrid factor (agp_lml_run_append) is the FP64 schedule on top of it
    eng.set_hybrid(-1)
    eng.set_prefix(1930)
    eng.run()
    got, info = eng.fetch()
    assert np.all(info == 0) and eng.hybrid_info()[0]
    eng.set_prefix(2048)
    eng.run_append()
    app, info = eng.fetch()
    assert np.all(info == 0)
    assert np.max(np.abs(app - lml_f) / np.abs(lml_f)) <= 1e-10


def test_hybrid_reports_a_failed_factorisation_like_the_fp64_schedule(eng):
    """A matrix that is not positive definite: the LAPACK info of the FP64 schedule, NaN score, the other particles of
    the batch unharmed (digit planes of NaN / huge entries are garbage bytes, never a hang)."""
    n = 700
    ts, xs = o.synthetic_series(n)
    tsd = ts.copy()
    tsd[500] = tsd[10]          # duplicate time point + zero noise: singular leading minor in the fourth block column
    cases = [(o.SquaredExponential(0.5, 1.0), 0.0), (o.SquaredExponential(0.1, 1.0), 0.1), (o.Constant(1.0), -2.0)]
    nodes, noises = [H.to_agp(k) for k, _ in cases], [nz for _, nz in cases]
    eng.set_hybrid(0)
    lml0, info0 = eng.lml_batch(nodes, noises, tsd, xs)
    eng.set_hybrid(1, 2, 2)
    lml1, info1 = eng.lml_batch(nodes, noises, tsd, xs)
    assert eng.hybrid_info()[0]
    assert info1[1] == 0 and info1[2] == info0[2] == 1 and info0[0] != 0
    # the singular minor's pivot is pure rounding noise: the hybrid schedule may or may not see it as non-positive
    assert np.isnan(lml1[2]) and (np.isnan(lml1[0]) if info1[0] != 0 else np.isfinite(lml1[0]))
    assert abs(lml1[1] - lml0[1]) <= 1e-10 * abs(lml0[1])
    ref_info = c_oracle.lml(o.encode_program(cases[0][0]), tsd, xs, 0.0)[1]
    assert ref_info != 0


@pytest.mark.parametrize("n,width", [(300, 1), (700, 2), (1100, 3), (1153, 4), (2048, 4)])
def test_hybrid_gradient_calls_match_the_fp64_schedule(eng, n, width):
    """agp_lml_grad_batch / agp_lml_grad_noise_batch on the hybrid schedule (the identity-augmented matrix: factorisation,
    trtri and lauum contractions on the int8 path; bound of the appended rows 1 / sqrt(noise)) against the FP64 schedule:
    LML, dLML/dparams, dLML/dnoise.  The gradients are sums of n^2 products of -K^{-1} entries with kernel derivatives:
    agreement to 1e-8 of the gradient's scale."""
    ts, xs, parts, nodes, noises = _batch(n, 4)
    eng.set_hybrid(0)
    lml0, g0, gn0, info0 = eng.lml_grad_batch(nodes, noises, ts, xs)
    lmln0, gnn0, infon0 = eng.lml_grad_noise_batch(nodes, noises, ts, xs)
    eng.set_hybrid(1, width, 2)
    lml1, g1, gn1, info1 = eng.lml_grad_batch(nodes, noises, ts, xs)
    lmln1, gnn1, infon1 = eng.lml_grad_noise_batch(nodes, noises, ts, xs)
    assert np.all(info0 == 0) and np.all(info1 == 0) and np.all(infon1 == 0)
    assert np.max(np.abs(lml1 - lml0) / np.abs(lml0)) <= 1e-10 and np.max(np.abs(lmln1 - lml0) / np.abs(lml0)) <= 1e-10
    for p in range(len(nodes)):
        scale = max(1.0, float(np.max(np.abs(g0[p]))) if len(g0[p]) else 1.0, abs(gn0[p]))
        assert np.max(np.abs(np.asarray(g1[p]) - np.asarray(g0[p]))) <= 1e-8 * scale, (p, g1[p], g0[p])
        assert abs(gn1[p] - gn0[p]) <= 1e-8 * scale and abs(gnn1[p] - gnn0[p]) <= 1e-8 * scale, (p, gn1[p], gn0[p], gnn1[p])
    again = eng.lml_grad_batch(nodes, noises, ts, xs)
    assert all(np.array_equal(a, b) for a, b in zip(g1, again[1])) and np.array_equal(gn1, again[2])


def test_hybrid_gradient_keeps_the_fp64_schedule_when_the_noise_bound_is_useless(eng):
    """The appended rows' scale is the a-priori bound 1 / sqrt(noise); with a noise far below what the kernel itself puts on the
    diagonal (a WhiteNoise node, noise ~ 0) that bound would cost the digits: such a batch stays on the FP64 schedule."""
    n = 700
    ts, xs = o.synthetic_series(n)
    nd = H.to_agp(o.synthetic_particle(0, "se+wn")[0])
    eng.set_hybrid(0)
    ref = eng.lml_grad_batch([nd], [1e-12], ts, xs)
    eng.set_hybrid(1, 2, 2)
    got = eng.lml_grad_batch([nd], [1e-12], ts, xs)
    assert not eng.hybrid_info()[0] and got[3][0] == 0
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1][0], ref[1][0]) and np.array_equal(got[2], ref[2])
    eng.lml_grad_batch([nd], [1e-3], ts, xs)
    assert eng.hybrid_info()[0]
    # the plain LML of the same batch has no such restriction (its row scales come from the Gram diagonal)
    eng.lml_batch([nd], [1e-12], ts, xs)
    assert eng.hybrid_info()[0]


def test_hybrid_full_size_n8192(eng):
    """configs[2] shape with the default settings (hybrid by size): against the oracle at the north_star tolerance."""
    n = 8192
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(3), o.synthetic_particle(7)]
    got, info = eng.lml_batch([H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs)
    assert eng.hybrid_info()[0] and np.all(info == 0)
    ref = o.log_marginal_likelihood(*parts[1], ts, xs)
    assert abs(got[1] - ref) <= 1e-9 * abs(ref)
