"""GPU suite: the CUDA path, called through the C-ABI (ctypes), against the CPU oracle and the
committed golden fixtures.

Stated tolerances
  * Gram entries: |K_gpu - K_oracle| <= 64 eps |K|.  +,* are exact (no FMA contraction); the
    slack is the libm difference (CUDA libdevice exp/sin/pow/tanh, each <= 2 ulp, vs glibc/NumPy)
    amplified by the exponent's condition number (|coef * sin^2| up to ~10 for the fixtures).
  * LML: relative error <= 1e-8 (BASELINE.json north_star); observed ~1e-15 on these cases, and
    a tighter 1e-11 regression bound is asserted where the problem is well conditioned.
"""
import threading

import numpy as np
import pytest

import autogp_oracle as o
import c_oracle
import helpers as H

pytestmark = pytest.mark.gpu

EPS = np.finfo(np.float64).eps
GRAM_RTOL = 64 * EPS
LML_RTOL = 1e-8
LML_RTOL_TIGHT = 1e-11


def oracle_lmls(parts, ts, xs):
    return np.array([o.log_marginal_likelihood(nd, nz, ts, xs) for nd, nz in parts])


def gpu_lmls(engine, parts, ts, xs):
    return engine.lml_batch([H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs)


# ---- site 1: Gram ---------------------------------------------------------------------------

def test_gram_fixture_sweep_vectorised_and_scalar(engine):
    """The 6 base + 108 composite kernels of test/test_GP.jl:24-33,54-56 on its 100-point grid."""
    import autogp.jl_b200 as agp

    _, ds = H.fixture_grid()
    for k in H.fixture_kernels():
        Kv = engine.gram(H.to_agp(k), 0.25, ds, agp.gp.FORM_VECTORIZED)
        Ks = engine.gram(H.to_agp(k), 0.25, ds, agp.gp.FORM_SCALAR)
        Rv = o.compute_cov_matrix_vectorized(k, 0.25, ds)
        Rs = c_oracle.gram(o.encode_program(k), ds, 0.25, form=1)
        assert np.all(np.abs(Kv - Rv) <= GRAM_RTOL * np.abs(Rv)), k
        assert np.all(np.abs(Ks - Rs) <= GRAM_RTOL * np.abs(Rs)), k
        assert np.array_equal(Kv, Kv.T)  # both triangles, Matrix(Symmetric(K))
        assert Kv.flags["F_CONTIGUOUS"]


def test_gram_against_committed_golden(engine):
    ts, grams, _ = H.golden()
    for k, G in zip(H.fixture_kernels(), grams):
        K = engine.gram(H.to_agp(k), 0.0, ts)
        assert np.all(np.abs(K - G) <= GRAM_RTOL * np.abs(G)), k


def test_gram_exact_for_transcendental_free_kernels(engine):
    """Constant / Linear / WhiteNoise / Plus / Times involve only +,*: bit-exact."""
    rng = np.random.default_rng(0)
    ts = rng.uniform(0, 1, 77)
    ts[5] = ts[40]  # duplicate time point: WhiteNoise equality test (GP.jl:139)
    k = o.Plus(o.Times(o.Linear(0.3, 1.1, 0.9), o.Constant(0.7)), o.Plus(o.WhiteNoise(0.4), o.Linear(0.6, 0.2, 1.7)))
    K = engine.gram(H.to_agp(k), 0.125, ts)
    R = o.compute_cov_matrix_vectorized(k, 0.125, ts)
    assert np.array_equal(K, R)
    assert K[5, 40] == R[5, 40] and K[5, 40] != K[5, 41]


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 130, 257])
def test_gram_ragged_sizes(engine, n):
    ts = np.random.default_rng(n).uniform(0, 1, n)
    k, nz = o.synthetic_particle(n % 5, "ge+per*lin")
    K = engine.gram(H.to_agp(k), nz, ts)
    R = o.compute_cov_matrix_vectorized(k, nz, ts)
    assert K.shape == (n, n)
    assert np.all(np.abs(K - R) <= GRAM_RTOL * np.abs(R))


def test_gram_changepoint_saturates_exactly(engine):
    """Model.jl:121 hard-codes scale = .001: tanh saturates to +-1, sigma to exactly 0 / 1."""
    ts = np.linspace(0, 1, 64)
    k = o.ChangePoint(o.Linear(0.5), o.SquaredExponential(0.2, 1.3), 0.4, 0.001)
    K = engine.gram(H.to_agp(k), 0.0, ts)
    R = o.compute_cov_matrix_vectorized(k, 0.0, ts)
    # next to the change point (1 - sigma) cancels catastrophically, so a 1-ulp tanh difference is
    # an ABSOLUTE error of a few eps * max|K| there; away from it the entries agree relatively
    assert np.all(np.abs(K - R) <= GRAM_RTOL * np.abs(R) + 8 * EPS * np.max(np.abs(R)))
    assert np.all(K[:20, 40:] == 0.0)  # opposite sides of the change point are independent
    assert np.array_equal(K[:20, :20], R[:20, :20])  # sigma == 1 exactly: pure Linear block, bit-exact


def test_gram_extreme_arguments_take_the_fallback_paths(engine):
    """The interpreter's own exp / sin^2 / constant division (agp_math.cuh) hand arguments outside their fast
    ranges to libdevice: exp arguments below -700 (tiny lengthscales: results underflow through the denormals to
    0), sin arguments above 1e5 (tiny periods), divisors near the ends of the exponent range, zero numerators
    (the diagonal) — all inside one warp next to ordinary entries."""
    ts = np.linspace(0.0, 1.0, 96)
    cases = [
        o.SquaredExponential(0.0262, 1.7),                  # -0.5 dx^2 / l^2 spans 0 .. -728: fast path, fallback, denormals
        o.Periodic(0.9, 1.7e-5, 0.8),                       # (pi / p) |dx| up to 1.8e5: Payne-Hanek fallback of sin
        o.Periodic(0.05, 0.31, 1.2),                        # -2/l^2 = -800: exp arguments down to -800
        o.SquaredExponential(1e-160, 1.0),                  # l^2 = 1e-320 (denormal divisor): generic division
        o.SquaredExponential(3e155, 1.0),                   # l^2 overflows to inf: every off-diagonal argument is -0
        o.GammaExponential(1e-3, 1.9, 1.1),                 # (|dx| / l)^gamma up to 5e5: exp fallback after pow
        o.Plus(o.Times(o.SquaredExponential(0.03, 1.0), o.Periodic(0.07, 2e-5, 1.0)), o.Linear(0.2, 0.1, 0.4)),
    ]
    for k in cases:
        with np.errstate(all="ignore"):
            R = o.compute_cov_matrix_vectorized(k, 0.0, ts)
        K = engine.gram(H.to_agp(k), 0.0, ts)
        assert np.all(np.isfinite(K)) == np.all(np.isfinite(R)), k
        big = np.abs(R) > 1e-290          # denormal results: libm implementations differ in the last denormal bits
        # exponent arguments here reach several hundred, so an ulp of argument error is hundreds of ulps of the result:
        # the tolerance scales with the magnitude of log K instead of the fixed 64 eps of the fixture sweep
        cond = np.maximum(1.0, np.abs(np.log(np.maximum(np.abs(R), 1e-300))))
        assert np.all(np.abs(K - R)[big] <= (16 * EPS * cond * np.abs(R))[big]), k
        assert np.all(np.abs(K[~big]) <= 1e-289), k
        assert np.array_equal(K, K.T)


def test_gram_does_not_mutate_inputs_and_empty(engine):
    ts = np.linspace(0, 1, 10)
    ts0 = ts.copy()
    engine.gram(H.to_agp(o.SquaredExponential(0.3)), 0.1, ts)
    assert np.array_equal(ts, ts0)
    assert engine.gram(H.to_agp(o.SquaredExponential(0.3)), 0.1, np.zeros(0)).shape == (0, 0)


# ---- site 2: LML ------------------------------------------------------------------------------

def test_lml_golden_fixture_batch_ragged_programs(engine):
    """All 114 fixture kernels as ONE ragged batch against the mpmath 50-digit golden values."""
    ts, _, lml = H.golden()
    xs = H.fixture_xs(ts)
    ks = H.fixture_kernels()
    got, info = engine.lml_batch([H.to_agp(k) for k in ks], [lml["noise"]] * len(ks), ts, xs)
    assert np.all(info == 0)
    truth = np.array(lml["fixture_mp"])
    assert np.all(np.abs(got - truth) <= 1e-12 * np.maximum(1.0, np.abs(truth)))


def test_lml_config0_n128_se_whitenoise(engine):
    """BASELINE.json configs[0]: n=128, 4 particles, SE + WhiteNoise."""
    _, _, lml = H.golden()
    ts, xs = o.synthetic_series(128)
    parts = [o.synthetic_particle(p, "se+wn") for p in range(4)]
    got, info = gpu_lmls(engine, parts, ts, xs)
    assert np.all(info == 0)
    assert H.rel_err(got, lml["config0_f64"]) <= LML_RTOL_TIGHT


@pytest.mark.parametrize("n", [1, 2, 3, 31, 127, 128, 129, 255, 256, 257, 300, 513, 640])
def test_lml_ragged_sizes_mixed_trees(engine, n):
    ts, xs = o.synthetic_series(n)
    trees = ["se*per+lin", "se+wn", "ge+per*lin", "cp(lin,se)"]
    parts = [o.synthetic_particle(p, trees[p % 4]) for p in range(7)]
    got, info = gpu_lmls(engine, parts, ts, xs)
    ref = oracle_lmls(parts, ts, xs)
    assert np.all(info == 0)
    assert H.rel_err(got, ref) <= LML_RTOL_TIGHT


def test_lml_hmc_benchmarks_and_predictive_identity(engine):
    """test/experiment_hmc.jl:111-132, 180-184: LML(joint) - LML(obs) = predictive logpdf."""
    ts = np.linspace(0.0, 10.0, 1000)
    rng = np.random.default_rng(7)
    idx = rng.permutation(1000)
    obs, test = idx[:200], idx[200:260]  # unsorted, like the shuffled reference path
    both = np.concatenate([obs, test])
    xs = H.fixture_xs(ts / 10.0) + 0.05 * rng.standard_normal(1000)
    parts = [(k, nz + o.JITTER) for k, nz in H.hmc_benchmarks()]
    l_obs, i1 = gpu_lmls(engine, parts, ts[obs], xs[obs])
    l_joint, i2 = gpu_lmls(engine, parts, ts[both], xs[both])
    assert np.all(i1 == 0) and np.all(i2 == 0)
    assert H.rel_err(l_obs, oracle_lmls(parts, ts[obs], xs[obs])) <= LML_RTOL
    assert H.rel_err(l_joint, oracle_lmls(parts, ts[both], xs[both])) <= LML_RTOL
    for (k, nz), a, b in zip(parts, l_joint, l_obs):
        mu, cov = o.predictive_mvn(k, nz, ts[obs], xs[obs], ts[test])
        assert o.mvn_logpdf(xs[test], mu, cov) == pytest.approx(a - b, rel=2e-7)


def test_lml_not_positive_definite_reports_lapack_info(engine):
    import autogp.jl_b200 as agp

    ts, xs = o.synthetic_series(200)
    cases = [(o.Constant(1.0), -2.0),                 # first pivot negative -> info 1
             (o.SquaredExponential(0.1, 1.0), 0.1),   # fine
             (o.Constant(1.0), 0.0)]                  # rank one: second pivot is exactly 0 -> info 2
    got, info = gpu_lmls(engine, cases, ts, xs)
    want = [c_oracle.lml(o.encode_program(k), ts, xs, nz)[1] for k, nz in cases]
    assert info.tolist() == want
    assert info[1] == 0 and np.isfinite(got[1])
    assert np.isnan(got[0]) and np.isnan(got[2])
    with pytest.raises(agp.PosDefException) as e:
        agp.log_marginal_likelihoods([H.to_agp(k) for k, _ in cases], [nz for _, nz in cases], ts, xs, engine=engine)
    assert e.value.info == 1 and e.value.particle == 0
    # failure deep inside the matrix (second block column) still reports the global 1-based index
    n = 300
    ts, xs = o.synthetic_series(n)
    tsd = ts.copy()
    tsd[200] = tsd[10]  # duplicate time point + zero noise => singular leading minor 201
    got, info = gpu_lmls(engine, [(o.SquaredExponential(0.5, 1.0), 0.0)], tsd, xs)
    ref_info = c_oracle.lml(o.encode_program(o.SquaredExponential(0.5, 1.0)), tsd, xs, 0.0)[1]
    assert info[0] != 0 and ref_info != 0


def test_lml_empty_inputs(engine):
    lml, info = engine.lml_batch([H.to_agp(o.Constant(1.0))] * 3, [0.1] * 3, np.zeros(0), np.zeros(0))
    assert lml.tolist() == [0.0, 0.0, 0.0] and info.tolist() == [0, 0, 0]  # empty mvnormal scores 0
    lml, info = engine.lml_batch([], [], np.linspace(0, 1, 5), np.zeros(5))
    assert lml.shape == (0,) and info.shape == (0,)


def test_lml_data_annealing_prefixes(engine):
    """inference_smc_anneal_data.jl:212-217: the same particles re-scored on growing prefixes."""
    n = 410
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p) for p in range(5)]
    engine.upload([H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs)
    for step in (41, 128, 205, 300, 410, 129):
        engine.set_prefix(step)
        engine.run()
        got, info = engine.fetch()
        assert np.all(info == 0)
        assert H.rel_err(got, oracle_lmls(parts, ts[:step], xs[:step])) <= LML_RTOL_TIGHT


def test_lml_deep_and_bushy_trees(engine):
    ts, xs = o.synthetic_series(150)
    leaves = [o.SquaredExponential(0.2 + 0.05 * i, 0.5 + 0.1 * i) for i in range(6)] + \
             [o.Periodic(0.5 + 0.1 * i, 0.2 + 0.05 * i, 0.7) for i in range(6)] + [o.Linear(0.1 * i, 0.3, 0.4) for i in range(4)]

    def right_chain(xs_, op):
        t = xs_[-1]
        for leaf in reversed(xs_[:-1]):
            t = op(leaf, t)
        return t

    def balanced(xs_, d=0):
        if len(xs_) == 1:
            return xs_[0]
        h = len(xs_) // 2
        op = o.Plus if d % 2 == 0 else o.Times
        return op(balanced(xs_[:h], d + 1), balanced(xs_[h:], d + 1))

    big = [o.SquaredExponential(0.1 + 0.01 * i, 0.05) for i in range(40)]  # 79 instructions: global-memory program path
    cp_nest = o.ChangePoint(o.ChangePoint(leaves[0], leaves[7], 0.3, 0.05), o.Plus(leaves[12], o.ChangePoint(leaves[1], leaves[8], 0.7, 0.1)), 0.5, 0.02)
    trees = [right_chain(leaves[:12], o.Plus), balanced(leaves), balanced(big), cp_nest,
             o.ChangePoint(leaves[12], balanced(leaves[:8]), 0.4, 0.01)]  # right operand deeper: swapped CP opcode
    parts = [(t, 0.05 + 0.01 * i) for i, t in enumerate(trees)]
    got, info = gpu_lmls(engine, parts, ts, xs)
    assert np.all(info == 0)
    assert H.rel_err(got, oracle_lmls(parts, ts, xs)) <= LML_RTOL_TIGHT
    K = engine.gram(H.to_agp(trees[2]), 0.1, ts)
    R = o.compute_cov_matrix_vectorized(trees[2], 0.1, ts)
    assert np.all(np.abs(K - R) <= GRAM_RTOL * np.abs(R))


def test_program_errors_are_reported_not_thrown(engine):
    from autogp.jl_b200 import _lib
    import ctypes as C

    lib = _lib.load()
    h = engine._h
    ts = np.linspace(0, 1, 8)
    K = np.empty((8, 8), order="F")

    def gram(ops, offs, params):
        ops = np.asarray(ops, np.int32)
        offs = np.asarray(offs, np.int32)
        params = np.asarray(params, np.float64)
        f = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        return lib.agp_gram(h, f(ops, C.c_int32), f(offs, C.c_int32), len(ops), f(params, C.c_double), len(params),
                            f(ts, C.c_double), 8, 0.1, 0, f(K, C.c_double))

    assert gram([42], [0], [1.0]) == _lib.AGP_ERR_PROGRAM and b"unknown node-type" in lib.agp_last_error(h)
    assert gram([1, 6], [0, 1], [1.0]) == _lib.AGP_ERR_PROGRAM and b"underflow" in lib.agp_last_error(h)
    assert gram([1, 1], [0, 1], [1.0, 2.0]) == _lib.AGP_ERR_PROGRAM  # two trees left on the stack
    assert gram([4], [0], [1.0, 2.5, 1.0]) == _lib.AGP_ERR_PROGRAM and b"gamma" in lib.agp_last_error(h)
    assert gram([2], [0], [1.0]) == _lib.AGP_ERR_PROGRAM  # Linear needs 3 params
    assert gram([1], [0], [1.0]) == 0  # handle still usable afterwards
    assert lib.agp_lml_run(h) in (0, _lib.AGP_ERR_STATE)


def test_lml_reentrant_engines_from_threads(engine):
    """up to nthreads concurrent callers, one handle each."""
    import autogp.jl_b200 as agp

    ts, xs = o.synthetic_series(260)
    parts = [o.synthetic_particle(p, "ge+per*lin") for p in range(6)]
    ref = oracle_lmls(parts, ts, xs)
    out = {}

    def work(i):
        eng = agp.Engine(0)
        for _ in range(3):
            out[i] = eng.lml_batch([H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs)[0]
        eng.close()

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(4):
        assert H.rel_err(out[i], ref) <= LML_RTOL_TIGHT


def test_smc_step_single_gpu(engine):
    from autogp.jl_b200 import smc

    ts, xs = o.synthetic_series(96)
    parts = [o.synthetic_particle(p, "se+wn") for p in range(8)]
    state = smc.ParticleState(nodes=[H.to_agp(nd) for nd, _ in parts], noises=[nz for _, nz in parts])
    prev = np.zeros(8)
    for step in (32, 64, 96):
        scores = smc.smc_step(state, ts[:step], xs[:step], engine=engine)
        ref = oracle_lmls(parts, ts[:step], xs[:step])
        assert H.rel_err(scores, ref) <= LML_RTOL_TIGHT
        prev = ref
    assert np.allclose(state.log_weights, prev, rtol=1e-10)  # sum of increments telescopes to the last score


def test_factor_reproduces_the_gram_matrix(engine):
    """agp_lml_copy_factor: L L' = K + noise I to rounding, and L' is the upper factor LAPACK dpotrf('U') returns."""
    import scipy.linalg as sla

    n = 300
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p, "se*per+lin") for p in range(3)]
    gpu_lmls(engine, parts, ts, xs)
    for p, (nd, nz) in enumerate(parts):
        L = engine.factor(p)[:n, :n]
        K = o.compute_cov_matrix_vectorized(nd, nz, ts)
        assert np.max(np.abs(L @ L.T - K)) <= 1e-12 * np.max(np.abs(K))
        U = sla.cholesky(K, lower=False)
        assert np.max(np.abs(L.T - U)) <= 1e-9 * np.max(np.abs(U))


def test_full_size_repeated_runs_never_corrupt_a_particle(engine):
    """The persistent kernel's operand pipeline (TMA copies, mbarriers, cross-CTA counters) under full load:
    every run of the headline workload must report info == 0 everywhere and reproduce the first run bit for bit
    (a rare pipeline race shows up as ONE particle with a wrong tile and a non-positive pivot)."""
    n, P = 2048, 64
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p) for p in range(P)]
    first, info = gpu_lmls(engine, parts, ts, xs)
    assert np.all(info == 0) and np.all(np.isfinite(first))
    for _ in range(7):
        again, info = gpu_lmls(engine, parts, ts, xs)
        assert np.all(info == 0) and np.array_equal(again, first)


# ---- BASELINE.json full sizes: size-independent properties ------------------------------------

def test_full_size_n2048_p64_properties(engine):
    """configs[1]: n=2048, 64 particles, Plus(Times(SE, Periodic), Linear)."""
    n, P = 2048, 64
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p) for p in range(P)]
    got, info = gpu_lmls(engine, parts, ts, xs)
    assert np.all(info == 0) and np.all(np.isfinite(got))
    # (a) spot-check against the oracle (0.5 s of CPU each)
    for p in (0, 17, 63):
        assert abs(got[p] - o.log_marginal_likelihood(*parts[p], ts, xs)) <= LML_RTOL * abs(got[p])
    # (b) determinism: bitwise identical on re-run
    again, _ = gpu_lmls(engine, parts, ts, xs)
    assert np.array_equal(got, again)
    # (c) the likelihood of exchangeable data is invariant to a joint permutation of (ts, xs)
    perm = np.random.default_rng(5).permutation(n)
    permuted, info = gpu_lmls(engine, parts, ts[perm], xs[perm])
    assert np.all(info == 0)
    assert H.rel_err(permuted, got) <= 1e-9
    # (d) chain rule on a prefix: LML(n) - LML(n-64) = predictive logpdf of the last 64 points
    p = 5
    head, _ = gpu_lmls(engine, [parts[p]], ts[: n - 64], xs[: n - 64])
    mu, cov = o.predictive_mvn(parts[p][0], parts[p][1], ts[: n - 64], xs[: n - 64], ts[n - 64:])
    assert o.mvn_logpdf(xs[n - 64:], mu, cov) == pytest.approx(got[p] - head[0], rel=1e-6)


def test_full_size_n8192_one_particle(engine):
    """configs[2] shape (n=8192): one particle against the oracle, one more for determinism."""
    n = 8192
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(3), o.synthetic_particle(3)]
    got, info = gpu_lmls(engine, parts, ts, xs)
    assert np.all(info == 0)
    assert got[0] == got[1]
    ref = o.log_marginal_likelihood(*parts[0], ts, xs)
    assert abs(got[0] - ref) <= LML_RTOL * abs(ref)


def test_scheduler_stress_repeated_runs_are_bitwise_stable(engine):
    """The persistent kernel resolves dependencies with spin-waits on counters; a missing
    dependency (like the diagonal-block inverses being overwritten by the next block column while
    the current one is still being solved) shows up as run-to-run differences.  Batch shapes that
    exercise the look-ahead split items (nt >= 4) with few and many particles, 6 runs each."""
    for n, P in ((640, 3), (1100, 7), (2048, 4), (2048, 16), (2300, 5)):
        ts, xs = o.synthetic_series(n)
        parts = [o.synthetic_particle(p) for p in range(P)]
        first, info = gpu_lmls(engine, parts, ts, xs)
        assert np.all(info == 0)
        ref = oracle_lmls(parts[:2], ts, xs)
        assert H.rel_err(first[:2], ref) <= LML_RTOL_TIGHT
        for _ in range(5):
            again, _ = gpu_lmls(engine, parts, ts, xs)
            assert np.array_equal(first, again), (n, P)
    # all particles identical => all results identical, whatever the interleaving of their items
    ts, xs = o.synthetic_series(1500)
    parts = [o.synthetic_particle(11)] * 24
    got, _ = gpu_lmls(engine, parts, ts, xs)
    assert np.all(got == got[0])


def test_gram_items_equal_the_separate_gram_launch_whatever_their_lead(monkeypatch):
    """The Gram units as items of the persistent kernel's queue (csrc/agp_chol_gram.cu) run the code of the separate
    agp_gramfill_kernel launch, so the LMLs must be bitwise the same — also with a lead of 1 or 5 items, where the first
    readers of a tile really wait for the unit's flag (a lead that small found the missing wait of the diagonal-tile
    item for the other row half of its tile)."""
    import autogp.jl_b200 as agp

    def run(fuse, lead, cases):
        monkeypatch.setenv("AGP_OZAKI", "0")   # the two Gram paths of the single-launch FP64 schedule (the hybrid schedule has its own tests)
        monkeypatch.setenv("AGP_FUSE_GRAM", "1" if fuse else "0")
        monkeypatch.setenv("AGP_GRAM_LEAD", str(lead))
        eng = agp.Engine(0)
        out = []
        for n, P in cases:
            ts, xs = o.synthetic_series(n)
            parts = [o.synthetic_particle(p, "se*per+lin" if p % 2 == 0 else "ge+per*lin") for p in range(P)]
            for _ in range(3):
                lml, info = gpu_lmls(eng, parts, ts, xs)
                assert np.all(info == 0) and eng.gram_items()[0] == bool(fuse)
                out.append(lml.copy())
        eng.close()
        return out

    cases = [(100, 3), (300, 70), (512, 64), (1100, 40), (2048, 24)]
    want = run(False, 0, cases)
    ref = oracle_lmls([o.synthetic_particle(p, "se*per+lin" if p % 2 == 0 else "ge+per*lin") for p in range(2)], *o.synthetic_series(300))
    assert H.rel_err(want[3][:2], ref) <= LML_RTOL_TIGHT
    for lead in (1, 5, 148, 0):
        got = run(True, lead, cases)
        for a, b in zip(want, got):
            assert np.array_equal(a, b), lead


# ---- §8 f-2: block-append continuation of the factorisation ------------------------------------

def test_lml_block_append_is_bitwise_the_full_recompute(engine):
    """agp_lml_run_append keeps the tile rows the previous prefix completed and computes only the new
    ones; the arithmetic per tile is the same, so the result must equal a from-scratch run on the
    same prefix bit for bit (and the oracle to the usual tolerance)."""
    import autogp.jl_b200 as agp
    from autogp.jl_b200 import _lib

    n_full, P = 1500, 6
    ts, xs = o.synthetic_series(n_full)
    parts = [o.synthetic_particle(p, t) for p, t in zip(range(P), ["se*per+lin", "se+wn", "ge+per*lin", "se*per+lin", "cp(lin,se)", "se*per+lin"])]
    nodes, noises = [H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
    fresh = agp.Engine(0)
    fresh.set_hybrid(0)   # the reference run of every prefix on the same FP64 schedule as the continuation (from 12 block columns on a
                          # from-scratch run would take the hybrid schedule: equal to 1e-11, not bit for bit — tests/test_hybrid_gpu.py)
    engine.upload(nodes, noises, ts, xs)
    with pytest.raises(_lib.AgpError):   # nothing factored yet
        engine.run_append()
    engine.set_prefix(200)
    engine.run()
    got, info = engine.fetch()
    assert np.all(info == 0)
    for n in (200, 456, 512, 513, 900, 1279, 1500):   # same size, inside a tile, tile-aligned, across tiles
        engine.set_prefix(n)
        engine.run_append()
        got, info = engine.fetch()
        ref, info_ref = fresh.lml_batch(nodes, noises, ts[:n], xs[:n])
        assert np.all(info == 0) and np.all(info_ref == 0)
        assert np.array_equal(got, ref), n
        assert H.rel_err(got[:2], oracle_lmls(parts[:2], ts[:n], xs[:n])) <= LML_RTOL_TIGHT
    # a shrinking prefix, or a failed factorisation, must be refused (caller falls back to run())
    engine.set_prefix(700)
    with pytest.raises(_lib.AgpError):
        engine.run_append()
    engine.upload([agp.Constant(1.0)], [-2.0], ts[:300], xs[:300])
    engine.set_prefix(150)
    engine.run()
    _, info = engine.fetch()
    assert info[0] != 0
    engine.set_prefix(300)
    with pytest.raises(_lib.AgpError):
        engine.run_append()
    fresh.close()


# ---- §8 f-3: predictive conditional MVN -----------------------------------------------------------

@pytest.mark.parametrize("n,m", [(200, 50), (128, 128), (333, 1), (700, 300), (50, 260)])
def test_predictive_mvn_matches_oracle(engine, n, m):
    """Distributions.MvNormal(node, noise, ts, xs, ts_pred) (src/GP.jl:731-758) for a ragged batch of
    kernels: mean and covariance against the oracle's dense restatement."""
    rng = np.random.default_rng(n * 1000 + m)
    ts, xs = o.synthetic_series(n)
    ts_pred = np.sort(rng.uniform(-0.1, 1.3, size=m))
    trees = ["se*per+lin", "se+wn", "ge+per*lin", "cp(lin,se)"]
    parts = [o.synthetic_particle(7 + p, t) for p, t in enumerate(trees)]
    nodes, noises = [H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
    noise_pred = [0.0, 0.3, noises[2], 1e-3]
    mean, cov, info = engine.predict_batch(nodes, noises, ts, xs, ts_pred, noise_pred)
    assert np.all(info == 0)
    for p, (nd, nz) in enumerate(parts):
        mu_ref, cov_ref = o.predictive_mvn(nd, nz, ts, xs, ts_pred, noise_pred=noise_pred[p])
        scale = max(1.0, float(np.max(np.abs(cov_ref))))
        assert np.max(np.abs(mean[p] - mu_ref)) <= 1e-8 * max(1.0, float(np.max(np.abs(mu_ref)))), p
        assert np.max(np.abs(cov[p] - cov_ref)) <= 1e-8 * scale, p
        assert np.array_equal(cov[p], cov[p].T)
    # default noise_pred = noise (src/GP.jl:738)
    mean2, cov2, _ = engine.predict_batch(nodes[:1], noises[:1], ts, xs, ts_pred)
    mu_ref, cov_ref = o.predictive_mvn(parts[0][0], parts[0][1], ts, xs, ts_pred)
    assert np.max(np.abs(cov2[0] - cov_ref)) <= 1e-8 * max(1.0, float(np.max(np.abs(cov_ref))))
    assert np.array_equal(mean2[0], mean[0])


def test_predictive_likelihood_identity_of_the_reference(engine):
    """test/experiment_hmc.jl:111-132: logpdf(MvNormal(node, noise, ts_obs, xs_obs, ts_test), xs_test)
    == LML(ts_all) - LML(ts_obs), on its three benchmark kernels (:180-184) — here with BOTH sides from
    the GPU path (predictive MVN from agp_predict_batch, LMLs from agp_lml_batch)."""
    import autogp.jl_b200 as agp

    ts_all = np.linspace(0, 10, 1000)
    rng = np.random.default_rng(42)
    xs_all = np.sin(ts_all) + 0.1 * rng.normal(size=ts_all.size)
    n_obs = 200
    cases = [(agp.SquaredExponential(2.0), 0.01), (agp.Plus(agp.Linear(0.5), agp.Periodic(2.0, 1.0)), 0.05),
             (agp.ChangePoint(agp.Linear(0.5), agp.Linear(1.5), 1.0, 0.001), 0.001)]
    for node, noise in cases:
        mean, cov = agp.predictive_mvn(node, noise, ts_all[:n_obs], xs_all[:n_obs], ts_all[n_obs:], engine=engine)
        lhs = o.mvn_logpdf(xs_all[n_obs:], mean, cov)
        lml_all = agp.mvnormal_logpdf(node, noise, ts_all, xs_all, engine=engine)
        lml_obs = agp.mvnormal_logpdf(node, noise, ts_all[:n_obs], xs_all[:n_obs], engine=engine)
        assert lhs == pytest.approx(lml_all - lml_obs, rel=1.5e-8)


def test_predictive_not_positive_definite_raises(engine):
    import autogp.jl_b200 as agp

    ts, xs = o.synthetic_series(150)
    with pytest.raises(agp.PosDefException):
        agp.predictive_mvn(agp.Constant(1.0), -2.0, ts, xs, np.linspace(0, 1, 5), engine=engine)


# ---- §8 f-1: gradient of the LML ---------------------------------------------------------------------

GRAD_RTOL = 2e-6   # finite-difference oracle: truncation + round-off of central differences, not the GPU path


def _grad_close(got, ref, rtol=GRAD_RTOL):
    got, ref = np.asarray(got), np.asarray(ref)
    return np.all(np.abs(got - ref) <= rtol * np.maximum(1.0, np.abs(ref)) + rtol * np.max(np.abs(ref)))


@pytest.mark.parametrize("n", [1, 37, 128, 200, 300])
def test_lml_gradient_matches_finite_difference_oracle(engine, n):
    """agp_lml_grad_batch against two CPU routes that never differentiate a kernel analytically
    (central differences of the oracle LML; 1/2 tr((aa' - K^-1) dK) with dK by central differences),
    on a ragged batch covering every leaf type and operator."""
    ts, xs = o.synthetic_series(max(n, 2))
    ts, xs = ts[:n], xs[:n]
    trees = ["se*per+lin", "se+wn", "ge+per*lin", "cp(lin,se)"]
    parts = [o.synthetic_particle(20 + p, t) for p, t in enumerate(trees)]
    parts.append((o.Plus(o.Times(o.Constant(0.7), o.GammaExponential(0.3, 1.4, 0.8)),
                         o.ChangePoint(o.Periodic(0.5, 0.3, 1.2), o.Linear(0.2, 0.4, 0.9), 0.45, 0.08)), 0.05))
    nodes, noises = [H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
    lml, grads, gnoise, info = engine.lml_grad_batch(nodes, noises, ts, xs)
    assert np.all(info == 0)
    assert H.rel_err(lml, oracle_lmls(parts, ts, xs)) <= LML_RTOL_TIGHT
    for p, (nd, nz) in enumerate(parts):
        g_fd, gn_fd = o.lml_grad_fd(nd, nz, ts, xs)
        g_dn, gn_dn = o.lml_grad_dense_fd(nd, nz, ts, xs)
        assert len(grads[p]) == len(g_fd)
        assert _grad_close(grads[p], g_dn), (p, grads[p], g_dn)
        assert _grad_close(grads[p], g_fd, 2e-5), (p, grads[p], g_fd)
        assert abs(gnoise[p] - gn_dn) <= 1e-8 * max(1.0, abs(gn_dn)), (p, gnoise[p], gn_dn)
        assert abs(gnoise[p] - gn_fd) <= 2e-5 * max(1.0, abs(gn_fd))


def test_lml_gradient_full_size_directional_derivative(engine):
    """n = 2048 (BASELINE.json configs[1] tree): the gradient must predict the change of the GPU LML
    itself along a random direction in parameter space; bitwise reproducible."""
    import autogp.jl_b200 as agp

    n, P = 2048, 4
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p) for p in range(P)]
    nodes, noises = [H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
    lml, grads, gnoise, info = engine.lml_grad_batch(nodes, noises, ts, xs)
    assert np.all(info == 0)
    again = engine.lml_grad_batch(nodes, noises, ts, xs)
    assert all(np.array_equal(a, b) for a, b in zip(grads, again[1])) and np.array_equal(gnoise, again[2])
    rng = np.random.default_rng(9)
    for p, (nd, nz) in enumerate(parts):
        params = o.encode_program(nd)[2]
        d = rng.normal(size=params.size) * np.abs(params)
        dn = rng.normal() * nz
        hstep = 1e-6
        up = H.to_agp(o.with_params(nd, params + hstep * d))
        dnn = H.to_agp(o.with_params(nd, params - hstep * d))
        f_up, _ = engine.lml_batch([up], [nz + hstep * dn], ts, xs)
        f_dn, _ = engine.lml_batch([dnn], [nz - hstep * dn], ts, xs)
        fd = (f_up[0] - f_dn[0]) / (2 * hstep)
        an = float(np.dot(grads[p], d) + gnoise[p] * dn)
        assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)), (p, fd, an)
    # at this size both calls take the hybrid schedule (int8 digit-plane contractions, tests/test_hybrid_gpu.py); the LML of
    # the gradient call equals the plain batch's to 1e-11 there and to 1e-13 on the FP64 schedule
    assert H.rel_err(lml, engine.lml_batch(nodes, noises, ts, xs)[0]) <= 1e-11
    engine.set_hybrid(0)
    try:
        lml_f = engine.lml_grad_batch(nodes, noises, ts, xs)[0]
        assert H.rel_err(lml_f, engine.lml_batch(nodes, noises, ts, xs)[0]) <= 1e-13
        assert H.rel_err(lml_f, lml) <= 1e-11
    finally:
        engine.set_hybrid(-1)


@pytest.mark.parametrize("n", [1, 77, 128, 300, 1000])
def test_noise_only_gradient_matches_the_full_gradient_call_and_the_oracle(engine, n):
    """agp_lml_grad_noise_batch (factorisation + trtri, 1/2 (|alpha|^2 - |L^-1|_F^2)) against agp_lml_grad_batch's noise
    gradient (1/2 tr(alpha alpha' - K^-1) out of the full inverse) and the dense oracle route; ragged batch, every node type."""
    ts, xs = o.synthetic_series(max(n, 2))
    ts, xs = ts[:n], xs[:n]
    trees = ["se*per+lin", "se+wn", "ge+per*lin", "cp(lin,se)"]
    parts = [o.synthetic_particle(40 + p, t) for p, t in enumerate(trees)]
    nodes, noises = [H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
    lml_f, _, gn_f, info_f = engine.lml_grad_batch(nodes, noises, ts, xs)
    lml, gn, info = engine.lml_grad_noise_batch(nodes, noises, ts, xs)
    assert np.all(info == 0) and np.all(info_f == 0)
    assert np.array_equal(lml, lml_f)                      # same factorisation items, same arithmetic
    assert H.rel_err(lml, oracle_lmls(parts, ts, xs)) <= LML_RTOL_TIGHT
    for p, (nd, nz) in enumerate(parts):
        # both are differences of two large sums (|alpha|^2 and tr K^-1): compare at the scale of the larger one
        K = o.compute_cov_matrix_vectorized(nd, nz, ts)
        Kinv = np.linalg.inv(K)
        alpha = Kinv @ xs
        scale = max(float(alpha @ alpha), float(np.trace(Kinv)), 1.0)
        ref = 0.5 * (float(alpha @ alpha) - float(np.trace(Kinv)))
        assert abs(gn[p] - gn_f[p]) <= 1e-12 * scale, (p, gn[p], gn_f[p])
        assert abs(gn[p] - ref) <= 1e-9 * scale, (p, gn[p], ref)
    again = engine.lml_grad_noise_batch(nodes, noises, ts, xs)
    assert np.array_equal(gn, again[1])                    # fixed summation order


def test_noise_only_gradient_full_size_and_edge_cases(engine):
    import autogp.jl_b200 as agp

    n, P = 2048, 6
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p) for p in range(P)]
    nodes, noises = [H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
    lml, gn, info = engine.lml_grad_noise_batch(nodes, noises, ts, xs)
    assert np.all(info == 0)
    hstep = 1e-6
    for p in range(P):   # the derivative of the GPU LML itself along the noise
        up, _ = engine.lml_batch([nodes[p]], [noises[p] * (1 + hstep)], ts, xs)
        dn, _ = engine.lml_batch([nodes[p]], [noises[p] * (1 - hstep)], ts, xs)
        fd = (up[0] - dn[0]) / (2 * hstep * noises[p])
        assert abs(fd - gn[p]) <= 1e-5 * max(1.0, abs(fd)), (p, fd, gn[p])
    # a plain LML batch after the augmented one is unaffected by the leftover state (at this size both take the hybrid
    # schedule: equal to 1e-11; bitwise on the FP64 schedule)
    assert H.rel_err(lml, engine.lml_batch(nodes, noises, ts, xs)[0]) <= 1e-11
    engine.set_hybrid(0)
    try:
        lml_f = engine.lml_grad_noise_batch(nodes, noises, ts, xs)[0]
        assert np.array_equal(engine.lml_batch(nodes, noises, ts, xs)[0], lml_f)
    finally:
        engine.set_hybrid(-1)
    # empty data, failed factorisation, and no program-size limit on this path
    lml, gn, info = engine.lml_grad_noise_batch([agp.SquaredExponential(0.3, 1.0)], [0.1], ts[:0], xs[:0])
    assert lml[0] == 0.0 and gn[0] == 0.0 and info[0] == 0
    lml, gn, info = engine.lml_grad_noise_batch([agp.Constant(1.0), agp.SquaredExponential(0.3, 1.0)], [-2.0, 0.1], ts[:150], xs[:150])
    assert info[0] != 0 and np.isnan(lml[0]) and np.isnan(gn[0]) and info[1] == 0 and np.isfinite(gn[1])
    big = agp.Constant(1.0)
    for _ in range(70):
        big = agp.Plus(big, agp.Constant(0.5))
    lml, gn, info = engine.lml_grad_noise_batch([big], [0.1], ts[:150], xs[:150])
    assert info[0] == 0 and np.isfinite(gn[0])


@pytest.mark.parametrize("n,m", [(60, 5), (300, 130), (1000, 33)])
def test_infer_gp_sum_matches_the_oracle(engine, n, m):
    """agp_predict_sum_batch against the restated infer_gp_sum (src/GP.jl:904-993): three particles with different
    summands in one call, prediction points inside and beyond the data, noise_pred default and explicit."""
    import autogp.jl_b200 as agp

    ts, xs = o.synthetic_series(n)
    tp = np.concatenate([np.linspace(1.0, 1.3, m - m // 2), ts[: m // 2]])
    sets = [
        [o.Linear(0.2, 0.4, 0.9), o.Periodic(0.5, 0.3, 1.2), o.GammaExponential(0.3, 1.4, 0.8)],
        [o.Times(o.SquaredExponential(0.4, 0.7), o.Periodic(0.9, 0.25, 1.1)), o.Constant(0.3), o.Linear(0.5, 0.1, 0.6)],
        [o.ChangePoint(o.Linear(0.3, 0.2, 0.8), o.SquaredExponential(0.2, 0.5), 0.6, 0.05), o.WhiteNoise(0.05), o.SquaredExponential(0.1, 0.2)],
    ]
    noises = [0.07, 0.02, 0.11]
    for npred in (None, [0.0, 0.3, 0.01]):
        mean, cov, info = engine.predict_sum_batch([[H.to_agp(nd) for nd in st] for st in sets], noises, ts, xs, tp, npred)
        assert np.all(info == 0)
        for p, st in enumerate(sets):
            mu_o, cov_o, idx = o.infer_gp_sum(st, noises[p], ts, xs, tp, noise_pred=None if npred is None else npred[p])
            cov_o = cov_o - o.GP_JITTER * np.eye(cov_o.shape[0])        # the C-ABI leaves the MvNormal's jitter to the caller
            scale = max(np.max(np.abs(cov_o)), 1e-12)
            assert np.max(np.abs(mean[p] - mu_o)) <= 1e-8 * max(1.0, np.max(np.abs(mu_o))), (p, np.max(np.abs(mean[p] - mu_o)))
            assert np.max(np.abs(cov[p] - cov_o)) <= 1e-8 * scale, (p, np.max(np.abs(cov[p] - cov_o)), scale)
            assert np.array_equal(cov[p], cov[p].T)
    # the single-particle mirror of the reference function (adds the jitter, raises on a non-PD training block)
    mu1, cov1, idx1 = agp.infer_gp_sum([H.to_agp(nd) for nd in sets[0]], noises[0], ts, xs, tp)
    mu_o, cov_o, idx_o = o.infer_gp_sum(sets[0], noises[0], ts, xs, tp)
    assert [list(r) for r in idx1["F"]] == [list(r) for r in idx_o["F"]] and list(idx1["X"]) == list(idx_o["X"])
    assert np.max(np.abs(cov1 - cov_o)) <= 1e-8 * np.max(np.abs(cov_o)) and np.max(np.abs(mu1 - mu_o)) <= 1e-8 * max(1.0, np.max(np.abs(mu_o)))


@pytest.mark.parametrize("n,m", [(300, 1), (300, 130), (1000, 300), (0, 5)])
def test_predictive_marginals_are_the_diagonal_of_the_predictive_covariance(engine, n, m):
    """agp_predict_marginals_batch against agp_predict_batch (bitwise: the same diagonal-tile items) and the oracle."""
    ts, xs = o.synthetic_series(max(n, 2))
    ts, xs = ts[:n], xs[:n]
    tp = np.linspace(0.9, 1.3, m)
    parts = [o.synthetic_particle(60 + p, t) for p, t in enumerate(["se*per+lin", "ge+per*lin", "cp(lin,se)"])]
    nodes, noises = [H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
    for npred in (None, [0.0, 0.2, 0.05]):
        mean, var, info = engine.predict_marginals_batch(nodes, noises, ts, xs, tp, npred)
        mean_f, cov_f, info_f = engine.predict_batch(nodes, noises, ts, xs, tp, npred)
        assert np.all(info == 0) and np.array_equal(mean, mean_f)
        assert np.array_equal(var, np.diagonal(cov_f, axis1=1, axis2=2))
        if n > 0:
            for p, (nd, nz) in enumerate(parts):
                mu_o, cov_o = o.predictive_mvn(nd, nz, ts, xs, tp, noise_pred=None if npred is None else npred[p])
                assert np.max(np.abs(var[p] - np.diag(cov_o))) <= 1e-8 * max(np.max(np.abs(cov_o)), 1e-12)


@pytest.mark.parametrize("n,m", [(200, 37), (128, 128), (1000, 5), (0, 40)])
def test_predictive_logpdf_through_the_append_identity(engine, n, m):
    """model.predictive_logpdfs (LML(ts u ts_new) - LML(ts), second factorisation continued from the first) against the
    oracle's logpdf of the dense predictive MVN — the identity of test/experiment_hmc.jl:111-132."""
    import autogp.jl_b200 as agp

    ts_all, xs_all = o.synthetic_series(n + m)
    ts, xs, tn, xn = ts_all[:n], xs_all[:n], ts_all[n:], xs_all[n:]
    parts = [o.synthetic_particle(70 + p, t) for p, t in enumerate(["se*per+lin", "ge+per*lin", "se+wn"])]
    got = agp.predictive_logpdfs([H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs, tn, xn, engine=engine)
    for p, (nd, nz) in enumerate(parts):
        if n > 0:
            mu, cov = o.predictive_mvn(nd, nz, ts, xs, tn)
        else:
            mu, cov = np.zeros(m), o.compute_cov_matrix_vectorized(nd, nz, tn)
        want = o.mvn_logpdf(xn, mu, cov)
        assert abs(got[p] - want) <= 1e-8 * max(abs(want), 1.0) + 1e-9 * abs(o.log_marginal_likelihood(nd, nz, ts_all, xs_all)), (p, got[p], want)


def test_predict_mvn_sum_flow_split_then_infer(engine):
    """predict_mvn_sum (src/api.jl): split every particle's kernel on a base-kernel type, then infer_gp_sum over the two
    sides — against the oracle on the same pairs, incl. a particle whose split has an empty (Constant(0)) side."""
    import autogp.jl_b200 as agp

    ts, xs = o.synthetic_series(200)
    tp = np.linspace(1.0, 1.25, 40)
    parts = [o.synthetic_particle(3, "se*per+lin"), o.synthetic_particle(4, "ge+per*lin"), o.synthetic_particle(5, "se+wn")]
    pairs = [agp.split_kernel_sop(H.to_agp(nd), agp.Periodic) for nd, _ in parts]
    assert pairs[2][0] == agp.Constant(0.0)                      # no Periodic factor anywhere in particle 2
    noises = [nz for _, nz in parts]
    mean, cov, info = engine.predict_sum_batch([list(pr) for pr in pairs], noises, ts, xs, tp)
    assert np.all(info == 0)
    for p, pr in enumerate(pairs):
        mu_o, cov_o, idx = o.infer_gp_sum([H.from_agp(nd) for nd in pr], noises[p], ts, xs, tp)
        cov_o = cov_o - o.GP_JITTER * np.eye(cov_o.shape[0])
        assert np.max(np.abs(mean[p] - mu_o)) <= 1e-8 * max(1.0, np.max(np.abs(mu_o)))
        assert np.max(np.abs(cov[p] - cov_o)) <= 1e-8 * np.max(np.abs(cov_o))
    assert np.all(mean[2, :40] == 0.0) and np.all(cov[2, :40, :] == 0.0)     # the empty side has a zero posterior


def test_infer_gp_sum_edge_cases(engine):
    import autogp.jl_b200 as agp
    from autogp.jl_b200 import _lib
    from autogp.jl_b200.model import PosDefException

    ts, xs = o.synthetic_series(150)
    tp = np.array([1.05])
    # one summand, one prediction point: F* and X* differ by the noise only
    mean, cov, info = engine.predict_sum_batch([[agp.SquaredExponential(0.3, 1.0)]], [0.1], ts, xs, tp)
    assert info[0] == 0 and mean.shape == (1, 2) and abs(mean[0, 0] - mean[0, 1]) <= 1e-12
    assert abs(cov[0, 1, 1] - cov[0, 0, 0] - 0.1) <= 1e-12 and abs(cov[0, 0, 1] - cov[0, 0, 0]) <= 1e-12
    # X* block == agp_predict_batch of the Plus kernel
    nodes = [agp.Linear(0.2, 0.4, 0.9), agp.Periodic(0.5, 0.3, 1.2)]
    tp = np.linspace(1.0, 1.2, 140)
    mean, cov, info = engine.predict_sum_batch([nodes], [0.05], ts, xs, tp)
    m2, c2, _ = engine.predict_batch([agp.Plus(nodes[0], nodes[1])], [0.05], ts, xs, tp)
    X = slice(2 * 140, 3 * 140)
    assert np.max(np.abs(mean[0, X] - m2[0])) <= 1e-10 and np.max(np.abs(cov[0, X, X] - c2[0])) <= 1e-10 * np.max(np.abs(c2[0]))
    # no observations: the posterior is the prior
    mean, cov, info = engine.predict_sum_batch([nodes], [0.05], ts[:0], xs[:0], tp[:4])
    K0 = o.compute_cov_matrix_vectorized(o.Linear(0.2, 0.4, 0.9), 0.0, tp[:4])
    assert np.all(mean == 0.0) and np.max(np.abs(cov[0, :4, :4] - K0)) <= 1e-14 and np.all(cov[0, :4, 4:8] == 0.0)
    # ragged summand counts and a non-PD training block
    with pytest.raises(ValueError):
        engine.predict_sum_batch([nodes, nodes[:1]], [0.05, 0.05], ts, xs, tp)
    with pytest.raises(PosDefException):
        agp.infer_gp_sum([agp.Constant(1.0)], -2.0, ts, xs, tp[:3], engine=engine)
    # a later plain LML batch is unaffected
    lml, info = engine.lml_batch([agp.Plus(nodes[0], nodes[1])], [0.05], ts, xs)
    assert info[0] == 0 and H.rel_err(lml, [o.log_marginal_likelihood(o.Plus(o.Linear(0.2, 0.4, 0.9), o.Periodic(0.5, 0.3, 1.2)), 0.05, ts, xs)]) <= LML_RTOL_TIGHT


def test_lml_gradient_edge_cases(engine):
    import autogp.jl_b200 as agp
    from autogp.jl_b200 import _lib

    ts, xs = o.synthetic_series(150)
    # empty data: score 0, zero gradient
    lml, grads, gnoise, info = engine.lml_grad_batch([agp.SquaredExponential(0.3, 1.0)], [0.1], ts[:0], xs[:0])
    assert lml[0] == 0.0 and np.all(grads[0] == 0.0) and gnoise[0] == 0.0 and info[0] == 0
    # not positive definite: NaN gradient, LAPACK info
    lml, grads, gnoise, info = engine.lml_grad_batch([agp.Constant(1.0), agp.SquaredExponential(0.3, 1.0)], [-2.0, 0.1], ts, xs)
    assert info[0] != 0 and np.isnan(lml[0]) and np.all(np.isnan(grads[0])) and np.isnan(gnoise[0])
    assert info[1] == 0 and np.all(np.isfinite(grads[1]))
    # a kernel beyond the general variant's tape (512 levels; 180 Periodic leaves push 540) is an error, not a crash
    big = agp.Periodic(0.5, 0.3, 1.0)
    for _ in range(179):
        big = agp.Plus(big, agp.Periodic(0.5, 0.3, 0.01))
    with pytest.raises(_lib.AgpError):
        engine.lml_grad_batch([big], [0.1], ts, xs)


def test_lml_gradient_of_kernels_beyond_64_nodes_and_64_parameters(engine):
    """Structure learning with max_depth = -1 proposes large kernels now and then; their gradient takes the general
    variant of agp_grad_kernel (program from global memory, parameters in windows of 64).  Checked against the
    finite-difference oracle, against the tuned variant on a small kernel riding in the same batch, and on a chain
    of 71 Constants whose 71 gradients must all be the same number."""
    import autogp.jl_b200 as agp

    n = 60
    ts, xs = o.synthetic_series(n)
    rng = np.random.default_rng(3)

    def leaf(k):
        u = lambda lo, hi: float(rng.uniform(lo, hi))
        return [lambda: o.Linear(u(0.1, 0.9), u(0.1, 0.5), u(0.05, 0.3)), lambda: o.SquaredExponential(u(0.1, 0.6), u(0.05, 0.3)),
                lambda: o.GammaExponential(u(0.1, 0.6), u(0.6, 1.8), u(0.05, 0.3)), lambda: o.Periodic(u(0.3, 1.2), u(0.1, 0.5), u(0.05, 0.3)),
                lambda: o.Constant(u(0.05, 0.3))][k % 5]()

    def grow(k0, k1):   # a balanced tree over leaves k0..k1-1 with Plus / Times / ChangePoint alternating by level
        if k1 - k0 == 1:
            return leaf(k0)
        mid = (k0 + k1) // 2
        left, right = grow(k0, mid), grow(mid, k1)
        kind = (k1 - k0) % 3
        return o.Plus(left, right) if kind == 0 else o.Times(left, right) if kind == 1 else o.ChangePoint(left, right, 0.4 + 0.01 * (k0 % 7), 0.05)

    big = grow(0, 48)                                  # 95 nodes, > 64 parameters, operand stack depth 6
    small = o.synthetic_particle(21, "se*per+lin")[0]
    assert len(o.unroll(big)) > 64 and len(o.encode_program(big)[2]) > 64
    nodes, noises = [H.to_agp(big), H.to_agp(small)], [0.07, 0.05]
    lml, grads, gnoise, info = engine.lml_grad_batch(nodes, noises, ts, xs)
    assert np.all(info == 0)
    assert H.rel_err(lml, oracle_lmls([(big, 0.07), (small, 0.05)], ts, xs)) <= LML_RTOL_TIGHT
    g_dn, gn_dn = o.lml_grad_dense_fd(big, 0.07, ts, xs)
    assert len(grads[0]) == len(g_dn)
    assert _grad_close(grads[0], g_dn), np.max(np.abs(grads[0] - g_dn))
    assert abs(gnoise[0] - gn_dn) <= 1e-8 * max(1.0, abs(gn_dn))
    # the small kernel: same numbers as from the tuned variant (a batch of its own)
    lml1, grads1, gnoise1, _ = engine.lml_grad_batch(nodes[1:], noises[1:], ts, xs)
    np.testing.assert_allclose(grads[1], grads1[0], rtol=1e-12, atol=0)
    assert lml[1] == lml1[0] and abs(gnoise[1] - gnoise1[0]) <= 1e-12 * abs(gnoise1[0])
    # 71 Constants: dK/dc = ones for every one of them
    chain = agp.Constant(1.0)
    for _ in range(70):
        chain = agp.Plus(chain, agp.Constant(0.5))
    lml, grads, gnoise, info = engine.lml_grad_batch([chain], [0.1], ts, xs)
    assert info[0] == 0 and len(grads[0]) == 71
    np.testing.assert_allclose(grads[0], grads[0][0], rtol=1e-13)
    g1 = engine.lml_grad_batch([agp.Constant(36.0)], [0.1], ts, xs)[1][0][0]
    assert grads[0][0] == pytest.approx(g1, rel=1e-10)


def test_reserve_sizes_the_buffers_once(engine):
    """agp_reserve: after reserving for the largest call, growing the series (the data-annealing pattern) re-allocates
    nothing — the factor's device address stays put — and results are those of an engine that never reserved."""
    import autogp.jl_b200 as agp

    ts, xs = o.synthetic_series(700)
    parts = [o.synthetic_particle(40 + p, "se*per+lin") for p in range(3)]
    nodes, noises = [H.to_agp(nd) for nd, _ in parts], [nz for _, nz in parts]
    eng = agp.Engine(0)
    eng.reserve(700, 3, gradient=True)
    ptrs, out = set(), []
    for n in (100, 300, 700):
        lml, info = eng.lml_batch(nodes, noises, ts[:n], xs[:n])
        ptrs.add(eng.device_results()[0])
        _, g, gn, ginfo = eng.lml_grad_batch(nodes, noises, ts[:n], xs[:n])
        assert np.all(info == 0) and np.all(ginfo == 0)
        out.append((lml, g, gn))
    for n, (lml, g, gn) in zip((100, 300, 700), out):
        lml0, _ = engine.lml_batch(nodes, noises, ts[:n], xs[:n])
        _, g0, gn0, _ = engine.lml_grad_batch(nodes, noises, ts[:n], xs[:n])
        assert np.array_equal(lml, lml0) and np.array_equal(gn, gn0) and all(np.array_equal(a, b) for a, b in zip(g, g0))
    with pytest.raises(Exception):
        eng.reserve(-1, 3)
    eng.close()
