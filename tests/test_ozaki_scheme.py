"""CPU checks of the arithmetic behind the hybrid factorisation (csrc/agp_ozaki.cu): the digit planes,
the exact integer products per weight group, the recombination, the a-priori row bounds.  Plain NumPy / Python integers
restate what the CUDA kernels do (the kernels themselves are checked on the GPU: tests/test_hybrid_gpu.py,
tools/ozaki_update_test.cu)."""
import numpy as np
import pytest

import autogp_oracle as o

TB = 128


def digits_of(x, e):
    """agp_ozaki_digits.cuh: v = rint(x 2^(55 - e)), clamped to 56 bits, peeled into seven balanced base-256 digits, leading digit first."""
    v = np.rint(x * np.exp2(55.0 - e)).astype(np.int64)
    lim = 127 * ((256 ** 7 - 1) // 255)          # what seven balanced byte digits reach: 0.99608 2^55
    v = np.clip(v, -lim, lim)
    out = np.zeros((7,) + x.shape, dtype=np.int64)
    for q in range(6, 0, -1):
        d = ((v + 128) & 255) - 128
        out[q] = d
        v = (v - d) >> 8
    out[0] = v
    return out


def row_exponent(kdiag):
    """agp_ozaki_rowscale_kernel: e = ceil(log2 sqrt(K_rr)) through frexp (K_rr = m 2^ex, 1/2 <= m < 1), one more when
    sqrt(K_rr) is above 0.99 of 2^e (the digits reach 0.996 of the scale)."""
    _, ex = np.frexp(kdiag)
    e = (ex + 1) >> 1
    return e + (kdiag > 0.98 * np.exp2(2.0 * e))


def int8_product(A, B, ea, eb):
    """sum_j A_ij B_kj from the digit planes: exact integer sums per weight group g = p + q <= 6, recombined as the epilogue of
    agp_ozaki_update2_kernel does (groups 0..3 and 4..6, int64, one conversion each, one fma)."""
    da, db = digits_of(A, ea[:, None]), digits_of(B, eb[:, None])
    assert np.min(da) >= -128 and np.max(da) <= 127 and np.min(db) >= -128 and np.max(db) <= 127      # signed bytes
    G = [sum(da[p] @ db[g - p].T for p in range(g + 1)) for g in range(7)]
    assert max(int(np.max(np.abs(g))) for g in G) < 2 ** 31      # the int32 accumulators of TMEM do not overflow
    t0 = (G[0] << 24) + (G[1] << 16) + (G[2] << 8) + G[3]
    t1 = (G[4] << 16) + (G[5] << 8) + G[6]
    d = t1.astype(np.float64) * 2.0 ** -24 + t0.astype(np.float64)          # the epilogue's fma (NumPy: two roundings, same grade)
    scale = np.exp2(ea.astype(np.float64))[:, None] * np.exp2(eb.astype(np.float64))[None, :]
    return d * 2.0 ** -38 * scale


def test_digit_planes_represent_the_entry_to_2_pow_minus_56_of_the_row_scale():
    rng = np.random.default_rng(3)
    e = rng.integers(-6, 7, size=400)
    x = rng.uniform(-0.99, 0.99, size=(400, 64)) * np.exp2(e)[:, None] * np.exp2(-rng.integers(0, 40, size=(400, 64)))
    x[:, 0] = 0.99 * np.exp2(e)   # the largest entries the row scales admit, both signs
    x[:, 1] = -0.99 * np.exp2(e)
    x[:, 2] = 0.0
    d = digits_of(x, e[:, None])
    assert np.min(d) >= -128 and np.max(d) <= 127
    rep = sum(d[p].astype(np.longdouble) * np.longdouble(2.0) ** (-7 - 8 * p) for p in range(7)) * np.exp2(e)[:, None].astype(np.longdouble)
    assert np.max(np.abs((rep - x.astype(np.longdouble)).astype(np.float64)) / np.exp2(e)[:, None]) <= 2.0 ** -56


@pytest.mark.parametrize("n,tree", [(512, "se*per+lin"), (640, "ge+per*lin")])
def test_int8_contraction_of_a_real_factor_is_fp64_grade(n, tree):
    """T = K - L L^T over the block columns left of the last tile row of a benchmark factor: the int8 scheme with ONE scale per
    row from the Gram diagonal against a long-double reference; FP64 accumulation for comparison."""
    ts, xs = o.synthetic_series(n)
    node, noise = o.synthetic_particle(0, tree)
    K = o.compute_cov_matrix_vectorized(node, noise, ts)
    L = np.linalg.cholesky(K)
    e = row_exponent(np.diag(K))
    assert np.all(np.max(np.abs(L), axis=1) <= 0.9901 * np.exp2(e.astype(np.float64)))  # |L_ij| <= sqrt(K_ii) <= 0.99 2^e_i
    assert np.all(np.exp2(e.astype(np.float64)) <= 2.0 * np.sqrt(np.diag(K)) / 0.98)
    k = n // TB - 1
    A = L[k * TB:, :k * TB]
    exact = A.astype(np.longdouble) @ A.astype(np.longdouble).T
    got = int8_product(A, A, e[k * TB:], e[k * TB:])
    bound = np.sqrt(np.outer(np.diag(K)[k * TB:], np.diag(K)[k * TB:]))
    err_i8 = float(np.max(np.abs((got.astype(np.longdouble) - exact).astype(np.float64)) / bound))
    err_f64 = float(np.max(np.abs(((A @ A.T).astype(np.longdouble) - exact).astype(np.float64)) / bound))
    assert err_i8 <= 8 * np.sqrt(k * TB) * 2.0 ** -55        # stated error model: about sqrt(depth) 2^-55 of sqrt(K_ii K_kk)
    assert err_i8 <= 20 * max(err_f64, 1e-17)                 # the same league as FP64 accumulation


def test_appended_rows_of_the_gradient_calls_obey_the_noise_bound():
    """The rows of L^{-T} (appended [I 0] rows of agp_lml_grad_batch): sum_j (L^{-1})_jr^2 = (K^{-1})_rr <= 1 / noise."""
    n = 300
    ts, xs = o.synthetic_series(n)
    for p, tree in enumerate(["se*per+lin", "ge+per*lin", "se+wn", "cp(lin,se)"]):
        node, noise = o.synthetic_particle(p, tree)
        K = o.compute_cov_matrix_vectorized(node, noise, ts)
        Linv = np.linalg.inv(np.linalg.cholesky(K))
        rows = Linv.T                                            # appended row r = column r of L^{-1}
        assert np.all(np.sum(rows ** 2, axis=1) <= (1.0 / noise) * (1 + 1e-9))
        _, ex = np.frexp(1.0 / noise)
        e = max((ex + 1) >> 1, 0) + 1                            # agp_ozaki_rowscale_kernel, rows >= ld_obs
        assert np.max(np.abs(rows)) <= 2.0 ** e / 2
