"""Paper study for DESIGN.md §9(c): how many int8 slices would an Ozaki-style error-free split of the Cholesky trailing
update need for FP64-grade results?  (CPU only: the slice products are integer arithmetic, emulated exactly in int64.)

The contraction of the persistent kernel's panel item, sum_j L_ij L_kj^T over 128-deep block columns, is rebuilt from
int8 slices of the two operand tiles: every (row, 128-column tile) gets its own power-of-two scale (the largest exponent
of its entries), the scaled entries are cut into `s` signed 7-bit slices, and all slice pairs (p, q) with p + q < s are
multiplied exactly (what tcgen05.mma kind::i8 with int32 TMEM accumulators does: 128 x 127^2 < 2^31) and recombined in
FP64 — s (s + 1) / 2 int8 products per FP64 product.  Reported: the error of K - L L^T over the last block row of the
benchmark factor (n = 2048, tree se*per+lin: where the contraction is deepest and cancels most), relative to the
largest entry of the result, against a long-double reference — for plain FP64 accumulation (what DMMA does) and for
s = 6 .. 10 slices."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import autogp_oracle as o  # noqa: E402  (developer tool: the oracle only supplies the benchmark matrix)

TB = 128


def slices_of(tile, s):
    """tile [rows, 128] -> (int8 slices [s, rows, 128], scale exponents [rows]): tile ~ sum_p slices[p] * 2^(e - 7 (p + 1))."""
    mx = np.max(np.abs(tile), axis=1)
    e = np.where(mx > 0, np.floor(np.log2(np.where(mx > 0, mx, 1.0))).astype(np.int64) + 1, 0)   # |x| < 2^e
    r = tile * np.exp2(-e)[:, None]            # |r| < 1, exact (power of two)
    out = np.zeros((s,) + tile.shape, dtype=np.int64)
    for p in range(s):
        r = r * 128.0                           # exact
        q = np.rint(r)                          # round to nearest: |q| <= 64 after the first slice, <= 128 at most
        q = np.clip(q, -127, 127)
        out[p] = q.astype(np.int64)
        r = r - q                               # exact: |r| <= 1/2
    return out, e


def slices_with_scale(tile, s, e):
    """the same with given per-row exponents e (|tile[r]| < 2^e[r] must hold)"""
    r = tile * np.exp2(-e.astype(np.float64))[:, None]
    out = np.zeros((s,) + tile.shape, dtype=np.int64)
    for p in range(s):
        r = r * 128.0
        q = np.clip(np.rint(r), -127, 127)
        out[p] = q.astype(np.int64)
        r = r - q
    return out


def ozaki_product_row_scale(A, B, s, ea, eb):
    """One scale per ROW for the whole contraction (known before the factorisation: |L_ij| <= sqrt(K_ii)), so the int32
    group sums can run over every 128-deep tile and be recombined ONCE per output tile."""
    m, K = A.shape
    n = B.shape[0]
    G = [np.zeros((m, n), dtype=np.int64) for _ in range(s)]
    for t in range(K // TB):
        As = slices_with_scale(A[:, t * TB:(t + 1) * TB], s, ea)
        Bs = slices_with_scale(B[:, t * TB:(t + 1) * TB], s, eb)
        for d in range(s):
            for p in range(d + 1):
                G[d] += As[p] @ Bs[d - p].T
    assert max(int(np.max(np.abs(g))) for g in G) < 2 ** 31, "int32 accumulators would overflow"
    acc = np.zeros((m, n))
    for d in range(s - 1, -1, -1):
        acc += G[d].astype(np.float64) * 2.0 ** (-7 * (d + 2))
    return acc * np.exp2(ea.astype(np.float64))[:, None] * np.exp2(eb.astype(np.float64))[None, :]


def ozaki_product(A, B, s):
    """sum over 128-deep tiles of A[:, t] B[:, t]^T from int8 slices (exact integer products, FP64 recombination)."""
    m, K = A.shape
    n = B.shape[0]
    C = np.zeros((m, n))
    for t in range(K // TB):
        As, ea = slices_of(A[:, t * TB:(t + 1) * TB], s)
        Bs, eb = slices_of(B[:, t * TB:(t + 1) * TB], s)
        acc = np.zeros((m, n))
        for d in range(s - 1, -1, -1):          # smallest terms first
            G = np.zeros((m, n), dtype=np.int64)
            for p in range(d + 1):
                G += As[p] @ Bs[d - p].T        # exact: |sum| <= (d + 1) 128 127^2 < 2^31 for d < 8
            acc += G.astype(np.float64) * 2.0 ** (-7 * (d + 2))
        C += acc * np.exp2(ea)[:, None] * np.exp2(eb)[None, :]
    return C


def main():
    n = 2048
    ts, xs = o.synthetic_series(n)
    node, noise = o.synthetic_particle(0, "se*per+lin")
    K = o.compute_cov_matrix_vectorized(node, noise, ts)
    L = np.linalg.cholesky(K)
    k = n // TB - 1
    rows = slice(k * TB, n)
    A = L[rows, :k * TB]                        # the last tile row: A = B for its diagonal tile
    exact = (A.astype(np.longdouble) @ A.astype(np.longdouble).T).astype(np.longdouble)
    Kt = K[rows, rows].astype(np.longdouble)
    res_exact = Kt - exact                      # K - L L^T before the last block column is factored
    scale = float(np.max(np.abs(res_exact)))
    print(f"n = {n}, last diagonal tile, contraction depth {k * TB}: max |K - L L^T| = {scale:.3e}, max |L L^T| = {float(np.max(np.abs(exact))):.3e}")
    fp64 = A @ A.T
    print(f"  FP64 accumulation (BLAS dgemm):  max error / max |result| = {float(np.max(np.abs(fp64.astype(np.longdouble) - exact))) / scale:.2e}")
    for s in (6, 7, 8, 9, 10):
        C = ozaki_product(A, A, s)
        err = float(np.max(np.abs(C.astype(np.longdouble) - exact))) / scale
        print(f"  {s:2d} slices ({s * (s + 1) // 2:2d} int8 products), one scale per (row, 128-column tile):  max error / max |result| = {err:.2e}")
    # one scale per row from the Gram diagonal: sum_j L_ij^2 = K_ii, so |L_ij| <= sqrt(K_ii) < 2^e_i
    diag = np.diag(K)[rows]
    e_row = np.floor(np.log2(np.sqrt(diag))).astype(np.int64) + 1
    for s in (7, 8, 9, 10):
        C = ozaki_product_row_scale(A, A, s, e_row, e_row)
        err = float(np.max(np.abs(C.astype(np.longdouble) - exact))) / scale
        print(f"  {s:2d} slices ({s * (s + 1) // 2:2d} int8 products), one scale per row (int32 sums over all 15 tiles): max error / max |result| = {err:.2e}")


if __name__ == "__main__":
    main()
