"""ctypes loader for oracle/oracle_c.c — TEST INFRASTRUCTURE ONLY (see autogp_oracle.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_c.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "oracle_c.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        lib = C.CDLL(_SO)
        i32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        lib.oracle_gram.argtypes = [i32p, i32p, C.c_int32, f64p, f64p, C.c_int32, C.c_double, C.c_int, f64p]
        lib.oracle_gram.restype = None
        lib.oracle_lml.argtypes = [i32p, i32p, C.c_int32, f64p, f64p, f64p, C.c_int32, C.c_double, f64p]
        lib.oracle_lml.restype = C.c_int
        lib.oracle_gram_upper_cols.argtypes = [i32p, i32p, C.c_int32, f64p, f64p, C.c_int32, C.c_double, C.c_int32, C.c_int32, f64p]
        lib.oracle_gram_upper_cols.restype = None
        _lib = lib
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def gram(program, ts, noise, form=0):
    ops, offs, params = program
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    params = np.ascontiguousarray(params if len(params) else np.zeros(1), dtype=np.float64)
    n = len(ts)
    K = np.empty((n, n), order="F")
    load().oracle_gram(_p(ops, C.c_int32), _p(offs, C.c_int32), len(ops), _p(params, C.c_double),
                       _p(ts, C.c_double), n, float(noise), int(form), _p(K, C.c_double))
    return K


def lml(program, ts, xs, noise):
    ops, offs, params = program
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    params = np.ascontiguousarray(params if len(params) else np.zeros(1), dtype=np.float64)
    out = C.c_double()
    info = load().oracle_lml(_p(ops, C.c_int32), _p(offs, C.c_int32), len(ops), _p(params, C.c_double),
                             _p(ts, C.c_double), _p(xs, C.c_double), len(ts), float(noise), C.byref(out))
    return out.value, info


def lml_cpu_best(program, ts, xs, noise, threads=1):
    """BASELINE.md §2 "CPU-best" context number (not the reference's path): fused Gram (upper triangle, no temporaries,
    column ranges of equal area over `threads` host threads) + LAPACK dpotrf / dtrtrs from SciPy's OpenBLAS."""
    from concurrent.futures import ThreadPoolExecutor

    from scipy.linalg import cholesky, solve_triangular

    ops, offs, params = program
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    params = np.ascontiguousarray(params if len(params) else np.zeros(1), dtype=np.float64)
    n = len(ts)
    K = np.zeros((n, n), order="F")
    lib = load()
    chunks = 4 * max(threads, 1)
    cuts = sorted({int(round(n * np.sqrt(c / chunks))) for c in range(chunks + 1)})  # equal triangle area per chunk

    def fill(rng):
        lib.oracle_gram_upper_cols(_p(ops, C.c_int32), _p(offs, C.c_int32), len(ops), _p(params, C.c_double), _p(ts, C.c_double), n,
                                   float(noise), rng[0], rng[1], _p(K, C.c_double))

    ranges = list(zip(cuts[:-1], cuts[1:]))
    if threads > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(fill, ranges))
    else:
        for rng in ranges:
            fill(rng)
    U = cholesky(K, lower=False, overwrite_a=True, check_finite=False)
    z = solve_triangular(U, xs, trans="T", lower=False, check_finite=False)
    return -0.5 * (n * np.log(2.0 * np.pi) + 2.0 * float(np.sum(np.log(np.diag(U))))) - 0.5 * float(z @ z)
