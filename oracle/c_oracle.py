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
