"""Accuracy of the interpreter's own FP64 exp / sin^2 / constant division (autogp.jl_b200/csrc/
agp_math.cuh), compiled for the host and compared with 50-digit mpmath.  The same source runs on
the device (lock-step over E entries); FMA is IEEE-exact on both sides, so the host result is the
device result."""
import ctypes as C
import os
import subprocess

import mpmath as mp
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("mathlib") / "libagp_math_host.so"
    src = os.path.join(ROOT, "tests", "host", "math_harness.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-o", str(out), src])
    return C.CDLL(str(out))


def _call(fn, *arrs):
    n = len(arrs[0])
    y = np.zeros(n)
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    fn(*[ptr(np.ascontiguousarray(a, dtype=np.float64)) for a in arrs], ptr(y), C.c_long(n))
    return y


def _ulp_err(got, exact_mp):
    errs = []
    for g, e in zip(got.tolist(), exact_mp):
        ref = float(e)
        ulp = np.spacing(abs(ref)) if ref != 0 else 5e-324
        errs.append(abs(mp.mpf(g) - e) / mp.mpf(float(ulp)))
    return float(max(errs))


def test_exp_within_one_ulp(lib):
    mp.mp.dps = 50
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-699.9, 699.9, 6000), rng.uniform(-40, 0.5, 6000), -np.logspace(-18, 2, 2000),
                        np.array([0.0, -0.0, 1e-300, -1e-300, 0.5 * np.log(2), -0.5 * np.log(2), 699.99, -699.99])])
    x = x[: len(x) // 4 * 4]
    y = _call(lib.agp_host_exp, x)
    assert _ulp_err(y, [mp.exp(mp.mpf(v)) for v in x.tolist()]) <= 1.0


def test_exp_fallback_range_matches_libm(lib):
    x = np.array([-700.0, -708.5, -745.0, -746.0, -1e4, 700.0, 709.7, 710.0, np.inf, -np.inf, 5e-324, -720.25])
    y = _call(lib.agp_host_exp, x)
    with np.errstate(over="ignore", under="ignore"):
        assert np.array_equal(y, np.exp(x))
    assert np.isnan(_call(lib.agp_host_exp, np.array([np.nan, 0.0, 0.0, 0.0]))[0])


def test_sin_squared_accuracy(lib):
    mp.mp.dps = 60
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(0, 10, 4000), rng.uniform(0, 1e5, 4000), np.logspace(-12, 0, 500),
                        np.arange(1, 200) * (np.pi / 2), np.arange(1, 200) * (np.pi / 4), np.array([0.0, 99999.5])])
    x = x[: len(x) // 4 * 4]
    y = _call(lib.agp_host_sin2, x)
    # sin r to ~1.3 ulp (half an ulp from the reduced argument, like libdevice) -> sin^2 within 3.5 ulp,
    # including the tiny values next to multiples of pi/2 (relative accuracy of the reduction)
    assert _ulp_err(y, [mp.sin(mp.mpf(v)) ** 2 for v in x.tolist()]) <= 3.5
    # what the reference computes, round(sin x)^2 rounded, is itself up to ~1.8 ulp off
    ref = np.sin(x) ** 2
    assert np.all(np.abs(y - ref) <= 5 * np.spacing(ref))


def test_sin_fallback_range(lib):
    x = np.array([1.0e5 + 0.5, 3.0e7, 1e300, 2.5e5])
    y = _call(lib.agp_host_sin2, x)
    assert np.all(np.abs(y - np.sin(x) ** 2) <= 4 * np.spacing(np.sin(x) ** 2))
    assert np.isnan(_call(lib.agp_host_sin2, np.array([np.inf, 0.0, 0.0, 0.0]))[0])


def test_constant_division_is_correctly_rounded(lib):
    rng = np.random.default_rng(2)
    a = np.exp(rng.uniform(-20, 20, 50000))
    x = -0.5 * rng.uniform(-30, 30, 50000) ** 2
    x[:100] = 0.0
    x[100:200] = -0.0
    x[200:300] = rng.uniform(-1, 1, 100) * 1e-200   # outside the fast window: generic division
    a[300:400] = 1.0 + np.arange(100) * np.finfo(float).eps  # divisors just above 1
    y = _call(lib.agp_host_div, x, a)
    assert np.array_equal(y, x / a)
