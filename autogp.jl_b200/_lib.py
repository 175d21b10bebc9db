"""ctypes binding of ``libagp_b200.so`` (C-ABI: include/agp_b200.h).

There is NO CPU fallback: if the shared library is missing, or a call is made without a CUDA
device, this module raises.  (The oracle under ``oracle/`` is test infrastructure and is never
imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGP_LIB") or os.path.join(_HERE, "libagp_b200.so")  # AGP_LIB: developer A/B builds

# every symbol include/agp_b200.h declares
EXPORTS = (
    "agp_create", "agp_destroy", "agp_reserve", "agp_last_error", "agp_version",
    "agp_gram", "agp_gram_device",
    "agp_lml_batch", "agp_lml_upload", "agp_lml_run", "agp_lml_fetch",
    "agp_lml_device_results", "agp_lml_set_prefix",
    "agp_stream", "agp_synchronize", "agp_launch_count", "agp_lml_time", "agp_lml_stage_times",
    "agp_queue_build", "agp_set_hybrid", "agp_dev_overlap_probe", "agp_hybrid_info", "agp_queue_build_hybrid", "agp_queue_build_general", "agp_queue_build_gram", "agp_gram_items", "agp_lml_trace",
    "agp_lml_run_append", "agp_predict_batch", "agp_predict_marginals_batch", "agp_predict_sum_batch", "agp_queue_build_marginals", "agp_lml_grad_batch", "agp_lml_grad_noise_batch", "agp_lml_copy_factor",
)

AGP_OK, AGP_ERR_ARG, AGP_ERR_PROGRAM, AGP_ERR_CUDA, AGP_ERR_NOMEM, AGP_ERR_STATE = 0, -1, -2, -3, -4, -5


class AgpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"agp error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C autogp.jl_b200/csrc`. There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    i32p, f64p, vp = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p
    lib.agp_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.agp_create.restype = C.c_int
    lib.agp_destroy.argtypes = [vp]
    lib.agp_destroy.restype = None
    lib.agp_last_error.argtypes = [vp]
    lib.agp_last_error.restype = C.c_char_p
    lib.agp_version.argtypes = []
    lib.agp_version.restype = C.c_char_p
    gram_args = [vp, i32p, i32p, C.c_int32, f64p, C.c_int32, f64p, C.c_int32, C.c_double, C.c_int32]
    lib.agp_gram.argtypes = gram_args + [f64p]
    lib.agp_gram.restype = C.c_int
    lib.agp_gram_device.argtypes = gram_args + [vp]
    lib.agp_gram_device.restype = C.c_int
    up_args = [vp, C.c_int32, i32p, i32p, i32p, i32p, f64p, f64p, f64p, f64p, C.c_int32]
    lib.agp_lml_batch.argtypes = up_args + [f64p, i32p]
    lib.agp_lml_batch.restype = C.c_int
    lib.agp_lml_upload.argtypes = up_args
    lib.agp_lml_upload.restype = C.c_int
    lib.agp_lml_run.argtypes = [vp]
    lib.agp_lml_run.restype = C.c_int
    lib.agp_lml_fetch.argtypes = [vp, f64p, i32p]
    lib.agp_lml_fetch.restype = C.c_int
    lib.agp_lml_copy_factor.argtypes = [vp, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    lib.agp_lml_copy_factor.restype = C.c_int
    lib.agp_lml_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.agp_lml_device_results.restype = C.c_int
    lib.agp_lml_set_prefix.argtypes = [vp, C.c_int32]
    lib.agp_lml_set_prefix.restype = C.c_int
    lib.agp_stream.argtypes = [vp]
    lib.agp_stream.restype = vp
    lib.agp_synchronize.argtypes = [vp]
    lib.agp_synchronize.restype = C.c_int
    lib.agp_launch_count.argtypes = [vp]
    lib.agp_launch_count.restype = C.c_int64
    lib.agp_lml_time.argtypes = [vp, C.c_int32, C.POINTER(C.c_float)]
    lib.agp_lml_time.restype = C.c_int
    lib.agp_lml_stage_times.argtypes = [vp, C.POINTER(C.c_float)]
    lib.agp_lml_stage_times.restype = C.c_int
    lib.agp_queue_build.argtypes = [C.c_int32, C.c_int32, C.c_int32, i32p, C.c_int64]
    lib.agp_queue_build.restype = C.c_int64
    lib.agp_queue_build_gram.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p, C.c_int64]
    lib.agp_queue_build_gram.restype = C.c_int64
    lib.agp_reserve.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.agp_reserve.restype = C.c_int
    lib.agp_set_hybrid.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    lib.agp_set_hybrid.restype = C.c_int
    lib.agp_hybrid_info.argtypes = [vp, i32p, i32p, C.POINTER(C.c_float)]
    lib.agp_hybrid_info.restype = C.c_int
    lib.agp_queue_build_hybrid.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p, C.c_int64, i32p, C.c_int32]
    lib.agp_queue_build_hybrid.restype = C.c_int64
    lib.agp_dev_overlap_probe.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
    lib.agp_dev_overlap_probe.restype = C.c_int
    lib.agp_gram_items.argtypes = [C.c_void_p, i32p, i32p]
    lib.agp_gram_items.restype = C.c_int
    lib.agp_queue_build_general.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p, C.c_int64]
    lib.agp_queue_build_general.restype = C.c_int64
    lib.agp_lml_run_append.argtypes = [vp]
    lib.agp_lml_run_append.restype = C.c_int
    lib.agp_predict_batch.argtypes = up_args + [f64p, C.c_int32, f64p, f64p, f64p, i32p]
    lib.agp_predict_batch.restype = C.c_int
    lib.agp_predict_marginals_batch.argtypes = up_args + [f64p, C.c_int32, f64p, f64p, f64p, i32p]
    lib.agp_predict_marginals_batch.restype = C.c_int
    lib.agp_queue_build_marginals.argtypes = [C.c_int32, C.c_int32, C.c_int32, i32p, C.c_int64]
    lib.agp_queue_build_marginals.restype = C.c_int64
    lib.agp_predict_sum_batch.argtypes = [vp, C.c_int32, C.c_int32] + up_args[2:] + [f64p, C.c_int32, f64p, f64p, f64p, i32p]
    lib.agp_predict_sum_batch.restype = C.c_int
    lib.agp_lml_grad_batch.argtypes = up_args + [f64p, f64p, f64p, i32p]
    lib.agp_lml_grad_batch.restype = C.c_int
    lib.agp_lml_grad_noise_batch.argtypes = up_args + [f64p, f64p, i32p]
    lib.agp_lml_grad_noise_batch.restype = C.c_int
    lib.agp_lml_trace.argtypes = [vp, C.POINTER(C.c_int64), C.c_int64]
    lib.agp_lml_trace.restype = C.c_int64
    _lib = lib
    return lib
