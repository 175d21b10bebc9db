"""Host-side mirror of the reference's GP operator API for the hot path (src/GP.jl).

Same names, argument meaning and error behaviour as the Julia functions they stand in for;
all arithmetic runs in ``libagp_b200.so`` on the GPU (no CPU fallback):

  ===========================================  =========================================
  reference (src/GP.jl)                         here
  ===========================================  =========================================
  ``WhiteNoise … ChangePoint`` structs :131-479  dataclasses below (same field order)
  ``unroll(node)``                    :111-113  :func:`unroll` (postfix; the wire order)
  ``size`` / ``depth``                :93-104   :func:`size` / :func:`depth`
  ``eval_cov(node, ts)``              :61       :func:`eval_cov`
  ``compute_cov_matrix_vectorized``   :666-668  :func:`compute_cov_matrix_vectorized`
  ``compute_cov_matrix``              :674-684  :func:`compute_cov_matrix`
  ``GPConfig`` node codes             :1101-1108  ``OP_*`` constants
  ===========================================  =========================================
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib

OP_CONSTANT, OP_LINEAR, OP_SQUARED_EXPONENTIAL, OP_GAMMA_EXPONENTIAL, OP_PERIODIC = 1, 2, 3, 4, 5
OP_PLUS, OP_TIMES, OP_CHANGEPOINT, OP_WHITE_NOISE = 6, 7, 8, 9
FORM_VECTORIZED, FORM_SCALAR = 0, 1


class Node:
    """``abstract type Node`` (src/GP.jl:39)."""

    def __add__(self, other: "Node") -> "Plus":  # Base.:+ (GP.jl:379)
        return Plus(self, other)

    def __mul__(self, other: "Node") -> "Times":  # Base.:* (GP.jl:380)
        return Times(self, other)


class LeafNode(Node):
    pass


class BinaryOpNode(Node):
    pass


@dataclass(frozen=True)
class WhiteNoise(LeafNode):
    value: float


@dataclass(frozen=True)
class Constant(LeafNode):
    value: float


@dataclass(frozen=True)
class Linear(LeafNode):
    intercept: float
    bias: float = 1.0
    amplitude: float = 1.0


@dataclass(frozen=True)
class SquaredExponential(LeafNode):
    lengthscale: float
    amplitude: float = 1.0


@dataclass(frozen=True)
class GammaExponential(LeafNode):
    lengthscale: float
    gamma: float
    amplitude: float = 1.0

    def __post_init__(self):
        # `@assert (0 < gamma <= 2)` in the Julia constructor (GP.jl:274)
        if not (0 < self.gamma <= 2):
            raise AssertionError("0 < gamma <= 2")


@dataclass(frozen=True)
class Periodic(LeafNode):
    lengthscale: float
    period: float
    amplitude: float = 1.0


@dataclass(frozen=True)
class Plus(BinaryOpNode):
    left: Node
    right: Node


@dataclass(frozen=True)
class Times(BinaryOpNode):
    left: Node
    right: Node


@dataclass(frozen=True)
class ChangePoint(BinaryOpNode):
    left: Node
    right: Node
    location: float
    scale: float


def size(node: Node) -> int:
    return 1 if isinstance(node, LeafNode) else 1 + size(node.left) + size(node.right)


def depth(node: Node) -> int:
    return 1 if isinstance(node, LeafNode) else 1 + max(depth(node.left), depth(node.right))


def unroll(node: Node) -> List[Node]:
    """Flat postfix list of all sub-kernels (src/GP.jl:111-113)."""
    out: List[Node] = []
    stack = [(node, False)]
    while stack:  # iterative: trees are unbounded a priori (max_depth = -1)
        nd, seen = stack.pop()
        if isinstance(nd, LeafNode) or seen:
            out.append(nd)
        else:
            stack.append((nd, True))
            stack.append((nd.right, False))
            stack.append((nd.left, False))
    return out


_LEAF_FIELDS = {
    WhiteNoise: (OP_WHITE_NOISE, ("value",)),
    Constant: (OP_CONSTANT, ("value",)),
    Linear: (OP_LINEAR, ("intercept", "bias", "amplitude")),
    SquaredExponential: (OP_SQUARED_EXPONENTIAL, ("lengthscale", "amplitude")),
    GammaExponential: (OP_GAMMA_EXPONENTIAL, ("lengthscale", "gamma", "amplitude")),
    Periodic: (OP_PERIODIC, ("lengthscale", "period", "amplitude")),
}


def _encode_into(node: Node, ops: list, offs: list, params: list, p0: int) -> None:
    """Appends the postfix program of `node` (``unroll`` order, src/GP.jl:111-113); parameter offsets count from `p0`.
    Iterative: trees are unbounded a priori (max_depth = -1)."""
    stack = [(node, False)]
    while stack:
        nd, seen = stack.pop()
        t = type(nd)
        leaf = _LEAF_FIELDS.get(t)
        if leaf is None and not seen:
            if not isinstance(nd, BinaryOpNode):
                raise TypeError(f"not a covariance kernel node: {nd!r}")
            stack.append((nd, True))
            stack.append((nd.right, False))
            stack.append((nd.left, False))
            continue
        offs.append(len(params) - p0)
        if leaf is not None:
            ops.append(leaf[0])
            for f in leaf[1]:
                params.append(float(getattr(nd, f)))
        elif t is Plus:
            ops.append(OP_PLUS)
        elif t is Times:
            ops.append(OP_TIMES)
        elif t is ChangePoint:
            ops.append(OP_CHANGEPOINT)
            params.append(float(nd.location))
            params.append(float(nd.scale))
        else:
            raise TypeError(f"not a covariance kernel node: {nd!r}")


def encode_program(node: Node) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Kernel tree -> wire format (ops, param_off, params) of include/agp_b200.h."""
    ops, offs, params = [], [], []
    _encode_into(node, ops, offs, params, 0)
    return (np.asarray(ops, dtype=np.int32), np.asarray(offs, dtype=np.int32),
            np.asarray(params, dtype=np.float64))


def _i32p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Engine:
    """One ``agp_handle``: a CUDA stream plus its device workspaces on one GPU."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.agp_create(int(device), C.byref(h))
        if rc != 0:
            raise _lib.AgpError(rc, f"agp_create(device={device}) failed (is a CUDA device visible?)")
        self._h = h
        self.device = int(device)
        self._P = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.agp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise _lib.AgpError(rc, self._lib.agp_last_error(self._h).decode())

    # ---- site 1 -------------------------------------------------------------------------
    def reserve(self, max_n: int, max_batch: int, *, max_pred: int = 0, gradient: bool = False) -> None:
        """Size the large device buffers once (``agp_reserve``): a data-annealing run otherwise re-allocates them every
        time the series grows."""
        self._check(self._lib.agp_reserve(self._h, int(max_n), int(max_pred), int(max_batch), 1 if gradient else 0))

    def gram(self, node: Node, noise: float, ts, form: int = FORM_VECTORIZED) -> np.ndarray:
        ops, offs, params = encode_program(node)
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        n = ts.shape[0]
        K = np.empty((n, n), dtype=np.float64, order="F")
        self._check(self._lib.agp_gram(self._h, _i32p(ops), _i32p(offs), len(ops), _f64p(params), len(params),
                                       _f64p(ts), n, float(noise), int(form), _f64p(K)))
        return K

    def gram_device(self, node: Node, noise: float, ts, out_ptr: int, form: int = FORM_VECTORIZED) -> None:
        ops, offs, params = encode_program(node)
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        self._check(self._lib.agp_gram_device(self._h, _i32p(ops), _i32p(offs), len(ops), _f64p(params), len(params),
                                              _f64p(ts), ts.shape[0], float(noise), int(form), C.c_void_p(out_ptr)))

    # ---- site 2 -------------------------------------------------------------------------
    @staticmethod
    def pack_batch(nodes: Sequence[Node], noises: Sequence[float]):
        """All particles' programs in the wire format of ``agp_lml_upload`` (one flat pass: this runs once per MH step
        of every particle, inside the end-to-end call)."""
        ops: List[int] = []
        offs: List[int] = []
        params: List[float] = []
        prog_len: List[int] = []
        n_params: List[int] = []
        for node in nodes:
            o0, p0 = len(ops), len(params)
            _encode_into(node, ops, offs, params, p0)
            prog_len.append(len(ops) - o0)
            n_params.append(len(params) - p0)
        noise = np.ascontiguousarray(noises, dtype=np.float64)
        if noise.shape[0] != len(prog_len):
            raise ValueError("one noise value per particle")
        return (np.asarray(prog_len, dtype=np.int32), np.asarray(ops, dtype=np.int32), np.asarray(offs, dtype=np.int32),
                np.asarray(n_params, dtype=np.int32), np.asarray(params, dtype=np.float64), noise)

    def lml_batch(self, nodes: Sequence[Node], noises: Sequence[float], ts, xs) -> Tuple[np.ndarray, np.ndarray]:
        """Host buffers in, host results out: one ``agp_lml_batch`` call."""
        packed = self.pack_batch(nodes, noises)
        return self.lml_batch_packed(packed, ts, xs)

    def lml_batch_packed(self, packed, ts, xs) -> Tuple[np.ndarray, np.ndarray]:
        prog_len, ops, offs, n_params, params, noise = packed
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        if ts.shape != xs.shape:
            raise ValueError("ts and xs must have equal length")
        P = len(prog_len)
        lml = np.empty(P, dtype=np.float64)
        info = np.empty(P, dtype=np.int32)
        self._check(self._lib.agp_lml_batch(self._h, P, _i32p(prog_len), _i32p(ops), _i32p(offs), _i32p(n_params),
                                            _f64p(params), _f64p(noise), _f64p(ts), _f64p(xs), ts.shape[0],
                                            _f64p(lml), _i32p(info)))
        self._P = P
        return lml, info

    def upload(self, nodes: Sequence[Node], noises: Sequence[float], ts, xs) -> None:
        self.upload_packed(self.pack_batch(nodes, noises), ts, xs)

    def upload_packed(self, packed, ts, xs) -> None:
        prog_len, ops, offs, n_params, params, noise = packed
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        if ts.shape != xs.shape:
            raise ValueError("ts and xs must have equal length")
        self._check(self._lib.agp_lml_upload(self._h, len(prog_len), _i32p(prog_len), _i32p(ops), _i32p(offs),
                                             _i32p(n_params), _f64p(params), _f64p(noise), _f64p(ts), _f64p(xs),
                                             ts.shape[0]))
        self._P = len(prog_len)

    def set_prefix(self, n_prefix: int) -> None:
        self._check(self._lib.agp_lml_set_prefix(self._h, int(n_prefix)))

    def run(self) -> None:
        self._check(self._lib.agp_lml_run(self._h))

    def run_append(self) -> None:
        """Continue the resident factorisation after ``set_prefix`` grew the data (block-append:
        only the new tile rows are computed).  Raises AgpError(AGP_ERR_STATE) when no clean factor
        is resident — the caller then falls back to :meth:`run`."""
        self._check(self._lib.agp_lml_run_append(self._h))

    def lml_grad_batch(self, nodes: Sequence[Node], noises: Sequence[float], ts, xs):
        """LML and its gradient: returns (lml[P], grads, grad_noise[P], info[P]) where ``grads[p]`` is the
        gradient with respect to particle p's parameters in ``encode_program(node)[2]`` order."""
        prog_len, ops, offs, n_params, params, noise = self.pack_batch(nodes, noises)
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        if ts.shape != xs.shape:
            raise ValueError("ts and xs must have equal length")
        P = len(prog_len)
        lml = np.empty(P, dtype=np.float64)
        gparams = np.empty(max(int(n_params.sum()), 1), dtype=np.float64)
        gnoise = np.empty(P, dtype=np.float64)
        info = np.empty(P, dtype=np.int32)
        self._check(self._lib.agp_lml_grad_batch(self._h, P, _i32p(prog_len), _i32p(ops), _i32p(offs), _i32p(n_params),
                                                 _f64p(params), _f64p(noise), _f64p(ts), _f64p(xs), ts.shape[0],
                                                 _f64p(lml), _f64p(gparams), _f64p(gnoise), _i32p(info)))
        self._P = P
        cuts = np.concatenate([[0], np.cumsum(n_params)])
        grads = [gparams[cuts[p]:cuts[p + 1]].copy() for p in range(P)]
        return lml, grads, gnoise, info

    def lml_grad_noise_batch(self, nodes: Sequence[Node], noises: Sequence[float], ts, xs):
        """LML and dLML/dnoise only (the noise move of the reference's HMC): returns (lml[P], grad_noise[P], info[P]).
        Factorisation + triangular inverse, no K^{-1}, no kernel-tree walk: ~0.6 of ``lml_grad_batch``."""
        prog_len, ops, offs, n_params, params, noise = self.pack_batch(nodes, noises)
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        if ts.shape != xs.shape:
            raise ValueError("ts and xs must have equal length")
        P = len(prog_len)
        lml = np.empty(P, dtype=np.float64)
        gnoise = np.empty(P, dtype=np.float64)
        info = np.empty(P, dtype=np.int32)
        self._check(self._lib.agp_lml_grad_noise_batch(self._h, P, _i32p(prog_len), _i32p(ops), _i32p(offs), _i32p(n_params),
                                                       _f64p(params), _f64p(noise), _f64p(ts), _f64p(xs), ts.shape[0],
                                                       _f64p(lml), _f64p(gnoise), _i32p(info)))
        self._P = P
        return lml, gnoise, info

    # ---- site 3 -------------------------------------------------------------------------
    def predict_batch(self, nodes: Sequence[Node], noises: Sequence[float], ts, xs, ts_pred,
                      noise_pred: Optional[Sequence[float]] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Conditional MVN of every particle at ``ts_pred`` (``Distributions.MvNormal(node, noise, ts, xs,
        ts_pred; noise_pred)``, src/GP.jl:731-758): returns (mean[P, m], cov[P, m, m], info[P])."""
        prog_len, ops, offs, n_params, params, noise = self.pack_batch(nodes, noises)
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        tp = np.ascontiguousarray(ts_pred, dtype=np.float64)
        if ts.shape != xs.shape:
            raise ValueError("ts and xs must have equal length")
        P, m = len(prog_len), tp.shape[0]
        npred = None if noise_pred is None else np.ascontiguousarray(noise_pred, dtype=np.float64)
        if npred is not None and npred.shape[0] != P:
            raise ValueError("one noise_pred value per particle")
        mean = np.empty((P, m), dtype=np.float64)
        cov = np.empty((P, m, m), dtype=np.float64)
        info = np.empty(P, dtype=np.int32)
        self._check(self._lib.agp_predict_batch(self._h, P, _i32p(prog_len), _i32p(ops), _i32p(offs), _i32p(n_params),
                                                _f64p(params), _f64p(noise), _f64p(ts), _f64p(xs), ts.shape[0],
                                                _f64p(tp), m, None if npred is None else _f64p(npred),
                                                _f64p(mean), _f64p(cov), _i32p(info)))
        self._P = P
        return mean, cov, info

    def predict_marginals_batch(self, nodes: Sequence[Node], noises: Sequence[float], ts, xs, ts_pred,
                                noise_pred: Optional[Sequence[float]] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Mean and marginal variance of X(ts_pred) | X(ts) = xs for every particle — what ``predict``'s quantiles read
        (src/api.jl:633-699 -> ``Distributions.quantile``, src/GP.jl:1006-1012: mean and sqrt(diag(cov)) only): returns
        (mean[P, m], var[P, m], info[P]); no m x m covariance is formed or copied."""
        prog_len, ops, offs, n_params, params, noise = self.pack_batch(nodes, noises)
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        tp = np.ascontiguousarray(ts_pred, dtype=np.float64)
        if ts.shape != xs.shape:
            raise ValueError("ts and xs must have equal length")
        P, m = len(prog_len), tp.shape[0]
        npred = None if noise_pred is None else np.ascontiguousarray(noise_pred, dtype=np.float64)
        if npred is not None and npred.shape[0] != P:
            raise ValueError("one noise_pred value per particle")
        mean = np.empty((P, m), dtype=np.float64)
        var = np.empty((P, m), dtype=np.float64)
        info = np.empty(P, dtype=np.int32)
        self._check(self._lib.agp_predict_marginals_batch(self._h, P, _i32p(prog_len), _i32p(ops), _i32p(offs), _i32p(n_params),
                                                          _f64p(params), _f64p(noise), _f64p(ts), _f64p(xs), ts.shape[0],
                                                          _f64p(tp), m, None if npred is None else _f64p(npred),
                                                          _f64p(mean), _f64p(var), _i32p(info)))
        self._P = P
        return mean, var, info

    def predict_sum_batch(self, summands: Sequence[Sequence[Node]], noises: Sequence[float], ts, xs, ts_pred,
                          noise_pred: Optional[Sequence[float]] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """``infer_gp_sum(nodes, noise, ts, xs, ts_pred; noise_pred)`` (src/GP.jl:904-993) for every particle:
        ``summands[p]`` are the M additive components of particle p's kernel (same M for all).  Returns
        (mean[P, d], cov[P, d, d], info[P]), d = (M + 1) m, over [F_1(T*); ...; F_M(T*); X(T*)] — without the JITTER
        the reference adds when it wraps the result in an MvNormal (:986)."""
        P = len(summands)
        M = len(summands[0]) if P else 1
        if M < 1 or any(len(sm) != M for sm in summands):
            raise ValueError("every particle needs the same number (>= 1) of summands")
        flat = [nd for sm in summands for nd in sm]
        prog_len, ops, offs, n_params, params, _ = self.pack_batch(flat, [0.0] * len(flat))
        noise = np.ascontiguousarray(noises, dtype=np.float64)
        if noise.shape[0] != P:
            raise ValueError("one noise value per particle")
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        tp = np.ascontiguousarray(ts_pred, dtype=np.float64)
        if ts.shape != xs.shape:
            raise ValueError("ts and xs must have equal length")
        m = tp.shape[0]
        d = (M + 1) * m
        npred = None if noise_pred is None else np.ascontiguousarray(noise_pred, dtype=np.float64)
        if npred is not None and npred.shape[0] != P:
            raise ValueError("one noise_pred value per particle")
        mean = np.empty((P, d), dtype=np.float64)
        cov = np.empty((P, d, d), dtype=np.float64)
        info = np.empty(P, dtype=np.int32)
        self._check(self._lib.agp_predict_sum_batch(self._h, P, M, _i32p(prog_len), _i32p(ops), _i32p(offs), _i32p(n_params),
                                                    _f64p(params), _f64p(noise), _f64p(ts), _f64p(xs), ts.shape[0],
                                                    _f64p(tp), m, None if npred is None else _f64p(npred),
                                                    _f64p(mean), _f64p(cov), _i32p(info)))
        self._P = P
        return mean, cov, info

    def fetch(self) -> Tuple[np.ndarray, np.ndarray]:
        lml = np.empty(self._P, dtype=np.float64)
        info = np.empty(self._P, dtype=np.int32)
        self._check(self._lib.agp_lml_fetch(self._h, _f64p(lml), _i32p(info)))
        return lml, info

    def factor(self, particle: int) -> np.ndarray:
        """Lower Cholesky factor of one particle of the resident batch ([ld, ld], rows >= n are padding); its
        transpose is the upper factor of cholesky(Symmetric(K)) that the reference's MvNormal holds."""
        ld = C.c_int32(0)
        self._check(self._lib.agp_lml_copy_factor(self._h, particle, None, C.byref(ld)))
        out = np.empty((ld.value, ld.value), dtype=np.float64)
        self._check(self._lib.agp_lml_copy_factor(self._h, particle, _f64p(out), C.byref(ld)))
        return np.tril(out)

    def device_results(self) -> Tuple[int, int]:
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self._lib.agp_lml_device_results(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def time_runs(self, reps: int) -> float:
        ms = C.c_float()
        self._check(self._lib.agp_lml_time(self._h, int(reps), C.byref(ms)))
        return float(ms.value)

    def stage_times(self) -> Tuple[float, float, float]:
        arr = (C.c_float * 3)()
        self._check(self._lib.agp_lml_stage_times(self._h, arr))
        return float(arr[0]), float(arr[1]), float(arr[2])

    def trace(self) -> np.ndarray:
        """Diagnostics: one traced run of the resident batch -> (items[n,4], stamps[n,8]) in queue
        order (see agp_lml_trace / agp_queue_build in include/agp_b200.h)."""
        n_items = int(self._lib.agp_lml_trace(self._h, None, 0))
        if n_items < 0:
            self._check(n_items)
        stamps = np.zeros((n_items, 8), dtype=np.int64)
        rc = int(self._lib.agp_lml_trace(self._h, stamps.ctypes.data_as(C.POINTER(C.c_int64)), n_items))
        if rc < 0:
            self._check(rc)
        return stamps

    def gram_items(self) -> Tuple[bool, int]:
        """(Gram units run as items of the persistent kernel's queue, their lead in items): agp_gram_items."""
        a, b = C.c_int32(), C.c_int32()
        self._check(self._lib.agp_gram_items(self._h, C.byref(a), C.byref(b)))
        return bool(a.value), int(b.value)

    def set_hybrid(self, mode: int = -1, width: int = 0, min_nt: int = 12) -> None:
        """Hybrid factorisation of plain LML runs (agp_set_hybrid): the long contractions as exact int8 digit-plane
        products on tcgen05, super-columns of `width` block columns (0 = by size); mode -1 = from `min_nt` block columns on, 0 = never,
        1 = whenever the batch has more than `width` block columns."""
        self._check(self._lib.agp_set_hybrid(self._h, int(mode), int(width), int(min_nt)))

    def hybrid_info(self) -> Tuple[bool, int, Tuple[float, float, float, float]]:
        """(the resident batch takes the hybrid schedule, super-column width, stage times of the last stage_times()
        call on a hybrid run: Gram + row scales, persistent-kernel segments, int8 updates, digit planes) in ms."""
        a, w = C.c_int32(), C.c_int32()
        ms = (C.c_float * 4)()
        self._check(self._lib.agp_hybrid_info(self._h, C.byref(a), C.byref(w), ms))
        return a.value > 0, int(w.value), tuple(float(x) for x in ms)

    def synchronize(self) -> None:
        self._check(self._lib.agp_synchronize(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.agp_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.agp_launch_count(self._h))


_default_engines = {}
_default_lock = threading.Lock()


def default_engine(device: int = 0) -> Engine:
    """Per-(thread, device) engine, like one handle per Julia thread."""
    key = (threading.get_ident(), int(device))
    with _default_lock:
        eng = _default_engines.get(key)
        if eng is None:
            eng = _default_engines[key] = Engine(device)
        return eng


def eval_cov(node: Node, ts, *, engine: Optional[Engine] = None) -> np.ndarray:
    """``eval_cov(node, ts::Vector{Float64})`` (src/GP.jl:61): n×n covariance matrix."""
    return (engine or default_engine()).gram(node, 0.0, ts, FORM_VECTORIZED)


def compute_cov_matrix_vectorized(node: Node, noise: float, ts, *, engine: Optional[Engine] = None) -> np.ndarray:
    """``GP.compute_cov_matrix_vectorized(node, noise, ts)`` (src/GP.jl:666-668)."""
    return (engine or default_engine()).gram(node, noise, ts, FORM_VECTORIZED)


def predictive_mvn(node: Node, noise: float, ts, xs, ts_pred, *, noise_pred: Optional[float] = None,
                   engine: Optional[Engine] = None) -> Tuple[np.ndarray, np.ndarray]:
    """``Distributions.MvNormal(node, noise, ts, xs, ts_pred; noise_pred)`` (src/GP.jl:731-758): the mean
    vector and covariance matrix of X(ts_pred) | X(ts) = xs.  Raises ``PosDefException`` when the
    training covariance is not positive definite (as the reference's ``\\`` would)."""
    eng = engine or default_engine()
    mean, cov, info = eng.predict_batch([node], [noise], ts, xs, ts_pred, None if noise_pred is None else [noise_pred])
    if info[0] != 0:
        from .model import PosDefException
        raise PosDefException(int(info[0]), 0)
    return mean[0], cov[0]


def marginal_quantiles(mean, var, quantiles: Sequence[float]) -> np.ndarray:
    """``Distributions.quantile(dist::MvNormal, p)`` (src/GP.jl:1006-1012): quantiles of the MARGINALS,
    ``quantile(Normal(mu_i, sqrt(cov_ii)), q)`` — [m, len(quantiles)] from the mean and marginal variance that
    ``Engine.predict_marginals_batch`` returns for one particle."""
    from statistics import NormalDist

    mean = np.asarray(mean, dtype=np.float64)
    std = np.sqrt(np.asarray(var, dtype=np.float64))
    z = np.array([-np.inf if q == 0.0 else np.inf if q == 1.0 else NormalDist().inv_cdf(float(q)) for q in quantiles])
    with np.errstate(invalid="ignore"):
        out = mean[:, None] + std[:, None] * z[None, :]
    out[std == 0.0, :] = mean[std == 0.0, None]   # a degenerate marginal is a point mass (0 * inf otherwise)
    return out


def split_kernel_sop(node: Node, leaf_type: type) -> Tuple[Node, Node]:
    """``GP.split_kernel_sop(node, T)`` (src/GP.jl:603-655): read the kernel as a sum of products and return
    ``(k_T, k_nT)`` — the addends with a factor of base-kernel type ``leaf_type`` and the addends without one, with
    ``Constant(0)`` standing for an empty side.  ``predict_mvn_sum`` (src/api.jl) feeds the pair to ``infer_gp_sum``."""
    def merge_plus(a, b):                                   # merge_split_operand(::Plus, ...), :642-648
        if a is None:
            return b
        return a if b is None else Plus(a, b)

    def helper(nd):
        if isinstance(nd, LeafNode):                        # :614-615
            return (nd, None) if type(nd) is leaf_type else (None, nd)
        la, lb = helper(nd.left)
        ra, rb = helper(nd.right)
        if isinstance(nd, Times):                           # :617-631: distribute the product over the two sides
            mult = lambda a, b: None if a is None or b is None else Times(a, b)
            l_sop = merge_plus(merge_plus(mult(la, ra), mult(la, rb)), mult(lb, ra))
            return l_sop, mult(lb, rb)
        if isinstance(nd, ChangePoint):                     # :650-656: an empty side becomes Constant(0)
            def merge_cp(a, b):
                if a is None and b is None:
                    return None
                return ChangePoint(Constant(0.0) if a is None else a, Constant(0.0) if b is None else b, nd.location, nd.scale)
            return merge_cp(la, ra), merge_cp(lb, rb)
        return merge_plus(la, ra), merge_plus(lb, rb)       # Plus, :633-639

    a, b = helper(node)
    return (Constant(0.0) if a is None else a), (Constant(0.0) if b is None else b)   # :604-608


def infer_gp_sum(nodes: Sequence[Node], noise: float, ts, xs, ts_pred, *, noise_pred: Optional[float] = None,
                 engine: Optional[Engine] = None):
    """``GP.infer_gp_sum(nodes, noise, ts, xs, ts_pred; noise_pred)`` (src/GP.jl:904-993): posterior over the latent
    summands F_i(ts_pred) and the observable X(ts_pred) of  X = sum_i F_i + eps.  Returns
    ``(mean, cov, indexes)`` where ``cov`` includes the reference's ``JITTER * I`` (:986; the GP module's own
    ``JITTER = 1e-8``, src/GP.jl:760 — not Model.jl's 1e-5) and ``indexes`` =
    ``{"F": [range per summand], "X": range}`` (0-based; :989-992)."""
    from .model import PosDefException

    JITTER = 1e-8  # src/GP.jl:760
    eng = engine or default_engine()
    mean, cov, info = eng.predict_sum_batch([list(nodes)], [noise], ts, xs, ts_pred, None if noise_pred is None else [noise_pred])
    if info[0] != 0:
        raise PosDefException(int(info[0]), 0)
    m, M = len(np.atleast_1d(ts_pred)), len(nodes)
    cov = cov[0] + JITTER * np.eye(cov.shape[1])
    return mean[0], cov, {"F": [range(i * m, (i + 1) * m) for i in range(M)], "X": range(M * m, (M + 1) * m)}


def compute_cov_matrix(node: Node, noise: float, ts, *, engine: Optional[Engine] = None) -> np.ndarray:
    """``GP.compute_cov_matrix(node, noise, ts)`` (src/GP.jl:674-684), scalar-form arithmetic."""
    return (engine or default_engine()).gram(node, noise, ts, FORM_SCALAR)
