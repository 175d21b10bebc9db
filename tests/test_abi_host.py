"""CPU suite, part 2: the C-ABI library and the host-side mirror (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import autogp_oracle as o
import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "agp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from autogp.jl_b200 import _lib

    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    assert set(declared) == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None, name
    assert b"sm_100a" in lib.agp_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    import autogp.jl_b200 as agp
    from autogp.jl_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.AgpError):
        agp.Engine(0)
    with pytest.raises(_lib.AgpError):
        agp.compute_cov_matrix_vectorized(agp.Constant(1.0), 0.1, np.linspace(0, 1, 4))


def test_product_never_imports_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "autogp.jl_b200")
    pat = re.compile(r"^\s*(import|from|#include)\b.*oracle|CDLL\(.*oracle|liboracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f


def test_encode_program_matches_oracle_encoder_and_unroll_order():
    import autogp.jl_b200 as agp

    for k in H.fixture_kernels():
        a = agp.encode_program(H.to_agp(k))
        b = o.encode_program(k)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    # postfix order of GP.unroll (src/GP.jl:111-113): left, right, node
    t = agp.Plus(agp.Times(agp.SquaredExponential(1.0), agp.Periodic(1.0, 2.0)), agp.Linear(0.5))
    names = [type(n).__name__ for n in agp.unroll(t)]
    assert names == ["SquaredExponential", "Periodic", "Times", "Linear", "Plus"]
    assert agp.size(t) == 5 and agp.depth(t) == 3
    # GPConfig integer codes (src/GP.jl:1101-1108)
    assert agp.encode_program(t)[0].tolist() == [3, 5, 7, 2, 6]


def test_operator_overloads_and_defaults():
    import autogp.jl_b200 as agp

    a, b = agp.Linear(0.5), agp.Periodic(2.0, 1.0)
    assert (a + b) == agp.Plus(a, b) and (a * b) == agp.Times(a, b)
    assert agp.Linear(0.5).bias == 1.0 and agp.Linear(0.5).amplitude == 1.0  # GP.jl:189
    assert agp.SquaredExponential(2.0).amplitude == 1.0
    with pytest.raises(AssertionError):
        agp.GammaExponential(1.0, 2.5)  # GP.jl:274


def test_transforms_match_oracle():
    import autogp.jl_b200 as agp

    assert agp.JITTER == o.JITTER
    for f in ("noise", "period", "gamma", "lengthscale"):
        for z in (-2.0, 0.0, 0.7):
            v = agp.transform_param(f, z)
            assert v == o.transform_param(f, z)
            assert agp.untransform_param(f, v) == pytest.approx(z, abs=1e-12)


def test_program_compiler_rejects_malformed_programs():
    """Error behaviour of the boundary: bad programs are rejected before any CUDA work."""
    from autogp.jl_b200 import _lib

    lib = _lib.load()
    # agp_create fails without a GPU, but argument checks on a null handle must not crash
    assert lib.agp_lml_run(None) == _lib.AGP_ERR_ARG
    assert lib.agp_synchronize(None) == _lib.AGP_ERR_ARG
    assert lib.agp_launch_count(None) == 0
    assert lib.agp_last_error(None) == b"null handle"
    h = C.c_void_p()
    assert lib.agp_create(-1, C.byref(h)) != 0 and not h.value


def test_shard_range_partitions():
    from autogp.jl_b200 import smc

    for P in (1, 7, 64, 512):
        for w in (1, 2, 3, 8):
            spans = [smc.shard_range(P, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == P
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_weight_consumers_match_oracle():
    from autogp.jl_b200 import smc

    rng = np.random.default_rng(3)
    lw = rng.normal(size=32) * 5
    assert smc.effective_sample_size(lw) == pytest.approx(o.effective_sample_size(o.normalize_weights(lw)[1]))
    assert np.allclose(smc.compute_particle_weights(lw), np.exp(o.normalize_weights(lw)[1]))
    assert np.array_equal(smc.resample_indices(lw, 11), smc.resample_indices(lw, 11))


def _replay_queue(buf, P, nt, nt_total, first_row, first_col=None, expect_trailing=True, gram_items=False, segments=None, width=0, aug=False,
                  slice_items=False):
    """Sequential replay of a work queue with the kernel's own wait rules (agp_fused.cu).  Every
    counter wait must already be satisfied by EARLIER items (deadlock-freedom of in-order popping),
    every tile an item reads must be final, and the items must tile every contraction exactly.
    first_col(i): first block column with a non-zero tile in tile row i (0 except for the rows of
    the inverse schedule)."""
    DIAG, POTF2, PANEL, GRAM, SLICE, PARTIAL, YINIT = 0, 1, 2, 3, 4, 1 << 9, 1 << 10
    sliced = set()
    first_col = first_col or (lambda i: 0)
    nts = nt_total  # the builders lay the counters out for nt_stride = nt_total
    n_tiles = nts * (nts + 1) // 2
    counters = np.zeros(32 + 3 * P * nts + P + 2 * P * n_tiles, dtype=np.int64)
    gflag = lambda p, i, k, h: 32 + 3 * P * nts + P + ((p * n_tiles + i * (i + 1) // 2 + k) * 2 + h)
    n_gram, gram_done = 0, set()
    rowdone = lambda p, i: 32 + p * nts + i
    diagu = lambda p, k: 32 + P * nts + p * nts + k
    ppre = lambda p, i: 32 + 2 * P * nts + p * nts + i
    fdone = lambda p: 32 + 3 * P * nts + p
    final = {}      # (p, row, col) -> finished halves of the solved panel L[row][col]
    for p in range(P):  # continuation: rows above first_row are final, first_row diagonal tiles factored
        for i in range(first_row):
            counters[rowdone(p, i)] = 2 * i
            for j in range(i):
                final[(p, i, j)] = 2
        counters[fdone(p)] = first_row
    covered, n_diag, stored = {}, {}, set()
    seg_first = {int(a): s_ for s_, a in enumerate(segments[:-1])} if segments is not None else {}
    seg_base = {}   # (p, i, k, h) -> block columns [0, base) were contracted by the int8 update kernel
    for idx_, (x, p, k, i, f4, f5, flag, need) in enumerate(buf.tolist()):
        if idx_ in seg_first and seg_first[idx_] > 0:
            # hybrid schedule: a new launch starts here.  Between the launches the int8 kernel brings every lower tile of
            # the super-column's block columns to coverage c0 — everything it reads must be final, nothing it writes
            # may have been touched, and all earlier items are complete (kernel boundary).
            c0_ = seg_first[idx_] * width
            if gram_items and seg_first[idx_] == 1:   # the row scales are read off the Gram diagonal after the first segment
                assert all((pp, c, c, hh) in gram_done for pp in range(P) for c in range(nt) for hh in (0, 1))
            for pp in range(P):
                assert counters[fdone(pp)] == c0_
                for ii in range(c0_, nt + (c0_ if aug else 0)):   # (aug: and the appended rows nt + a, a < c0, over [a, c0))
                    assert all(final.get((pp, ii, j), 0) == 2 for j in range(first_col(ii), c0_)), (pp, ii)
                    if slice_items:   # what the int8 launch reads: planes from the first non-zero tile on, and one zero tile before it
                        assert all((pp, ii, j, hh) in sliced for j in range(max(first_col(ii) - 1, 0), c0_) for hh in (0, 1)), (pp, ii)
                    for kk in range(c0_, min(nt, c0_ + width)):
                        if kk <= ii:
                            for hh in (0, 1):
                                assert (pp, ii, kk, hh) not in covered
                                assert not gram_items or (pp, ii, kk, hh) in gram_done   # the update subtracts from the Gram tile
                                covered[(pp, ii, kk, hh)] = c0_
                                seg_base[(pp, ii, kk, hh)] = c0_
        t, h, partial, yinit = x & 0xFF, (x >> 8) & 1, bool(x & PARTIAL), bool(x & YINIT)
        j0, j1, need_k, need_i = f4 & 0xFFFF, f4 >> 16, f5 & 0xFFFF, f5 >> 16
        assert 0 <= p < P and 0 <= k < nt_total
        if t == GRAM:   # Gram unit of tile half (i, k, h): bumps that half's flag, exactly once, before its reader
            assert gram_items and k <= i < nt and (p, i, k, h) not in gram_done
            # the two units of a diagonal tile share the flag of its first half (a DIAG item reads blocks of both halves)
            assert flag == gflag(p, i, k, 0 if i == k else h) and counters[flag] == (len({(p, i, k, 0), (p, i, k, 1)} & gram_done) if i == k else 0)
            assert (p, i, k, h) not in covered and (i != k or (p, i, k, 1 - h) not in covered)
            gram_done.add((p, i, k, h))
            counters[flag] += 1
            n_gram += 1
            continue
        if t == SLICE:   # digit planes of a finished tile half: its tile row has finished the panels up to block column k
            assert slice_items and k < nt and k < i < nt_total and (p, i, k, h) not in sliced and flag == -1
            assert counters[rowdone(p, i)] >= need_i and need_i == max(0, 2 * (k - first_col(i) + 1))
            assert final.get((p, i, k), 0) == 2 or k < first_col(i)
            sliced.add((p, i, k, h))
            continue
        if t == POTF2:
            assert k < nt and counters[diagu(p, k)] >= need and need == n_diag.get((p, k), 0), (p, k, need)
            if segments is not None and k > 0:   # the diagonal tile is complete: DIAG items or the int8 update
                assert all(covered.get((p, k, k, hh)) == k for hh in (0, 1)), (p, k)
            assert counters[fdone(p)] == k          # block columns are factored in order
            counters[fdone(p)] += 1
            continue
        assert (t == DIAG and i == k) or (t == PANEL and k < i < nt_total)
        assert i >= first_row and 0 <= j0 <= j1 <= min(k, nt)
        start = max(first_col(i), first_col(k)) if j1 > 0 else 0
        assert covered.get((p, i, k, h), min(start, j1)) == j0, "contraction ranges must be contiguous"
        covered[(p, i, k, h)] = j1
        # the kernel's waits
        assert counters[rowdone(p, k)] >= need_k and (t == DIAG or counters[rowdone(p, i)] >= need_i)
        first_touch = gram_items and j0 == 0 and start == 0
        if first_touch:   # the accumulators start from minus the Gram tile half: its unit must be done
            if t == DIAG:
                assert flag == gflag(p, k, k, 0) and need == 2 and counters[flag] == 2
            else:
                assert flag == gflag(p, i, k, h) and need == 1 and counters[flag] == 1
        elif flag >= 0:
            assert flag == (diagu(p, k) if t == DIAG else ppre(p, i)) and counters[flag] >= need
        if not first_touch and (p, i, k, h) not in seg_base:
            assert (flag >= 0) == (j0 > start)
        # what it reads is final
        for j in range(j0, j1):
            assert final.get((p, k, j), 0) == 2 and final.get((p, i, j), 0) == 2, (t, p, k, i, j)
        if t == DIAG:
            counters[diagu(p, k)] += 1
            n_diag[(p, k)] = n_diag.get((p, k), 0) + 1
            if partial and k >= nt:
                stored.add((p, i, k, h))
        elif partial:
            counters[ppre(p, i)] += 1
            if k >= nt:
                stored.add((p, i, k, h))
        else:
            assert k < nt and j1 == k
            assert yinit == (k == first_col(i)), "the first panel of a tile row starts the forward-solve entry"
            assert counters[fdone(p)] >= k + 1      # L_kk ready for the triangular solve
            counters[rowdone(p, i)] += 1
            final[(p, i, k)] = final.get((p, i, k), 0) + 1
    assert n_gram == (P * nt * (nt + 1) if gram_items else 0)   # every lower tile half exactly once
    for p in range(P):
        assert counters[fdone(p)] == nt
        for i in range(first_row, nt_total):
            assert counters[rowdone(p, i)] == 2 * (min(i, nt) - first_col(i))
            for k in range(first_col(i), i + 1):
                for h in (0, 1):
                    if k < nt:
                        assert covered[(p, i, k, h)] == k         # finished: contraction over all of [first, k)
                    elif not expect_trailing:
                        assert (p, i, k, h) not in covered
                    elif first_col(k) <= first_col(i):
                        assert covered[(p, i, k, h)] == nt and (p, i, k, h) in stored   # trailing (Schur complement) tile
    return sliced


@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("P,nt", [(1, 1), (3, 2), (2, 5), (5, 16), (2, 23)])
def test_work_queue_replay_never_waits_for_a_later_item(P, nt, order):
    """Deadlock-freedom of the persistent kernel (agp_fused.cu): CTAs pop items in queue order and a
    CTA only ever spins on counters, so replaying the queue sequentially with the kernel's own wait
    rules must find every wait already satisfied."""
    from autogp.jl_b200 import _lib

    lib = _lib.load()
    n_items = lib.agp_queue_build(P, nt, order, None, 0)
    buf = np.zeros((n_items, 8), dtype=np.int32)
    assert lib.agp_queue_build(P, nt, order, buf.ctypes.data_as(C.POINTER(C.c_int32)), n_items) == n_items
    _replay_queue(buf, P, nt, nt, 0)


@pytest.mark.parametrize("slices", [0, 1])
@pytest.mark.parametrize("P,nt,width", [(1, 2, 1), (3, 5, 2), (2, 8, 4), (3, 16, 4), (2, 16, 3), (1, 9, 2)])
def test_hybrid_schedule_of_the_gradient_calls_replay(P, nt, width, slices):
    """The identity-augmented batch on the hybrid schedule: the panels of the appended rows nt + a (rows of L^{-T}) in
    the segments, their contraction over [a, c0) on the int8 path; no FP64 item touches the trailing block (the lauum
    pass is one int8 launch)."""
    from autogp.jl_b200 import _lib

    lib = _lib.load()
    i32p = C.POINTER(C.c_int32)
    n_items = lib.agp_queue_build_hybrid(P, nt, width, 0, 1 + 2 * slices, None, 0, None, 0)
    buf = np.zeros((n_items, 8), dtype=np.int32)
    seg = np.zeros((nt + width - 1) // width + 1, dtype=np.int32)
    assert lib.agp_queue_build_hybrid(P, nt, width, 0, 1 + 2 * slices, buf.ctypes.data_as(i32p), n_items, seg.ctypes.data_as(i32p), len(seg)) == n_items
    assert seg[0] == 0 and seg[-1] == n_items
    sliced = _replay_queue(buf, P, nt, 2 * nt, 0, first_col=lambda i: i - nt if i >= nt else 0, expect_trailing=False, segments=seg.tolist(), width=width,
                           aug=True, slice_items=bool(slices))
    if slices:   # the lauum launch after the last segment reads every appended row from one tile before its first non-zero one
        assert all((p, nt + a, j, hh) in sliced for p in range(P) for a in range(nt) for j in range(max(a - 1, 0), nt) for hh in (0, 1))


@pytest.mark.parametrize("gram_lead", [0, 1, 8, 296, -1])
@pytest.mark.parametrize("P,nt,width", [(1, 2, 1), (3, 5, 2), (2, 8, 4), (5, 16, 4), (2, 16, 2), (2, 23, 4), (1, 9, 3)])
def test_hybrid_schedule_replay(P, nt, width, gram_lead):
    """The super-column schedule of the hybrid factorisation (agp_queue_build_hybrid): every segment is a launch of
    its own, the int8 update between two launches covers block columns [0, c0) of the next super-column's tiles, and the
    items inside a segment obey the same wait rules as the single-launch schedule."""
    from autogp.jl_b200 import _lib

    lib = _lib.load()
    i32p = C.POINTER(C.c_int32)
    slices = 2 if gram_lead < 0 else 0   # (-1: the default schedule — Gram launch in front, SLICE items in the segments)
    gram_lead = max(gram_lead, 0)
    n_items = lib.agp_queue_build_hybrid(P, nt, width, gram_lead, slices, None, 0, None, 0)
    buf = np.zeros((n_items, 8), dtype=np.int32)
    seg = np.zeros((nt + width - 1) // width + 1, dtype=np.int32)
    assert lib.agp_queue_build_hybrid(P, nt, width, gram_lead, slices, buf.ctypes.data_as(i32p), n_items, seg.ctypes.data_as(i32p), len(seg)) == n_items
    assert seg[0] == 0 and seg[-1] == n_items and np.all(np.diff(seg) > 0)
    _replay_queue(buf, P, nt, nt, 0, segments=seg.tolist(), width=width, gram_items=gram_lead > 0, slice_items=slices > 0)


@pytest.mark.parametrize("order", [0, 2, 3])
@pytest.mark.parametrize("P,nt", [(1, 1), (3, 2), (2, 5), (5, 16), (2, 23)])
def test_work_queue_with_gram_items_replay(P, nt, order):
    """The plain schedule with the Gram units as queue items (ITEM_GRAM, csrc/agp_chol_gram.cu): every lower tile half is
    evaluated exactly once, before the first item that reads it, and that item waits for the unit's flag; stripped of the
    GRAM items and of those waits the queue is the plain one."""
    from autogp.jl_b200 import _lib

    lib = _lib.load()
    n_items = lib.agp_queue_build(P, nt, 300 + order, None, 0)
    buf = np.zeros((n_items, 8), dtype=np.int32)
    assert lib.agp_queue_build(P, nt, 300 + order, buf.ctypes.data_as(C.POINTER(C.c_int32)), n_items) == n_items
    _replay_queue(buf, P, nt, nt, 0, gram_items=True)
    n_plain = lib.agp_queue_build(P, nt, order, None, 0)
    plain = np.zeros((n_plain, 8), dtype=np.int32)
    lib.agp_queue_build(P, nt, order, plain.ctypes.data_as(C.POINTER(C.c_int32)), n_plain)
    rest = buf[(buf[:, 0] & 0xFF) != 3].copy()
    assert n_items - n_plain == P * nt * (nt + 1) and rest.shape == plain.shape
    first = (plain[:, 6] < 0) & ((plain[:, 0] & 0xFF) != 1) & ((plain[:, 4] & 0xFFFF) == 0)
    assert np.array_equal(rest[~first], plain[~first]) and np.array_equal(rest[first][:, :6], plain[first][:, :6])


@pytest.mark.parametrize("P,nt,nt_total,first_row", [(2, 4, 4, 2), (3, 6, 6, 5), (1, 3, 3, 0), (2, 4, 6, 0), (1, 1, 2, 0),
                                                     (2, 0, 2, 0), (3, 5, 6, 0), (2, 16, 19, 0), (1, 9, 12, 0)])
def test_continuation_queues_replay(P, nt, nt_total, first_row):
    """Block-append (tile rows >= first_row only) and predictive (extra tile rows below the factored
    block) schedules obey the same rules."""
    from autogp.jl_b200 import _lib

    lib = _lib.load()
    n_items = lib.agp_queue_build_general(P, nt, nt_total, first_row, None, 0)
    assert n_items > 0
    buf = np.zeros((n_items, 8), dtype=np.int32)
    lib.agp_queue_build_general(P, nt, nt_total, first_row, buf.ctypes.data_as(C.POINTER(C.c_int32)), n_items)
    _replay_queue(buf, P, nt, nt_total, first_row)


@pytest.mark.parametrize("P,nt,order", [(1, 1, 2), (2, 3, 0), (2, 5, 2), (1, 9, 2)])
def test_inverse_schedule_replay(P, nt, order):
    """Identity-augmented factorisation of agp_lml_grad_batch: tile row nt + a solves to block row a of
    L^{-T} (non-zero from block column a on), the trailing tiles receive -K^{-1}."""
    from autogp.jl_b200 import _lib

    lib = _lib.load()
    n_items = lib.agp_queue_build(P, nt, 100 + order, None, 0)
    buf = np.zeros((n_items, 8), dtype=np.int32)
    lib.agp_queue_build(P, nt, 100 + order, buf.ctypes.data_as(C.POINTER(C.c_int32)), n_items)
    _replay_queue(buf, P, nt, 2 * nt, 0, first_col=lambda i: i - nt if i >= nt else 0)


@pytest.mark.parametrize("P,nt,order", [(1, 1, 3), (2, 4, 3), (1, 9, 2)])
def test_trtri_only_schedule_replay(P, nt, order):
    """agp_lml_grad_noise_batch: the identity-augmented schedule without its lauum pass — the same items as the full
    inverse schedule up to the first trailing store-only item, and nothing after."""
    from autogp.jl_b200 import _lib

    lib = _lib.load()
    n_items = lib.agp_queue_build(P, nt, 200 + order, None, 0)
    buf = np.zeros((n_items, 8), dtype=np.int32)
    lib.agp_queue_build(P, nt, 200 + order, buf.ctypes.data_as(C.POINTER(C.c_int32)), n_items)
    _replay_queue(buf, P, nt, 2 * nt, 0, first_col=lambda i: i - nt if i >= nt else 0, expect_trailing=False)
    n_full = lib.agp_queue_build(P, nt, 100 + order, None, 0)
    full = np.zeros((n_full, 8), dtype=np.int32)
    lib.agp_queue_build(P, nt, 100 + order, full.ctypes.data_as(C.POINTER(C.c_int32)), n_full)
    assert n_full - n_items == P * nt * (nt + 1)          # two half items per trailing lower tile
    assert np.array_equal(full[:n_items], buf)
    assert not np.any(buf[:, 2] >= nt)                    # no item factors or updates a trailing block column


def test_split_kernel_sop_follows_the_reference_examples():
    """GP.split_kernel_sop (src/GP.jl:603-655): the docstring examples (:589-600) and the ChangePoint rule."""
    import autogp.jl_b200 as agp

    l, p, c = agp.Linear(1.0), agp.Periodic(1.0, 1.0), agp.Constant(1.0)
    zero = agp.Constant(0.0)
    split = agp.split_kernel_sop
    assert split(l, agp.Linear) == (l, zero)
    assert split(l, agp.Periodic) == (zero, l)
    assert split(l * p + l * c, agp.Periodic) == (l * p, l * c)
    assert split(p * p, agp.Periodic) == (p * p, zero)
    a, b = split((l + p) * (l + p), agp.Periodic)
    assert a == (p * p + p * l) + l * p and b == l * l          # "p*l+p*p+l*p, l*l" up to the order of the addends
    cp = agp.ChangePoint(l * p, c, 0.4, 0.01)
    assert split(cp, agp.Periodic) == (agp.ChangePoint(l * p, zero, 0.4, 0.01), agp.ChangePoint(zero, c, 0.4, 0.01))
    assert split(agp.ChangePoint(l, c, 0.4, 0.01), agp.Periodic) == (zero, agp.ChangePoint(l, c, 0.4, 0.01))
    # the two sides add up to the kernel (sum-of-products expansion), checked on the oracle's Gram matrices
    import autogp_oracle as o
    from helpers import from_agp

    ts = np.linspace(0, 1, 17)
    tree = agp.Plus(agp.Times(agp.Plus(agp.Linear(0.3, 0.2, 0.7), agp.Periodic(0.6, 0.25, 1.1)), agp.SquaredExponential(0.4, 0.9)),
                    agp.ChangePoint(agp.Periodic(0.5, 0.3, 0.8), agp.GammaExponential(0.3, 1.2, 0.5), 0.5, 0.05))
    ka, kb = split(tree, agp.Periodic)
    total = o.eval_cov(from_agp(ka), ts) + o.eval_cov(from_agp(kb), ts)
    np.testing.assert_allclose(total, o.eval_cov(from_agp(tree), ts), rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("P,nt,nt_total", [(2, 4, 6), (1, 16, 19), (3, 1, 2)])
def test_marginals_queue_replay(P, nt, nt_total):
    """agp_predict_marginals_batch: the predictive schedule with only the diagonal tiles of the trailing block."""
    from autogp.jl_b200 import _lib

    lib = _lib.load()
    n_items = lib.agp_queue_build_marginals(P, nt, nt_total, None, 0)
    buf = np.zeros((n_items, 8), dtype=np.int32)
    lib.agp_queue_build_marginals(P, nt, nt_total, buf.ctypes.data_as(C.POINTER(C.c_int32)), n_items)
    n_full = lib.agp_queue_build_general(P, nt, nt_total, 0, None, 0)
    extra = nt_total - nt
    assert n_full - n_items == P * extra * (extra - 1)            # two half items per off-diagonal trailing tile
    # same replay rules; off-diagonal trailing tiles must be absent
    trailing = [(int(r[2]), int(r[3])) for r in buf if r[2] >= nt]
    assert trailing and all(k == i for k, i in trailing)
    full = np.zeros((n_full, 8), dtype=np.int32)
    lib.agp_queue_build_general(P, nt, nt_total, 0, full.ctypes.data_as(C.POINTER(C.c_int32)), n_full)
    keep = np.array([not (r[2] >= nt and r[2] != r[3]) for r in full])
    assert np.array_equal(full[keep], buf)                        # the full predictive queue minus those tiles


def test_marginal_quantiles_match_the_normal_quantile_function():
    """Distributions.quantile(::MvNormal, p) of the reference (src/GP.jl:1006-1012) reads mean and sqrt(diag(cov)) only."""
    import scipy.stats

    import autogp.jl_b200 as agp

    mean = np.array([0.0, 1.5, -2.0, 3.0])
    var = np.array([1.0, 0.25, 4.0, 0.0])
    qs = [0.025, 0.5, 0.975]
    got = agp.marginal_quantiles(mean, var, qs)
    want = np.stack([scipy.stats.norm.ppf(q, loc=mean, scale=np.sqrt(var)) for q in qs], axis=1)
    want[3, :] = 3.0                                      # a degenerate marginal is a point mass
    assert got.shape == (4, 3)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    assert np.all(agp.marginal_quantiles(mean[:3], var[:3], [0.0, 1.0]) == np.array([[-np.inf, np.inf]] * 3))
