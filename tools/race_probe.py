"""Repeated-run probe of the persistent Cholesky kernel (developer tool, GPU box).

    AGP_LIB=gpurun_tmp/libagp_X.so python tools/race_probe.py n P runs [label]

Runs the same batch `runs` times and counts the runs whose LMLs are not bitwise the first run's.  For the first few
bad runs the factor of the differing particle is diffed tile by tile against the factor of a repeated (clean) run, which
gives the first wrong tile in dependency order.  A build with -DAGP_X_DEBUGOPS=1 exports agp_debug_read: its records say
what a fragment read from an operand stage held against the same words in global memory
(profiles/r02_race_experiments.txt).
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200 import _lib  # noqa: E402
from autogp.jl_b200.workloads import synthetic_particle, synthetic_series  # noqa: E402


def first_bad_tiles(A, B, limit=4):
    nt = A.shape[0] // 128
    out = []
    for k in range(nt):
        for i in range(k, nt):
            a, b = A[i * 128:(i + 1) * 128, k * 128:(k + 1) * 128], B[i * 128:(i + 1) * 128, k * 128:(k + 1) * 128]
            neq = ~((a == b) | (np.isnan(a) & np.isnan(b)))
            if neq.any():
                rr, cc = np.nonzero(neq)
                out.append(f"(i={i},k={k}) {int(neq.sum())} entries rows {rr.min()}..{rr.max()} cols {cc.min()}..{cc.max()}")
                if len(out) >= limit:
                    return out
    return out


def main():
    n, P, runs = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    label = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(os.environ.get("AGP_LIB", "product"))
    eng = agp.Engine(0)
    ts, xs = synthetic_series(n)
    parts = [synthetic_particle(p) for p in range(P)]
    nodes, noises = [nd for nd, _ in parts], [nz for _, nz in parts]
    lib = _lib.load()
    dbg = getattr(lib, "agp_debug_read", None)
    eng.upload(nodes, noises, ts, xs)
    eng.run()
    first, info0 = eng.fetch()
    first = first.copy()
    assert np.all(info0 == 0), f"first run has info != 0: {info0}"
    ms = eng.time_runs(5) if hasattr(eng, "time_runs") else float("nan")
    if dbg is not None:
        buf = (C.c_ulonglong * (16 * 1024))()
        dbg.argtypes = [C.POINTER(C.c_ulonglong), C.c_int, C.c_int]
        dbg.restype = C.c_int
        dbg(buf, 16 * 1024, 1)
    bad_runs, shown = 0, 0
    t0 = time.time()
    for r in range(runs):
        eng.run()
        lml, info = eng.fetch()
        same = (lml == first) | (np.isnan(lml) & np.isnan(first))
        if not same.all() or np.any(info != 0):
            bad_runs += 1
            if shown < 3:
                shown += 1
                ps = np.nonzero(~same | (info != 0))[0]
                p = int(ps[0])
                badf = eng.factor(p)
                for _ in range(3):
                    eng.run()
                    l2, i2 = eng.fetch()
                    if l2[p] == first[p]:
                        break
                goodf = eng.factor(p)
                print(f"  [{label}] run {r}: particles {ps.tolist()} info {info[ps].tolist()} first bad tiles of p={p}: {first_bad_tiles(goodf, badf)}", flush=True)
    dt = time.time() - t0
    line = f"[{label}] n={n} P={P}: {bad_runs} bad runs / {runs}  ({ms:.3f} ms per run device, {dt:.1f} s wall)"
    if dbg is not None:
        cnt = dbg(buf, 16 * 1024, 1)
        line += f"  operand mismatches recorded: {cnt}"
        w = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 16)
        f64 = w.view(np.float64)
        for j in range(min(cnt, 40)):
            d = w[j]
            idx, lo = int(d[0]) >> 32, int(d[0]) & 0xffffffff
            ch, ks, f, warp, lane = lo >> 16, (lo >> 12) & 15, (lo >> 8) & 15, (lo >> 5) & 7, lo & 31
            m = int(d[1])
            p_, k_, i_, h_, j0_, j1_ = m >> 48, (m >> 36) & 0xfff, (m >> 24) & 0xfff, (m >> 20) & 15, (m >> 10) & 0x3ff, m & 0x3ff
            sv, gv, sv2, gv2, gnext, gprev = (f64[j][c] for c in (2, 3, 4, 5, 6, 7))
            what = ("stage==next use" if sv == gnext else "stage==prev use" if sv == gprev else "stage==other")
            later = ("stage later==global" if sv2 == gv else "stage later==next use" if sv2 == gnext else "stage later differs")
            gl = "global stable" if gv == gv2 else "GLOBAL CHANGED"
            print(f"    item {idx} (p={p_} k={k_} i={i_} h={h_} [{j0_},{j1_})) chunk {ch}/{int(d[11])} ks {ks} frag {'B' if f >= 4 else 'A'}{f & 3} warp {warp} lane {lane} "
                  f"G0={int(d[10])} cta {int(d[13])} row {int(d[14])} col {int(d[15])}: {what}; {later}; {gl}  sv={sv:.6g} gv={gv:.6g} next={gnext:.6g} prev={gprev:.6g} t={int(d[12])}")
    print(line, flush=True)


if __name__ == "__main__":
    main()
