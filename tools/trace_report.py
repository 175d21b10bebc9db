"""Per-phase / per-slot utilisation of one traced run of the persistent kernel (GPU box)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200 import _lib  # noqa: E402
from autogp.jl_b200.workloads import synthetic_batch, synthetic_series  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    order = int(os.environ.get("AGP_ORDER", "3"))
    eng = agp.Engine(0)
    ts, xs = synthetic_series(n)
    nodes, noises = synthetic_batch(P)
    eng.upload(nodes, noises, ts, xs)
    for _ in range(3):
        eng.run()
    eng.synchronize()
    ms = eng.time_runs(10) / 10
    st = eng.trace()
    nt = -(-n // 128)
    lib = _lib.load()
    items = np.zeros((st.shape[0], 8), dtype=np.int32)
    fused, lead = eng.gram_items()
    hyb, width, _ = eng.hybrid_info()
    seg = None
    if hyb:   # the segments of the hybrid schedule: stamps of all launches in one array, items in agp_queue_build_hybrid order
        fused = False
        seg = np.zeros(-(-nt // width) + 1, dtype=np.int32)
        i32p = C.POINTER(C.c_int32)
        assert lib.agp_queue_build_hybrid(P, nt, width, 0, 0, items.ctypes.data_as(i32p), st.shape[0], seg.ctypes.data_as(i32p), len(seg)) == st.shape[0]
        print(f"hybrid schedule, super-column width {width}: segments start at items {seg.tolist()}")
        for s_ in range(len(seg) - 1):
            a, b = seg[s_], seg[s_ + 1]
            w = st[a:b]
            busy = (w[:, 5] - w[:, 0]).sum() * 1e-3
            span_s = (w[:, 5].max() - w[:, 0].min()) * 1e-3
            slots_s = len(np.unique(w[:, 7]))
            print(f"  segment {s_}: items {b - a:6d}  span {span_s:8.0f} us  busy {busy / 1e3:8.1f} ms*cta  occupancy {busy / (slots_s * span_s):.3f}")
    elif fused:
        assert lib.agp_queue_build_gram(P, nt, order, lead, items.ctypes.data_as(C.POINTER(C.c_int32)), st.shape[0]) == st.shape[0]
    else:
        assert lib.agp_queue_build(P, nt, order, items.ctypes.data_as(C.POINTER(C.c_int32)), st.shape[0]) == st.shape[0]
    typ = items[:, 0] & 0xFF
    t0 = st[:, 0].min()
    t_end = st[:, 5].max()
    span = (t_end - t0) * 1e-3
    print(f"n={n} P={P} order={order}: untraced {ms*1e3:.0f} us/run, traced span {span:.0f} us, items {len(st)}")
    names = {0: "DIAG", 1: "POTF2", 2: "PANEL"}
    slots = len(np.unique(st[:, 7]))
    busy_total = 0.0
    if fused:
        m = typ == 3
        s = st[m].astype(np.float64) * 1e-3
        tot = s[:, 5] - s[:, 0]
        busy_total += tot.sum()
        print(f"GRAM   items {m.sum():6d}  total {tot.sum()/1e3:9.1f} ms*cta  mean {tot.mean():7.1f} us  (lead {lead}); window [{(s[:,0].min()-t0*1e-3):8.0f}, {(s[:,5].max()-t0*1e-3):8.0f}] us")
        # waits of the first readers for their Gram flag: inside wait1 of DIAG / PANEL items below
    for t in (0, 1, 2):
        m = typ == t
        s = st[m].astype(np.float64) * 1e-3
        tot = s[:, 5] - s[:, 0]
        busy_total += tot.sum()
        line = f"{names[t]:6s} items {m.sum():6d}  total {tot.sum()/1e3:9.1f} ms*cta  mean {tot.mean():7.1f} us"
        if t == 1:
            w = s[:, 1] - s[:, 0]
            line += f" | wait {w.mean():6.1f}  work {(s[:,5]-s[:,1]).mean():6.1f} = load {(s[:,2]-s[:,1]).mean():5.1f} + panels {(s[:,3]-s[:,2]).mean():5.1f} + tail {(s[:,5]-s[:,3]).mean():5.1f}"
        else:
            w1 = s[:, 1] - s[:, 0]
            mm = s[:, 2] - s[:, 1]
            if t == 2:
                fin = (items[m, 0] & (1 << 9)) == 0   # final panels only carry the solve stamps
                s, w1, mm = s[fin], w1[fin], mm[fin]
                gr = s[:, 3] - s[:, 2]
                w2 = s[:, 4] - s[:, 3]
                tr = s[:, 5] - s[:, 4]
                line += f" | wait1 {w1.mean():6.1f} mma {mm.mean():6.1f} xst {gr.mean():6.1f} waitF {w2.mean():6.1f} trsm {tr.mean():6.1f}"
                line += f" | sums(ms*cta): wait1 {w1.sum()/1e3:.1f} mma {mm.sum()/1e3:.1f} xst {gr.sum()/1e3:.1f} waitF {w2.sum()/1e3:.1f} trsm {tr.sum()/1e3:.1f}"
            else:
                rest = s[:, 5] - s[:, 2]
                kk = items[m, 2] > 0
                line += f" | (k>0) wait1 {w1[kk].mean():6.1f} mma {mm[kk].mean():6.1f} store {rest[kk].mean():6.1f}"
        print(line)
        if t == 1 and (st[m][:, 4] != 0).any():  # clocks thread 0 spent in the three phases of the item's steps
            c = st[m][:, 4]
            p1, p2, p3 = ((c >> 42) & 0x1fffff) * 16, ((c >> 21) & 0x1fffff) * 16, (c & 0x1fffff) * 16
            print(f"       potf2 clocks per item: phase 1 (diagonal block, warp 0) {np.mean(p1):8.0f}  phase 2 (row substitution) {np.mean(p2):8.0f}  phase 3 (DMMA update) {np.mean(p3):8.0f}")
    print(f"slots {slots}  busy {busy_total/1e3:.1f} ms*cta  capacity {slots*span/1e3:.1f} ms*cta  occupancy {busy_total/(slots*span):.3f}")
    # per-block-column view of the panels
    print("per block column k: panel items mean us (wait1, mma, gram, waitF, trsm) and the wall-clock window of the column")
    for k in range(nt):
        m = (typ == 2) & (items[:, 2] == k) & ((items[:, 0] & (1 << 9)) == 0)   # final panel items
        if not m.any():
            continue
        s = st[m].astype(np.float64) * 1e-3
        print(f"  k={k:2d} n={m.sum():5d} wait1 {np.mean(s[:,1]-s[:,0]):6.1f} mma {np.mean(s[:,2]-s[:,1]):6.1f} xst {np.mean(s[:,3]-s[:,2]):6.1f} "
              f"waitF {np.mean(s[:,4]-s[:,3]):6.1f} trsm {np.mean(s[:,5]-s[:,4]):6.1f}  window [{(s[:,0].min()-t0*1e-3):8.0f}, {(s[:,5].max()-t0*1e-3):8.0f}] us")
    m = typ == 1
    s = st[m].astype(np.float64) * 1e-3
    for k in range(nt):
        mk = items[m, 2] == k
        print(f"  potf2 k={k:2d} wait {np.mean(s[mk,1]-s[mk,0]):6.1f} work {np.mean(s[mk,5]-s[mk,1]):6.1f} window [{(s[mk,0].min()-t0*1e-3):8.0f}, {(s[mk,5].max()-t0*1e-3):8.0f}]")
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"trace_n{n}_P{P}_o{order}.npz"), items=items, stamps=st)


if __name__ == "__main__":
    main()
