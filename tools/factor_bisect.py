"""Developer tool: find the first tile in which the factor of a suspect build differs from a reference run.

  python tools/factor_bisect.py save  OUT.npz [n P runs]      run with the library selected by AGP_LIB; save lml/info and,
                                                               for every run, the factor of each particle whose result
                                                               differs from the first run (plus the first run's factors
                                                               of those particles are NOT available: use `ref` below)
  python tools/factor_bisect.py ref   OUT.npz particles...     save the factors of the given particles (reference library)
  python tools/factor_bisect.py diff  REF.npz BAD.npz          per differing particle: tiles (i,k) whose bits differ, in
                                                               dependency order (column-major), first one first
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def setup(n, P):
    import autogp_oracle as o
    import autogp.jl_b200 as agp
    from tools.dev_check import to_agp
    eng = agp.Engine(0)
    ts, xs = o.synthetic_series(n)
    parts = [o.synthetic_particle(p) for p in range(P)]
    return eng, [to_agp(nd) for nd, _ in parts], [nz for _, nz in parts], ts, xs


def main():
    mode = sys.argv[1]
    if mode == "save":
        out = sys.argv[2]
        n, P, runs = (int(a) for a in (sys.argv[3:6] + ["2048", "64", "12"][len(sys.argv) - 3:]))
        eng, nodes, noises, ts, xs = setup(n, P)
        first = None
        saved = {}
        for r in range(runs):
            lml, info = eng.lml_batch(nodes, noises, ts, xs)
            if first is None:
                first = lml.copy()
                continue
            same = (lml == first) | (np.isnan(lml) & np.isnan(first))
            for p in np.nonzero(~same)[0]:
                saved[f"run{r}_p{p}"] = eng.factor(int(p))
                print(f"run {r}: particle {p} differs (info {info[p]}, lml {lml[p]} vs {first[p]})", flush=True)
        np.savez_compressed(out, n=n, P=P, first=first, **saved)
        print("saved", list(saved))
    elif mode == "ref":
        out = sys.argv[2]
        ps = [int(a) for a in sys.argv[3:]]
        eng, nodes, noises, ts, xs = setup(2048, 64)
        lml, info = eng.lml_batch(nodes, noises, ts, xs)
        np.savez_compressed(out, lml=lml, **{f"p{p}": eng.factor(p) for p in ps})
    else:
        ref, bad = np.load(sys.argv[2]), np.load(sys.argv[3])
        for key in bad.files:
            if not key.startswith("run"):
                continue
            p = int(key.split("_p")[1])
            if f"p{p}" not in ref.files:
                print(key, "no reference factor saved for this particle")
                continue
            A, B = ref[f"p{p}"], bad[key]
            nt = A.shape[0] // 128
            diffs = []
            for k in range(nt):
                for i in range(k, nt):
                    a, b = A[i * 128:(i + 1) * 128, k * 128:(k + 1) * 128], B[i * 128:(i + 1) * 128, k * 128:(k + 1) * 128]
                    neq = ~((a == b) | (np.isnan(a) & np.isnan(b)))
                    if neq.any():
                        rr, cc = np.nonzero(neq)
                        diffs.append((i, k, int(neq.sum()), int(rr.min()), int(rr.max()), int(cc.min()), int(cc.max()),
                                      float(np.nanmax(np.abs(a - b)[neq]))))
            print(key, "differing tiles:", len(diffs))
            for d in diffs[:6]:
                print("   tile (i=%d,k=%d): %d entries differ, rows %d..%d cols %d..%d, max |diff| %.3e" % d)


if __name__ == "__main__":
    main()
