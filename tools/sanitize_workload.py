"""One small invocation of every entry point of the library, for tools/sanitize.sh (compute-sanitizer)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200.workloads import synthetic_particle, synthetic_series  # noqa: E402


def main():
    eng = agp.Engine(0)
    n, P = 300, 5   # three block columns, ragged last tile
    ts, xs = synthetic_series(n)
    parts = [synthetic_particle(p, "se*per+lin" if p % 2 == 0 else "ge+per*lin") for p in range(P)]
    nodes, noises = [nd for nd, _ in parts], [nz for _, nz in parts]
    K = eng.gram(nodes[0], noises[0], ts[:100])
    lml, info = eng.lml_batch(nodes, noises, ts, xs)
    assert np.all(info == 0) and np.all(np.isfinite(lml)) and np.all(np.isfinite(K)) and eng.gram_items()[0] is True
    # seven block columns: the Gram fill as its own launch, look-ahead items, the right-looking tail (n = 300 above: three
    # block columns, the Gram units as queue items)
    ts7, xs7 = synthetic_series(800)
    lml7, info7 = eng.lml_batch(nodes[:3], noises[:3], ts7, xs7)
    assert np.all(info7 == 0) and np.all(np.isfinite(lml7)) and eng.gram_items()[0] is False
    eng.upload(nodes, noises, ts, xs)
    eng.set_prefix(140)
    eng.run()
    eng.fetch()
    eng.set_prefix(n)
    eng.run_append()
    app, _ = eng.fetch()
    assert np.array_equal(app, lml)
    _, grads, gnoise, ginfo = eng.lml_grad_batch(nodes[:2], noises[:2], ts[:200], xs[:200])
    _, gn2, ninfo = eng.lml_grad_noise_batch(nodes[:2], noises[:2], ts[:200], xs[:200])
    assert np.all(ginfo == 0) and np.all(ninfo == 0)
    mean, cov, pinfo = eng.predict_batch(nodes[:2], noises[:2], ts[:200], xs[:200], ts[200:260])
    m2, var, _ = eng.predict_marginals_batch(nodes[:2], noises[:2], ts[:200], xs[:200], ts[200:260])
    assert np.all(pinfo == 0) and np.all(np.isfinite(cov)) and np.all(np.isfinite(var))
    summands = [[parts[0][0].left, parts[0][0].right]]
    smean, scov, sinfo = eng.predict_sum_batch(summands, noises[:1], ts[:200], xs[:200], ts[200:230])
    assert sinfo[0] == 0
    # the hybrid schedule, forced on at small sizes: int8 tcgen05 update kernel (CTA pairs; odd numbers of tile rows), digit
    # planes, row scales, the FP64 segments; then the gradient calls on it (appended rows, lauum pass on the int8 path)
    eng.set_hybrid(1, 2, 2)
    ts5, xs5 = synthetic_series(600)
    lml5, info5 = eng.lml_batch(nodes[:3], noises[:3], ts5, xs5)
    assert eng.hybrid_info()[0] and np.all(info5 == 0) and np.all(np.isfinite(lml5))
    _, hg, hgn, hinfo = eng.lml_grad_batch(nodes[:2], noises[:2], ts5[:400], xs5[:400])
    _, hgn2, hinfo2 = eng.lml_grad_noise_batch(nodes[:2], noises[:2], ts5[:400], xs5[:400])
    assert np.all(hinfo == 0) and np.all(hinfo2 == 0) and np.all(np.isfinite(hgn)) and np.all(np.isfinite(hgn2))
    eng.set_hybrid(-1)
    # a kernel beyond 64 nodes / 64 parameters: the general variant of agp_grad_kernel (parameter windows, big tape)
    big = agp.Periodic(0.5, 0.3, 0.2)
    for j in range(36):
        leaf = agp.Periodic(0.5 + 0.01 * j, 0.3, 0.05) if j % 2 else agp.Linear(0.3, 0.1, 0.05)
        big = agp.Plus(big, leaf) if j % 3 else agp.Times(big, leaf)
    _, bg, bgn, binfo = eng.lml_grad_batch([big, nodes[0]], [0.1, noises[0]], ts[:150], xs[:150])
    assert binfo[0] == 0 and len(bg[0]) > 64 and np.all(np.isfinite(bg[0])) and np.all(np.isfinite(bg[1]))
    print(f"sanitize workload ok: {eng.launch_count} launches")
    eng.close()


if __name__ == "__main__":
    main()
