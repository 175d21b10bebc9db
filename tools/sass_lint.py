"""Structural check of the built persistent kernel (runs on the CPU: cuobjdump only).

Round 1 shipped agp_chol_kernel behind an md5 of its SASS because "semantically equal" main loops produced a wrong tile
in up to one run of four.  The cause (profiles/r02_race_experiments.txt) was a missing cross-proxy fence: the MMA warps
read an operand stage with LDS (generic proxy) and released it to the TMA producer (async proxy) without
fence.proxy.async in between, so a late LDS could see the next box.  What has to hold is therefore a property of the
code, not one particular ptxas schedule:

    in the functions that read TMA-fed stages (agp_chol_kernel in agp_chol_kernel.o: the panel item; do_diag in
    agp_chol_diag.o: the diagonal-tile item), walking back from every
    stage release (SYNCS.ARRIVE ... .A1T0 = mbarrier.arrive without expect_tx) a FENCE.VIEW.ASYNC must come before
    the first LDS; the function must still use TMA tensor copies (UTMALDG); and its main loop must be free of
    local-memory traffic (spills there cost the round-2 builds up to 2.4x, agp_chol_common.cuh).

    and, for the hybrid schedule's int8 kernel (agp_ozaki_update2_kernel in agp_ozaki.o): the 5th-generation tensor path must
    really be there — UTCIMMA (tcgen05.mma kind::i8; .2CTA for the pair variant), UTCBAR (tcgen05.commit), LDTM (tcgen05.ld),
    UTMALDG (TMA tensor loads) — and the MMA issue sequence must be free of per-instruction election loops (no BRA between two
    UTCIMMA of a product's four k-steps: with the issue loop inside `if (lane == 0)` ptxas emitted one, 79 instead of 64 clocks).

    python tools/sass_lint.py [object files]      exit code 1 on violation
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "autogp.jl_b200", "csrc")
DEFAULT_OBJS = [os.path.join(CSRC, f) for f in ("agp_chol_kernel.o", "agp_chol_diag.o", "agp_chol_kernel_solo.o", "agp_chol_diag_solo.o")]
FUNCS = ("agp_chol_kernel", "do_diag")  # the functions that read TMA-fed operand stages (panel item inlined in the kernel; diagonal-tile item)


def function_sass(obj):
    """{function name: [instruction text]} for the phase functions found in `obj`."""
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out = {}
    for part in txt.split("Function : ")[1:]:
        name = part.split("\n", 1)[0].strip()
        hit = [f for f in FUNCS if f in name]
        if not hit:
            continue
        out[hit[0]] = [(int(m.group(1), 16), m.group(2).strip())
                       for m in (re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l) for l in part.split("\n")) if m]
    return out


def lint_function(name, code):
    problems = []
    addr = [a for a, _ in code]
    ins = [t for _, t in code]
    releases = [i for i, t in enumerate(ins) if "SYNCS.ARRIVE" in t and "A1T0" in t]
    for i in releases:
        j = i - 1
        while j >= 0 and "FENCE.VIEW.ASYNC" not in ins[j]:
            if re.search(r"\bLDS(\.|\b)", ins[j]):
                problems.append(f"{name}: stage release at instruction {i} ({ins[i]}): LDS at {j} ({ins[j]}) is not fenced from it")
                break
            if "SYNCS.PHASECHK" in ins[j]:  # reached the wait that opened this stage use without meeting a shared load
                break
            j -= 1
        # the main loop = the innermost backward branch around the release: no local-memory instruction inside
        loop = None
        for kk in range(i, len(ins)):
            m = re.search(r"\bBRA(?:\.U)?(?:\s+\S+,)?\s+(0x[0-9a-f]+)", ins[kk])
            if m and int(m.group(1), 16) <= addr[i]:
                loop = (int(m.group(1), 16), kk)
                break
        if loop is None:
            problems.append(f"{name}: no loop found around the release at {i}, update this check")
            continue
        body = [k for k in range(len(ins)) if loop[0] <= addr[k] and k <= loop[1]]
        spills = [k for k in body if re.match(r"(@!?U?P\d+ )?(STL|LDL)", ins[k])]
        if spills:
            problems.append(f"{name}: {len(spills)} local-memory instruction(s) among the {len(body)} of the main loop around the release at {i} (first: {ins[spills[0]]})")
    tma = [i for i, t in enumerate(ins) if "UTMALDG" in t]
    if not releases:
        problems.append(f"{name}: no stage release (SYNCS.ARRIVE ... A1T0) found: the function's structure changed, update this check")
    if not tma:
        problems.append(f"{name}: no UTMALDG found: the operand pipeline no longer uses TMA tensor copies, update this check")
    return len(releases), len(tma), problems


def lint_int8(obj=None):
    """Problems of the int8 update kernels in agp_ozaki.o (see the module docstring)."""
    obj = obj or os.path.join(CSRC, "agp_ozaki.o")
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    problems, seen = [], 0
    for part in txt.split("Function : ")[1:]:
        name = part.split("\n", 1)[0].strip()
        if "agp_ozaki_update2_kernel" not in name:
            continue
        seen += 1
        pair = "ILi2E" in name
        ins = [m.group(1).strip() for m in (re.search(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", l) for l in part.split("\n")) if m]
        for op in (("UTCIMMA.2CTA" if pair else "UTCIMMA"), "UTCBAR", "LDTM", "UTMALDG"):
            if not any(op in t for t in ins):
                problems.append(f"{name}: no {op} in the SASS")
        mma = [i for i, t in enumerate(ins) if "UTCIMMA" in t]
        # the four k-steps of a digit-plane product are issued back to back: at most a handful of uniform-register instructions between them
        runs = sum(1 for a, b in zip(mma, mma[1:]) if b - a <= 6)
        if mma and runs < len(mma) // 2:
            problems.append(f"{name}: only {runs} of {len(mma)} UTCIMMA are issued back to back (per-instruction election loop?)")
    if seen < 2:
        problems.append(f"agp_ozaki_update2_kernel<1> / <2>: {seen} of 2 found in {obj}")
    return problems


def lint(objs=None):
    """Returns (n_releases, n_tma_loads, problems) over the phase functions."""
    found, rel, tma, problems = set(), 0, 0, []
    for obj in (objs or DEFAULT_OBJS):
        for name, ins in function_sass(obj).items():
            found.add(name)
            r, t, p = lint_function(name, ins)
            rel, tma, problems = rel + r, tma + t, problems + p
    for f in FUNCS:
        if f not in found:
            problems.append(f"{f}: not found in {objs or DEFAULT_OBJS}")
    return rel, tma, problems


if __name__ == "__main__":
    objs = sys.argv[1:] or DEFAULT_OBJS
    r, t, problems = lint(objs)
    if not sys.argv[1:]:
        p8 = lint_int8()
        print(f"agp_ozaki.o: int8 update kernels, {len(p8)} problem(s)")
        problems += p8
    print(f"{', '.join(os.path.basename(o) for o in objs)}: {r} stage release(s), {t} UTMALDG, {len(problems)} problem(s)")
    for p in problems:
        print("  " + p)
    sys.exit(1 if problems else 0)
