"""Structural check of the built persistent kernel (runs on the CPU: cuobjdump only).

Round 1 shipped agp_chol_kernel behind an md5 of its SASS because "semantically equal" main loops produced a wrong tile
in up to one run of four.  The cause (profiles/r02_race_experiments.txt) was a missing cross-proxy fence: the MMA warps
read an operand stage with LDS (generic proxy) and released it to the TMA producer (async proxy) without
fence.proxy.async in between, so a late LDS could see the next box.  What has to hold is therefore a property of the
code, not one particular ptxas schedule:

    in agp_chol_kernel, walking back from every stage release (SYNCS.ARRIVE ... .A1T0 = mbarrier.arrive without
    expect_tx) a FENCE.VIEW.ASYNC must come before the first LDS, and every UTMALDG (TMA tensor read of L) must be
    preceded by a FENCE.VIEW.ASYNC within its basic block chain since the last global acquire.

    python tools/sass_lint.py [autogp.jl_b200/csrc/agp_fused.o]      exit code 1 on violation
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_OBJ = os.path.join(ROOT, "autogp.jl_b200", "csrc", "agp_fused.o")


def chol_sass(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    f = txt[txt.index("Function : _ZN3agp15agp_chol_kernel"):]
    nxt = f.find("Function :", 10)
    f = f[:nxt] if nxt > 0 else f
    return [m.group(1).strip() for m in (re.search(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", l) for l in f.split("\n")) if m]


def lint(obj=DEFAULT_OBJ):
    """Returns (n_releases, n_tma_loads, problems)."""
    ins = chol_sass(obj)
    problems = []
    releases = [i for i, t in enumerate(ins) if "SYNCS.ARRIVE" in t and "A1T0" in t]
    for i in releases:
        j = i - 1
        while j >= 0 and "FENCE.VIEW.ASYNC" not in ins[j]:
            if re.search(r"\bLDS(\.|\b)", ins[j]):
                problems.append(f"stage release at instruction {i} ({ins[i]}): LDS at {j} ({ins[j]}) is not fenced from it")
                break
            if "SYNCS.PHASECHK" in ins[j]:  # reached the wait that opened this stage use without meeting a shared load
                break
            j -= 1
    tma = [i for i, t in enumerate(ins) if "UTMALDG" in t]
    if not releases:
        problems.append("no stage release (SYNCS.ARRIVE ... A1T0) found: the kernel's structure changed, update this check")
    if not tma:
        problems.append("no UTMALDG found: the operand pipeline no longer uses TMA tensor copies, update this check")
    n_fence = sum("FENCE.VIEW.ASYNC" in t for t in ins)
    if n_fence < len(releases) + 1:
        problems.append(f"only {n_fence} FENCE.VIEW.ASYNC for {len(releases)} releases + the item prologue")
    return len(releases), len(tma), problems


if __name__ == "__main__":
    obj = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_OBJ
    r, t, problems = lint(obj)
    print(f"{obj}: {r} stage release(s), {t} UTMALDG, {len(problems)} problem(s)")
    for p in problems:
        print("  " + p)
    sys.exit(1 if problems else 0)
