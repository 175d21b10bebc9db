"""md5 over the SASS instruction text of agp_chol_kernel (and the functions ptxas kept out of line inside it).

The persistent kernel's main loop is sensitive to how ptxas schedules it (profiles/r01_diag_item_bisect.txt): a build
whose fingerprint differs from the recorded one has to pass the repeated-run soak (tools/stress_check.py) again.

    python tools/sass_fingerprint.py [autogp.jl_b200/csrc/agp_fused.o]
"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECORDED = "ebe3f50d55fa383bae221107a61be640"   # nvcc 12.9.86, flags of csrc/Makefile; soak-tested (8000+ runs)


def fingerprint(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    f = txt[txt.index("Function : _ZN3agp15agp_chol_kernel"):]
    nxt = f.find("Function :", 10)
    f = f[:nxt] if nxt > 0 else f
    ins = [m.group(1) for m in (re.search(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", l) for l in f.split("\n")) if m]
    return len(ins), hashlib.md5("\n".join(ins).encode()).hexdigest()


if __name__ == "__main__":
    obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "autogp.jl_b200", "csrc", "agp_fused.o")
    n, h = fingerprint(obj)
    print(f"{obj}: {n} instructions, md5 {h} ({'matches the soak-tested build' if h == RECORDED else 'DIFFERS from the soak-tested build ' + RECORDED})")
