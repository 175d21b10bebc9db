"""The built persistent kernel keeps the cross-proxy fence that round 1 lacked (profiles/r02_race_experiments.txt):
every release of a TMA-fed operand stage is preceded by FENCE.VIEW.ASYNC before any shared-memory load.  CPU test:
needs only cuobjdump and the object files `make -C autogp.jl_b200/csrc` leaves behind.  The same check keeps the
register-critical main loops free of local-memory spills."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
OBJS = [os.path.join(ROOT, "autogp.jl_b200", "csrc", f) for f in ("agp_chol_kernel.o", "agp_chol_diag.o")]


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_every_stage_release_is_fenced_from_the_loads_before_it():
    if not all(os.path.exists(o) for o in OBJS):
        pytest.skip("phase-function objects not built (python -c 'import __graft_entry__ as g; g.build()')")
    import sass_lint

    releases, tma_loads, problems = sass_lint.lint(OBJS)
    assert releases >= 2 and tma_loads >= 3
    assert problems == []


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_int8_update_kernels_are_on_the_tcgen05_path():
    """The hybrid schedule's int8 kernels really use the 5th-generation tensor path (UTCIMMA with TMEM accumulators, UTCBAR,
    LDTM, UTMALDG; .2CTA in the pair variant) and issue a product's k-steps back to back."""
    obj = os.path.join(ROOT, "autogp.jl_b200", "csrc", "agp_ozaki.o")
    if not os.path.exists(obj):
        pytest.skip("agp_ozaki.o not built (python -c 'import __graft_entry__ as g; g.build()')")
    import sass_lint

    assert sass_lint.lint_int8(obj) == []
