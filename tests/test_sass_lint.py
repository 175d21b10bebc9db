"""The built persistent kernel keeps the cross-proxy fence that round 1 lacked (profiles/r02_race_experiments.txt):
every release of a TMA-fed operand stage is preceded by FENCE.VIEW.ASYNC before any shared-memory load.  CPU test:
needs only cuobjdump and the object file `make -C autogp.jl_b200/csrc` leaves behind."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
OBJ = os.path.join(ROOT, "autogp.jl_b200", "csrc", "agp_fused.o")


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_every_stage_release_is_fenced_from_the_loads_before_it():
    if not os.path.exists(OBJ):
        pytest.skip("agp_fused.o not built (python -c 'import __graft_entry__ as g; g.build()')")
    import sass_lint

    releases, tma_loads, problems = sass_lint.lint(OBJ)
    assert releases >= 1 and tma_loads >= 2
    assert problems == []
