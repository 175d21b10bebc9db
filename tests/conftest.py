import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_engine_available() -> str:
    """'' when an engine can be created on cuda:0, else the reason (library not built / no device)."""
    try:
        import autogp.jl_b200 as agp

        agp.Engine(0).close()
        return ""
    except Exception as e:  # ImportError (library missing) or AgpError (no CUDA device)
        return f"{type(e).__name__}: {e}"


def pytest_collection_modifyitems(config, items):
    # GPU-marked tests are skipped (not failed) on a machine without a CUDA device, so that a plain `pytest tests`
    # shows CPU-suite regressions; on the GPU box nothing is skipped and a missing library is a hard error there.
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    if os.path.exists("/dev/nvidiactl"):  # a GPU box: never skip, a library that does not load must fail loudly
        return
    why = _cuda_engine_available()
    if why:
        skip = pytest.mark.skip(reason=f"needs a CUDA device and the built library ({why[:120]})")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def engine():
    import autogp.jl_b200 as agp

    eng = agp.Engine(0)
    yield eng
    eng.close()
