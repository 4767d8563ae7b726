import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

SCENES = ROOT / "tests" / "scenes"
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _have_gpu() -> bool:
    try:
        import xray_projection_render_b200 as X

        return X._lib.load().XRayDeviceCount() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the product path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def X():
    """The product package; building the in-tree library first if it is missing."""
    lib = ROOT / "xray_projection_render_b200" / "lib" / "libcuda_render.so"
    if not lib.exists():
        import __graft_entry__ as g

        g.build()
    import xray_projection_render_b200 as pkg

    pkg._lib.load()
    return pkg


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle

    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def scenes():
    return SCENES
