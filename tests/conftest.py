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


# The analytic-scene suites run twice: through the interval renderer (render_span.cu, the default for scenes of convex
# primitives) and with it switched off, so that the marching kernels it falls back on keep their full coverage.
SPAN_MODULES = {"test_gpu_parity", "test_gpu_fuzz", "test_gpu_properties"}


@pytest.fixture(autouse=True)
def kernel_path(request, monkeypatch):
    mode = getattr(request, "param", None)
    if mode == "march":
        monkeypatch.setenv("XRAY_NO_SPAN", "1")
    elif mode == "span":
        monkeypatch.delenv("XRAY_NO_SPAN", raising=False)
    return mode


def pytest_generate_tests(metafunc):
    if metafunc.module.__name__.split(".")[-1] in SPAN_MODULES and "kernel_path" in metafunc.fixturenames:
        metafunc.parametrize("kernel_path", ["span", "march"], indirect=True)


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


@pytest.fixture(scope="session")
def host_replay(X, tmp_path_factory):
    """tests/c/host_loader_replay.c built with gcc: the Go host's dlopen/dlsym/call sequence from plain C."""
    import shutil
    import subprocess

    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = tmp_path_factory.mktemp("replay") / "host_loader_replay"
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c" / "host_loader_replay.c"),
                           "-o", str(exe), "-ldl", "-lm"])
    return str(exe)
