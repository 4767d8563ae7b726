"""integration/go_host.patch must apply cleanly to the reference checkout (git apply --check needs no Go toolchain)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
PATCH = ROOT / "integration" / "go_host.patch"


def test_patch_touches_the_files_the_boundary_names():
    text = PATCH.read_text()
    for f in ("cuda_backend.go", "cuda_path.go", "api.go", "main.go", "xray_projection_render/xray_renderer.py"):
        assert f"+++ b/{f}" in text
    for sym in ("XRaySceneCompileJSON", "XRaySceneFree", "XRaySceneNumVoxelSlots", "XRayRenderOptsInit", "XRayRenderSceneCUDA",
                "XRayLastError"):
        assert f'"{sym}"' in text, sym
    assert "params.UseCuda" in text and "use_cuda: not supported via API path" in text   # api.go:159 replaced


def test_struct_layout_in_patch_matches_the_header():
    """The XRayRenderOpts / XRayCameraParams64 typedefs the patch adds to the cgo preamble must have the library's layout."""
    src = r'''
#include <stddef.h>
#include <stdio.h>
#include "xray_cuda_render.h"
typedef struct { double eye[3]; double view[16]; double fov_y; double R; } Cam64;
typedef struct {
    unsigned int struct_size; int integration; int precision; int out_dtype;
    double ds; double flat_field; double density_multiplier;
    int num_devices; int devices[16];
    unsigned long long stream; unsigned long long* stats;
    int view_begin; int reserved[7];
} Opts;
int main(void) {
    int ok = sizeof(Cam64) == sizeof(XRayCameraParams64) && sizeof(Opts) == sizeof(XRayRenderOpts)
        && offsetof(Opts, ds) == offsetof(XRayRenderOpts, ds) && offsetof(Opts, num_devices) == offsetof(XRayRenderOpts, num_devices)
        && offsetof(Opts, stream) == offsetof(XRayRenderOpts, stream) && offsetof(Opts, stats) == offsetof(XRayRenderOpts, stats)
        && offsetof(Opts, view_begin) == offsetof(XRayRenderOpts, view_begin);
    printf("%d %zu\n", ok, sizeof(Opts));
    return ok ? 0 : 1;
}
'''
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        c = Path(td) / "t.c"
        c.write_text(src)
        subprocess.check_call(["gcc", "-I", str(ROOT / "include"), str(c), "-o", str(Path(td) / "t")])
        assert subprocess.run([str(Path(td) / "t")]).returncode == 0
    # and the typedef text in the patch is the one compiled above
    text = PATCH.read_text()
    assert "unsigned int struct_size; int integration; int precision; int out_dtype;" in text
    assert "int num_devices; int devices[16];" in text and "unsigned long long stream; unsigned long long* stats;" in text


def test_patch_applies_to_the_reference_checkout(tmp_path):
    if not REF.exists() or not shutil.which("git"):
        pytest.skip("reference checkout or git not available")
    work = tmp_path / "ref"
    shutil.copytree(REF, work, ignore=shutil.ignore_patterns(".git"))
    subprocess.check_call(["git", "init", "-q", "."], cwd=work)
    r = subprocess.run(["git", "apply", "--check", str(PATCH)], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    subprocess.check_call(["git", "apply", str(PATCH)], cwd=work)
    assert "cuda_dl_RenderSceneCUDA" in (work / "cuda_backend.go").read_text()
    assert "renderSceneCUDA(objJSON, dfJSON, cams64" in (work / "cuda_path.go").read_text()
