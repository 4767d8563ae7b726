"""Host-side logic of the product, runnable without a GPU: scene / deformation file front end
(main.go:50-120 + FromMap type rules), native scene compiler vs the oracle, camera helpers,
output formats, and the C ABI surface (symbols, struct layouts, argument validation)."""
import ctypes
import json
import math
import re
import struct
import zlib
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
ALL = ["cube_w_hole", "balls", "box_w_pped", "pillar_array", "lattice", "gyroid_example"]


# ---- C ABI surface ------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(X):
    L = X._lib.load()
    hdr = (ROOT / "include" / "xray_cuda_render.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b([A-Z][A-Za-z0-9]+)\s*\(", hdr)) - {"XRAY"}
    names = {n for n in names if n.startswith(("XRay", "Assemble", "Render"))}
    assert len(names) >= 22
    for n in sorted(names):
        assert hasattr(L, n), f"libcuda_render.so does not export {n}"
    assert set(X._lib.LEGACY_SYMBOLS) <= names and set(X._lib.EXTENDED_SYMBOLS) <= names


def test_go_host_load_sequence_from_c(X, host_replay, monkeypatch):
    """cuda_backend.go:27-43,103-114 replayed by a C program that includes only include/xray_cuda_render.h:
    dlopen(RTLD_LAZY|RTLD_LOCAL) of the path in XRAY_CUDA_LIB, the three legacy symbols (all required), struct sizes,
    and non-zero returns (never a crash or exit) for bad arguments.  No compute call: runs without a GPU."""
    import subprocess

    lib = str(ROOT / "xray_projection_render_b200" / "lib" / "libcuda_render.so")
    r = subprocess.run([host_replay, lib, "symbols"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    monkeypatch.setenv("XRAY_CUDA_LIB", lib)
    r = subprocess.run([host_replay, "-", "symbols"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok symbols", r.stderr
    r = subprocess.run([host_replay, "/nonexistent/libcuda_render.so", "symbols"], capture_output=True, text=True)
    assert r.returncode == 1  # the loader's -1: dlopen failed
    r = subprocess.run([host_replay, str(ROOT / "oracle" / "libxray_oracle.so"), "symbols"], capture_output=True, text=True)
    assert r.returncode == 2  # a library without the three symbols is refused (-2), as the Go loader does


def test_legacy_struct_layouts(X):
    # SURVEY.md 8b [probe]: CylinderParams 32 B, XRayCameraParams 84 B (view@12, fov_y@76, R@80), align 4
    C, P = X._lib.CylinderParams, X._lib.XRayCameraParams
    assert ctypes.sizeof(C) == 32 and ctypes.alignment(C) == 4
    assert ctypes.sizeof(P) == 84 and ctypes.alignment(P) == 4
    assert (P.view.offset, P.fov_y.offset, P.R.offset) == (12, 76, 80)
    assert (C.p1.offset, C.radius.offset, C.rho.offset) == (12, 24, 28)
    assert ctypes.sizeof(X._lib.XRayCameraParams64) == 21 * 8  # eye 3 + view 16 + fov + R


def test_opts_struct_matches_library(X):
    o = X._lib.make_opts()
    assert o.struct_size == ctypes.sizeof(X._lib.XRayRenderOpts)
    assert (o.integration, o.precision, o.out_dtype) == (1, 0, 0)  # hierarchical is the reference default (main.go:39)
    assert o.ds == -1.0 and o.density_multiplier == 1.0 and o.flat_field == 0.0


def test_legacy_argument_validation_without_gpu(X):
    """Error contract of cuda_backend.cu:95-101: non-zero on null pointers / bad dims, never a crash."""
    L = X._lib.load()
    fp = ctypes.POINTER(ctypes.c_float)
    vol = (ctypes.c_float * 8)()
    out = (ctypes.c_float * 4)()
    cam = (X._lib.XRayCameraParams * 1)()
    assert L.RenderVolumeProjectionsCUDA(None, 2, 2, 2, cam, 1, 2, 0.1, 0.0, out) == 1
    assert L.RenderVolumeProjectionsCUDA(vol, 2, 2, 2, None, 1, 2, 0.1, 0.0, out) == 1
    assert L.RenderVolumeProjectionsCUDA(vol, 2, 2, 2, cam, 1, 2, 0.1, 0.0, None) == 1
    assert L.RenderVolumeProjectionsCUDA(vol, 0, 2, 2, cam, 1, 2, 0.1, 0.0, out) == 2
    assert L.RenderVolumeProjectionsCUDA(vol, 2, 2, 2, cam, 0, 2, 0.1, 0.0, out) == 2
    assert L.RenderVolumeProjectionsCUDA(vol, 2, 2, 2, cam, 1, 0, 0.1, 0.0, out) == 2
    assert L.RenderVolumeProjectionsCUDA(vol, 2, 2, 2, cam, 1, 2, 0.0, 0.0, out) == 2
    assert L.XRayLastError() != b""
    cyl = (X._lib.CylinderParams * 1)()
    assert L.AssembleVoxelGridCUDA(None, 1, 4, 1.0, out) == 1
    assert L.AssembleVoxelGridCUDA(cyl, 0, 4, 1.0, out) == 2
    assert L.AssembleVoxelGridCUDA(cyl, 1, 0, 1.0, out) == 2
    offs = (ctypes.c_int * 9)(*([0] * 9))
    assert L.AssembleVoxelGridSpatialCUDA(cyl, 1, 4, 1.0, 2, None, None, 0, out) == 1
    assert L.AssembleVoxelGridSpatialCUDA(cyl, 1, 4, 1.0, 0, offs, None, 0, out) == 2
    bad = (ctypes.c_int * 9)(0, 1, 1, 1, 1, 1, 1, 1, 3)  # claims 3 indices, none given
    idx = (ctypes.c_int * 1)(0)
    assert L.AssembleVoxelGridSpatialCUDA(cyl, 1, 4, 1.0, 2, bad, idx, 1, out) == 2


def test_scene_compile_errors(X):
    L = X._lib.load()
    h = ctypes.c_void_p()
    assert L.XRaySceneCompileJSON(b'{"type":"torus"}', None, ctypes.byref(h)) != 0
    assert b"unknown object type `torus`" in L.XRayLastError()  # objects.go:676
    assert L.XRaySceneCompileJSON(b'{"type":"sphere","center":[0,0,0],"rho":1.0}', None, ctypes.byref(h)) != 0
    assert b"radius is not a float64" in L.XRayLastError()
    assert L.XRaySceneCompileJSON(b'{not json', None, ctypes.byref(h)) != 0
    nested = {"type": "object_collection", "objects": [{"type": "object_collection", "objects": []}]}
    assert L.XRaySceneCompileJSON(json.dumps(nested).encode(), None, ctypes.byref(h)) != 0
    assert b"unknown object type" in L.XRayLastError()  # objects.go:407: collections cannot nest
    ok = b'{"type":"sphere","center":[0,0,0],"radius":0.5,"rho":1.0}'
    assert L.XRaySceneCompileJSON(ok, b'{"type":"twist"}', ctypes.byref(h)) != 0
    assert b"unknown deformation type twist" in L.XRayLastError()
    assert L.XRaySceneCompileJSON(ok, b'{"type":"sigmoid","amplitude":1,"center":0,"lengthscale":1,"direction":"w"}',
                                  ctypes.byref(h)) != 0


# ---- scene front end ------------------------------------------------------------------------
@pytest.mark.parametrize("name", ALL)
def test_compiled_scene_matches_oracle_on_host(X, O, scenes, name):
    sc = X.Scene(str(scenes / f"{name}.json"))
    osc = O.OracleScene(str(scenes / f"{name}.json"))
    assert sc.auto_ds() == osc.auto_ds()
    rng = np.random.default_rng(0)
    for p in rng.uniform(-1.1, 1.1, size=(3000, 3)):
        assert sc.density_host(*p) == osc.density(*p)


@pytest.mark.parametrize("deform", ["deformation_sigmoid", "deformation_linear"])
def test_compiled_deformation_matches_oracle_on_host(X, O, scenes, deform):
    sc = X.Scene(str(scenes / "gyroid_example.json"), str(scenes / f"{deform}.json"))
    osc = O.OracleScene(str(scenes / "gyroid_example.json"), str(scenes / f"{deform}.json"), density_multiplier=1.5)
    rng = np.random.default_rng(1)
    for p in rng.uniform(-1.1, 1.1, size=(3000, 3)):
        assert sc.density_host(*p, 1.5) == osc.density(*p)


def test_scene_bounds_are_conservative(X, O, scenes):
    """Outside XRaySceneBounds density() must be exactly 0 (ray clipping relies on it), warps included."""
    cases = [("cube_w_hole", None), ("lattice", None), ("gyroid_example", "deformation_sigmoid"),
             ("cube_w_hole", "deformation_linear"), ("balls", None)]
    rng = np.random.default_rng(2)
    for name, d in cases:
        dp = str(scenes / f"{d}.json") if d else None
        sc = X.Scene(str(scenes / f"{name}.json"), dp)
        osc = O.OracleScene(str(scenes / f"{name}.json"), dp)
        lo, hi = (np.array(v) for v in sc.bounds())
        pts = rng.uniform(-1.8, 1.8, size=(20000, 3))
        outside = np.any((pts < lo) | (pts > hi), axis=1)
        assert outside.sum() > 100
        for p in pts[outside][:4000]:
            assert osc.density(*p) == 0.0
        inside_nonzero = sum(osc.density(*p) != 0.0 for p in pts[~outside][:2000])
        assert inside_nonzero > 0


def test_yaml_type_strictness(X, tmp_path):
    """objects.go `.(float64)` assertions: YAML ints are rejected where the Go code asserts float64."""
    good = tmp_path / "a.yaml"
    good.write_text("type: sphere\ncenter: [0, 0, 0]\nradius: 0.5\nrho: 1.0\n")
    X.Scene(str(good))
    bad = tmp_path / "b.yaml"
    bad.write_text("type: sphere\ncenter: [0, 0, 0]\nradius: 1\nrho: 1.0\n")
    with pytest.raises(X.SceneError, match="radius is not a float64"):
        X.Scene(str(bad))
    cube_bad = tmp_path / "c.yaml"
    cube_bad.write_text("type: cube\ncenter: [0, 0, 0]\nside: 1.5\nrho: 0.7\n")  # cube center elements must be floats
    with pytest.raises(X.SceneError, match="center"):
        X.Scene(str(cube_bad))
    box_ok = tmp_path / "d.yaml"
    box_ok.write_text("type: box\ncenter: [0, 0, 0]\nsides: [1, 2, 3]\nrho: 1\n")  # ToVec / ToFloat64 accept ints
    assert X.Scene(str(box_ok)).min_feature_size() == pytest.approx(0.1)
    cyl = tmp_path / "e.yaml"
    cyl.write_text("type: cylinder\np0: [0, 0, -1]\np1: [0, 0, 1]\nradius: 0.5\n")  # rho defaults to 1.0
    assert X.Scene(str(cyl)).density_host(0, 0, 0) == 1.0
    sci = tmp_path / "f.yaml"
    sci.write_text("type: sphere\ncenter: [0.0, 0.0, 0.0]\nradius: 5e-1\nrho: 1.0\n")  # yaml.v3 reads 5e-1 as a float
    assert X.Scene(str(sci)).min_feature_size() == 0.5


def test_extension_sniffing(X, tmp_path):  # main.go:63,97: last four characters decide
    p = tmp_path / "scene.yml"
    p.write_text("type: sphere\ncenter: [0.0, 0.0, 0.0]\nradius: 0.5\nrho: 1.0\n")
    with pytest.raises(ValueError, match="Unknown file extension"):
        X.Scene(str(p))
    j = tmp_path / "scene.json"
    j.write_text('{"type":"sphere","center":[0,0,0],"radius":1,"rho":1}')  # JSON numbers are always float64
    assert X.Scene(str(j)).min_feature_size() == 1.0


def test_unit_cell_forces_greedy_and_ignores_types(X, O):
    # objects.go:481-487: uc.objects is parsed as a collection whatever its type says, greedy forced on
    uc = {"objects": {"objects": [{"type": "sphere", "center": [0.5, 0.5, 0.5], "radius": 0.3, "rho": 0.8},
                                  {"type": "sphere", "center": [0.5, 0.5, 0.5], "radius": 0.3, "rho": 0.8}]},
          "xmin": 0.0, "xmax": 1.0, "ymin": 0.0, "ymax": 1.0, "zmin": 0.0, "zmax": 1.0}
    t = {"type": "tessellated_obj_coll", "uc": uc, "xmin": -2.0, "xmax": 2.0, "ymin": -2.0, "ymax": 2.0, "zmin": -2.0, "zmax": 2.0}
    assert X.Scene(t).density_host(0.5, 0.5, 0.5) == 0.8 == O.OracleScene(t).density(0.5, 0.5, 0.5)
    assert X.Scene(t).density_host(-0.5, 1.5, 0.5) == 0.8


def test_voxel_raw_loader(X, O, tmp_path):
    rng = np.random.default_rng(3)
    for dtype, npdt in [("uint8", "<u1"), ("uint16", "<u2"), ("uint32", "<u4"), ("float32", "<f4"), ("float64", "<f8")]:
        nx, ny, nz = 3, 4, 2
        raw = (rng.random(nx * ny * nz) * (1.0 if "float" in dtype else 200.0)).astype(npdt)
        f = tmp_path / f"v_{dtype}.raw"
        raw.tofile(f)
        scene = {"type": "voxel_grid", "path": f.name, "resolution": [nx, ny, nz], "dtype": dtype}
        sf = tmp_path / f"v_{dtype}.json"
        sf.write_text(json.dumps(scene))
        sc, osc = X.Scene(str(sf)), O.OracleScene(str(sf))
        assert sc.min_feature_size() == osc.min_feature_size() == 2.0 / 4
        for p in rng.uniform(-1, 1, size=(200, 3)):
            a, b = sc.density_host(*p), osc.density(*p)
            assert a == b
    with pytest.raises(X.SceneError, match="file size"):
        X.voxel_grid_from_raw(str(tmp_path / "v_uint8.raw"), [5, 5, 5], "uint8")


# ---- camera ---------------------------------------------------------------------------------
def test_camera_matches_oracle_bitwise(X, O):
    for az, pol, R in [(90, 90, 4), (0, 90, 4), (123.4, 70.0, 4.0), (359.0, 45.0, 3.0), (200.0, 135.0, 6.0)]:
        cam = X.camera_from_angles(az, pol, R, 40.0)
        eye, m = O.camera_from_angles(az, pol, R)
        assert list(cam.eye) == list(eye)
        assert np.array_equal(X.camera_matrix(cam), m)
        assert cam.fov_y == 40.0 and cam.R == R


def test_generate_camera_angles_and_sharding(X):
    a = X.generate_camera_angles(360)
    assert len(a) == 360 and a[0] == {"azimuthal": 90.0, "polar": 90.0} and a[1]["azimuthal"] == 91.0
    shards = [X.generate_camera_angles(10, r, 4) for r in range(4)]  # --jobs_modulo 4 --job r
    seen = sorted(x["azimuthal"] for s in shards for x in s)
    assert seen == sorted(x["azimuthal"] for x in X.generate_camera_angles(10))
    assert [len(s) for s in shards] == [3, 3, 2, 2]
    assert X.parse_float_list("1, 2.5 ,3") == [1.0, 2.5, 3.0] and X.parse_float_list("") == []
    assert X.parse_float_list("90,,180, ") == [90.0, 180.0]  # main.go:268-270 skips empty fields
    with pytest.raises(ValueError, match="invalid float value 'x'"):
        X.parse_float_list("1,x")


def test_legacy_camera_narrowing_roundtrip(X):
    cams = X.cameras_from_angles([(37.0, 60.0)], 4.0, 40.0)
    c32 = X.to_legacy(cams)
    back = X.from_legacy(c32)
    assert back[0].eye[0] == float(np.float32(cams[0].eye[0]))
    assert back[0].view[5] == float(np.float32(cams[0].view[5]))
    assert back[0].R == 4.0 and back[0].fov_y == 40.0


# ---- output formats -------------------------------------------------------------------------
def test_png_quantisation_and_orientation(X, tmp_path):
    img = np.array([[0.0, 0.25], [0.5, 1.0]])  # img[i][j]
    rgba = X.image_to_rgba8(img)
    # main.go:495-498: uint16(val*0xffff) >> 8, pixel (i,j) at x=i, y=res-1-j
    q = lambda v: int(v * 0xFFFF) >> 8
    assert rgba[1, 0, 0] == q(0.0) and rgba[0, 0, 0] == q(0.25) and rgba[1, 1, 0] == q(0.5) and rgba[0, 1, 0] == q(1.0)
    assert (rgba[..., 3] == 255).all()
    t = X.image_to_rgba8(img, transparency=True)
    assert t[0, 1, 3] == 0 and t[1, 0, 3] == 255  # alpha 0 only where val == 1 (main.go:486-492)
    assert tuple(t[0, 1]) == (0, 0, 0, 0)          # png.Encode un-premultiplies: alpha 0 -> colour 0
    # Go's png.Encode of an opaque *image.RGBA writes 8-bit RGB (colour type 2) ...
    p = tmp_path / "x.png"
    X.write_png(str(p), rgba)
    data = p.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    w, h, depth, ctype = struct.unpack(">IIBB", data[16:26])
    assert (w, h, depth, ctype) == (2, 2, 8, 2)
    idat = data[data.index(b"IDAT") + 4:data.index(b"IEND") - 8]
    raw = zlib.decompress(idat)
    assert raw == b"".join(b"\x00" + rgba[y, :, :3].tobytes() for y in range(2))
    # ... and 8-bit RGBA (colour type 6) as soon as one pixel is transparent
    X.write_png(str(p), t)
    data = p.read_bytes()
    assert struct.unpack(">IIBB", data[16:26]) == (2, 2, 8, 6)
    raw = zlib.decompress(data[data.index(b"IDAT") + 4:data.index(b"IEND") - 8])
    assert raw == b"".join(b"\x00" + t[y].tobytes() for y in range(2))


def test_object_json_follows_the_reference_tomap_schema(X, scenes):
    """main.go:538-546 writes lat[0].ToMap(): per type exactly the keys of the ToMap methods (objects.go) -- no greedy flag,
    voxel grids as nx / ny / nz / dtype float64 / path."""
    from xray_projection_render_b200.scene import reference_object_map

    sc = X.Scene(str(scenes / "lattice.json"))
    m = reference_object_map(sc.object_map)
    assert set(m) == {"type", "uc", "xmin", "xmax", "ymin", "ymax", "zmin", "zmax"}
    assert set(m["uc"]) == {"type", "objects", "xmin", "xmax", "ymin", "ymax", "zmin", "zmax"} and m["uc"]["type"] == "unit_cell"
    assert set(m["uc"]["objects"]) == {"type", "objects"}
    assert set(m["uc"]["objects"]["objects"][0]) == {"type", "p0", "p1", "radius", "rho"}
    v = reference_object_map({"type": "voxel_grid", "resolution": [4.0, 5.0, 6.0], "path": "vol.raw"})
    assert v == {"type": "voxel_grid", "nx": 4, "ny": 5, "nz": 6, "dtype": "float64", "path": "vol.raw"}


def test_renderer_parameter_validation(X):
    r = X.XRayRenderer()
    with pytest.raises(ValueError, match="'input' parameter is required"):
        r.render({})
    out = r.render({"input": "x.json", "ds": 0})
    assert out["success"] is False and "ds is 0" in out["error"]  # api.go:96-98
    out = r.render({"input": "x.json", "density_multiplier": 0})
    assert out["success"] is False and "density_multiplier is 0" in out["error"]  # api.go:99-101


def _gyroid_records(X, obj, deform):
    """(second float4, third float4) of the first gyroid's fp32 record in the compiled program (program.h layout)."""
    blob = X.Scene(obj, deform).program_bytes()
    hdr = struct.unpack_from("<16I", blob, 0)
    n_instr, instr_off, f32_off, f32_count = hdr[3], hdr[4], hdr[5], hdr[6]
    instr = np.frombuffer(blob, dtype=np.uint32, count=n_instr * 8, offset=instr_off).reshape(-1, 8)
    f32 = np.frombuffer(blob, dtype=np.float32, count=f32_count * 4, offset=f32_off).reshape(-1, 4)
    run = next(r for r in instr if r[0] == 5)  # OP_GYROID
    return f32[run[4] + 1], f32[run[4] + 2]


def test_gyroid_second_order_bound_in_the_program(X):
    """The third float4 of a gyroid record carries the second-order skip bound M2 = 1.01 (2 (|J|/scale)^2 + 3 curv/scale)
    of render_fast.cu prim_gyroid_so.  It must be EXACTLY zero (rule off, first-order Lipschitz bound used) for warp
    chains without a curvature bound -- a denormal there once switched the rule on with the un-warped ray direction."""
    cell = {"type": "tessellated_obj_coll", "xmin": -0.7, "xmax": 0.7, "ymin": -0.6, "ymax": 0.6, "zmin": -0.65, "zmax": 0.65,
            "uc": {"xmin": -1.0, "xmax": 1.0, "ymin": -1.0, "ymax": 1.0, "zmin": -1.0, "zmax": 1.0,
                   "objects": {"objects": [{"type": "gyroid", "center": [0.0, 0.0, 0.0], "scale": 0.125, "thickness": 0.25, "rho": 0.8}]}}}
    b, c = _gyroid_records(X, cell, None)
    assert b[0] == np.float32(8.0) and c[0] == pytest.approx(1.01 * 2 * 64.0, rel=1e-6) and c[2] == pytest.approx(1.0, rel=1e-6)
    A, L = 0.2, 0.1
    b, c = _gyroid_records(X, cell, {"type": "sigmoid", "amplitude": A, "center": 0.0, "lengthscale": L, "direction": "z"})
    jac, curv = 1 + A / (4 * L), A * 0.0962251 / L**2
    assert c[0] == pytest.approx(1.01 * (2 * (jac / 0.125) ** 2 + 3 * curv / 0.125), rel=1e-6) and c[2] == pytest.approx(jac, rel=1e-6)
    for d in ({"type": "rigid", "displacements": [0.1, 0.0, 0.0]}, {"type": "linear", "strains": [0.01, 0.0, 0.0, 0.02, 0.0, 0.0]},
              {"type": "affine", "matrix": [[1.1, 0.0, 0.0], [0.0, 0.9, 0.1], [0.0, 0.0, 1.0]]}):
        assert _gyroid_records(X, cell, d)[1][0] > 0.0
    for d in ({"type": "gaussian", "amplitudes": [0.1, 0.0, -0.1], "sigmas": [0.3, 0.4, 0.5], "centers": [0.1, 0.0, -0.2]},
              {"type": "composed", "deformations": [{"type": "rigid", "displacements": [0.1, 0.0, 0.0]},
                                                    {"type": "linear", "strains": [0.01, 0.02, 0.03, 0.0, 0.0, 0.05]}]}):
        assert _gyroid_records(X, cell, d)[1][0] == 0.0  # exactly: the kernel tests `> 0`


def test_scene_compiler_survives_malformed_json():
    """Contract of the plugin boundary (api.cu header; cuda_backend.cu never exits either): whatever bytes arrive, the
    compiler returns a code.  2000 malformed variants of the bundled scenes, in a subprocess so that a crash is a failure."""
    import subprocess
    import sys
    from pathlib import Path

    script = Path(__file__).resolve().parent / "dev" / "fuzz_compile.py"
    r = subprocess.run([sys.executable, str(script), "11", "2000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    assert "2000 inputs" in r.stdout and " compiled, " in r.stdout


def test_scene_compiler_survives_edge_values():
    """Well-formed scenes with degenerate numbers (zero-size unit cells, zero / negative radii, 1e308 extents, denormals): the
    compiler answers promptly with 0 or an error code -- no crash, no runaway candidate-grid construction."""
    import subprocess
    import sys
    from pathlib import Path

    script = Path(__file__).resolve().parent / "dev" / "fuzz_compile_values.py"
    r = subprocess.run([sys.executable, str(script), "21", "1200"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    assert "1200 inputs" in r.stdout and ", 0 slow" in r.stdout, r.stdout[-500:]


def test_renderer_batches_views_and_threads_the_png_writes(X, scenes, tmp_path, monkeypatch):
    """XRayRenderer.render pushes the views through the library in bounded batches and encodes frames on a thread pool;
    file names, bytes, frame order and transforms.json must not depend on the batching.  The library call is replaced by
    a stub here (no GPU in the CPU suite); tests/test_gpu_properties.py and test_gpu_reference_go_output.py run the real one."""
    import hashlib
    import json

    import xray_projection_render_b200.renderer as RR

    calls = []

    def fake_render_scene(scene, cams, res, **kw):
        calls.append(len(cams))
        out = np.empty((len(cams), res, res), dtype=np.float32)
        ii, jj = np.meshgrid(np.arange(res), np.arange(res), indexing="ij")
        for k, c in enumerate(cams):
            phase = float(c.view[3]) + 2.0 * float(c.view[7])  # the eye position: differs per view
            out[k] = 0.5 + 0.5 * np.sin(0.37 * ii + 0.11 * jj + phase)
        return out

    monkeypatch.setattr(RR, "render_scene", fake_render_scene)
    digests = {}
    for label, batch_bytes in (("one_batch", 1 << 30), ("two_views_per_batch", 2 * 40 * 40 * 4), ("one_view_per_batch", 1)):
        calls.clear()
        out_dir = tmp_path / label / "images"
        r = X.XRayRenderer()
        monkeypatch.setattr(r, "RENDER_BATCH_BYTES", batch_bytes, raising=False)
        res = r.render({"input": str(scenes / "cube_w_hole.json"), "output_dir": str(out_dir), "resolution": 40, "num_images": 7,
                        "transforms_file": str(tmp_path / label / "transforms.json")})
        assert res["success"] and res["rendered"] == 7
        assert calls == {"one_batch": [7], "two_views_per_batch": [2, 2, 2, 1], "one_view_per_batch": [1] * 7}[label]
        files = sorted(p.name for p in out_dir.iterdir())
        assert files == [f"image_{k:03d}.png" for k in range(7)]
        tf = json.loads((tmp_path / label / "transforms.json").read_text())
        assert [f["file_path"] for f in tf["frames"]] == [f"images/image_{k:03d}.png" for k in range(7)]
        digests[label] = [hashlib.sha256((out_dir / f).read_bytes()).hexdigest() for f in files] + [json.dumps(tf["frames"])]
    assert digests["one_batch"] == digests["two_views_per_batch"] == digests["one_view_per_batch"]
    assert len(set(digests["one_batch"][:7])) == 7  # seven different frames
    # a frame that cannot be written fails the call as a whole (the error of a worker thread is not lost)
    blocked = tmp_path / "blocked" / "images"
    blocked.mkdir(parents=True)
    (blocked / "image_003.png").mkdir()
    bad = X.XRayRenderer().render({"input": str(scenes / "cube_w_hole.json"), "output_dir": str(blocked), "resolution": 40,
                                   "num_images": 7, "transforms_file": str(tmp_path / "blocked" / "transforms.json")})
    assert bad["success"] is False and "render failed" in bad["error"]


def test_build_info_names_the_target_and_the_compiler(X):
    info = X._lib.library_info()
    assert "sm_100a" in info["build"] and "nvcc 12." in info["build"] and "DEVELOPMENT" not in info["build"]
    assert len(info["sha256"]) == 64 and info["bytes"] > 1 << 20 and info["path"].endswith("libcuda_render.so")
