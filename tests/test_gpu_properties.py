"""Size-independent properties of the hot path at BASELINE-sized detectors, where the oracle can only
afford random pixels: view-sharding equivalence, determinism, batch invariance, Beer-Lambert
multiplicativity, device-resident vs host path identity, multi-GPU equivalence, plus the
end-to-end renderer mirror and the committed golden vectors."""
import json
import math
import os

import numpy as np
import pytest

from helpers import FOV, R, TOL_FP32, TOL_FP64

pytestmark = pytest.mark.gpu


def test_full_size_lattice_random_pixels(X, O, scenes):
    """BASELINE config 2 geometry at its full 1024x1024 detector: 4 of the 360 views, 160 random pixels each."""
    obj = str(scenes / "lattice.json")
    sc, osc = X.Scene(obj), O.OracleScene(obj)
    ds = sc.auto_ds()
    res = 1024
    angles = X.generate_camera_angles(360)
    picks = [angles[k] for k in (0, 97, 181, 333)]
    cams = X.cameras_from_angles(picks, R, FOV)
    img = X.render_scene(sc, cams, res, ds=ds)
    rng = np.random.default_rng(5)
    for v, a in enumerate(picks):
        ij = rng.integers(256, 768, size=(160, 2))
        eye, cm = O.camera_from_angles(a["azimuthal"], a["polar"], R)
        ref, _ = osc.render_pixels(eye, cm, res, FOV, R, ds, ij)
        got = img[v][ij[:, 0], ij[:, 1]].astype(np.float64)
        assert np.abs(got - ref).max() <= TOL_FP32
        assert ref.min() < 0.9  # the sample really crosses the lattice


def test_full_size_gyroid_sigmoid_random_pixels(X, O, scenes):
    """BASELINE config 3 as specified (gyroid + sigmoid warp, 1024x1024 detector, ds = 4e-4: 8700 coarse steps per ray):
    4 of the 720 views, 4000 random pixels each -- 1.4e8 density() calls through the oracle, every one of which the
    lane-asynchronous kernel has to classify like the reference (second-order skips, guard band at 1.1x its bound)."""
    obj, d = str(scenes / "gyroid_example.json"), str(scenes / "deformation_sigmoid.json")
    sc, osc = X.Scene(obj, d), O.OracleScene(obj, d)
    ds = sc.auto_ds()
    res = 1024
    angles = X.generate_camera_angles(720)
    picks = [angles[k] for k in (3, 250, 433, 611)]
    cams = X.cameras_from_angles(picks, R, FOV)
    img = X.render_scene(sc, cams, res, ds=ds)
    rng = np.random.default_rng(11)
    worst = 0.0
    for v, a in enumerate(picks):
        ij = rng.integers(200, 824, size=(4000, 2))
        eye, cm = O.camera_from_angles(a["azimuthal"], a["polar"], R)
        ref, _ = osc.render_pixels(eye, cm, res, FOV, R, ds, ij)
        got = img[v][ij[:, 0], ij[:, 1]].astype(np.float64)
        worst = max(worst, float(np.abs(got - ref).max()))
        assert ref.min() < 0.9
    assert worst <= TOL_FP32, worst


def test_full_size_pillar_random_pixels(X, O, scenes):
    """BASELINE config 5 geometry at its full 4096x4096 detector: one view, random pixels."""
    obj = str(scenes / "pillar_array.json")
    sc, osc = X.Scene(obj), O.OracleScene(obj)
    ds = sc.auto_ds()
    res = 4096
    cams = X.cameras_from_angles([(123.0, 90.0)], R, FOV)
    img = X.render_scene(sc, cams, res, ds=ds)
    ij = np.random.default_rng(6).integers(0, res, size=(400, 2))
    eye, cm = O.camera_from_angles(123.0, 90.0, R)
    ref, _ = osc.render_pixels(eye, cm, res, FOV, R, ds, ij)
    assert np.abs(img[0][ij[:, 0], ij[:, 1]].astype(np.float64) - ref).max() <= TOL_FP32


def test_view_sharding_and_batch_invariance(X, scenes):
    """Rendering views in shards (--jobs_modulo semantics) or one by one gives bit-identical images."""
    sc = X.Scene(str(scenes / "lattice.json"))
    angles = X.generate_camera_angles(12)
    cams = X.cameras_from_angles(angles, R, FOV)
    full = X.render_scene(sc, cams, 96)
    for world in (2, 3):
        for r in range(world):
            part = X.render_scene(sc, X.cameras_from_angles(X.generate_camera_angles(12, r, world), R, FOV), 96)
            assert np.array_equal(part, full[r::world])
    single = X.render_scene(sc, X.cameras_from_angles([angles[7]], R, FOV), 96)
    assert np.array_equal(single[0], full[7])
    again = X.render_scene(sc, cams, 96)
    assert np.array_equal(again, full)  # deterministic


def test_counting_kernel_variant_is_identical(X, scenes):
    sc = X.Scene(str(scenes / "pillar_array.json"))
    cams = X.cameras_from_angles(X.generate_camera_angles(3), R, FOV)
    a = X.render_scene(sc, cams, 64)
    b, st = X.render_scene(sc, cams, 64, return_stats=True)
    assert np.array_equal(a, b)
    assert st["rays"] == 3 * 64 * 64 and st["ref_samples"] >= st["rays"] * 174 and st["launches"] >= 1


def test_beer_lambert_multiplicativity(X, scenes):
    """T is linear in density_multiplier and additive in flat_field for the fixed-step integrator:
    I(dm=2) = I(dm=1)^2, I(ff) = I(0)*exp(-ff)."""
    sc = X.Scene(str(scenes / "cube_w_hole.json"))
    cams = X.cameras_from_angles([(60.0, 90.0)], R, FOV)
    i1 = X.render_scene(sc, cams, 128, integration="simple", precision="fp64")
    i2 = X.render_scene(sc, cams, 128, integration="simple", precision="fp64", density_multiplier=2.0)
    iff = X.render_scene(sc, cams, 128, integration="simple", precision="fp64", flat_field=0.3)
    assert np.abs(i2 - i1 * i1).max() <= 1e-12
    assert np.abs(iff - i1 * math.exp(-0.3)).max() <= 1e-12
    assert i1.min() < 0.5 and i1.max() == 1.0


def test_fp32_and_fp64_modes_agree(X, scenes):
    for name in ("lattice", "pillar_array", "cube_w_hole"):
        sc = X.Scene(str(scenes / f"{name}.json"))
        cams = X.cameras_from_angles([(10.0, 90.0), (100.0, 80.0)], R, FOV)
        a = X.render_scene(sc, cams, 192, precision="fp32")
        b = X.render_scene(sc, cams, 192, precision="fp64")
        assert a.dtype == np.float32 and b.dtype == np.float64
        assert np.abs(a.astype(np.float64) - b).max() <= 2e-6  # rounding only: every classification matches


def test_device_resident_equals_host_path(X, scenes):
    torch = pytest.importorskip("torch")
    sc = X.Scene(str(scenes / "lattice.json"))
    cams = X.cameras_from_angles(X.generate_camera_angles(5), R, FOV)
    out = torch.zeros((5, 80, 80), dtype=torch.float32, device="cuda")
    X.render_scene_device(sc, cams, 80, out, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = X.render_scene(sc, cams, 80)
    assert np.array_equal(out.cpu().numpy(), host)
    pinned = torch.empty((5, 80, 80), dtype=torch.float32, pin_memory=True)
    X.render_scene(sc, cams, 80, out=pinned.numpy())
    assert np.array_equal(pinned.numpy(), host)
    out64 = torch.zeros((5, 80, 80), dtype=torch.float64, device="cuda")
    X.render_scene_device(sc, cams, 80, out64, precision="fp64", stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out64.cpu().numpy(), X.render_scene(sc, cams, 80, precision="fp64"))


def test_pageable_output_at_any_alignment(X, scenes):
    """The drain into the caller's pageable buffer (multi-threaded, streaming stores with a memcpy head and tail)
    delivers the same bytes whatever the buffer's alignment, and writes nothing outside it (cuda_backend.go:294-337
    hands over Go-heap slices)."""
    sc = X.Scene(str(scenes / "balls.json"))
    nv, res = 3, 1024  # 12.6 MB: three copy threads, every piece above the streaming threshold
    cams = X.cameras_from_angles(X.generate_camera_angles(nv), R, FOV)
    dev_ref = X.render_scene(sc, cams, res)
    n = nv * res * res
    for off in (1, 2, 3, 5):  # float32 elements: 4, 8, 12, 20 bytes past a 64-byte boundary
        raw = np.full(n + 64, -7.0, dtype=np.float32)
        base = (-raw.ctypes.data // 4) % 16  # element index of a 64-byte boundary
        out = raw[base + off:base + off + n].reshape(nv, res, res)
        assert out.ctypes.data % 64 == 4 * off
        X.render_scene(sc, cams, res, out=out)
        assert np.array_equal(out, dev_ref)
        assert np.all(raw[:base + off] == -7.0) and np.all(raw[base + off + n:] == -7.0)
    old = os.environ.get("XRAY_NO_STREAM_COPY")
    os.environ["XRAY_NO_STREAM_COPY"] = "1"
    try:
        assert np.array_equal(X.render_scene(sc, cams, res), dev_ref)
    finally:
        if old is None:
            del os.environ["XRAY_NO_STREAM_COPY"]
        else:
            os.environ["XRAY_NO_STREAM_COPY"] = old


def test_multi_gpu_view_sharding_in_one_call(X, scenes):
    n = X._lib.load().XRayDeviceCount()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    sc = X.Scene(str(scenes / "lattice.json"))
    cams = X.cameras_from_angles(X.generate_camera_angles(9), R, FOV)
    one = X.render_scene(sc, cams, 64)
    many = X.render_scene(sc, cams, 64, devices=list(range(min(n, 4))))
    assert np.array_equal(one, many)
    vol = np.random.default_rng(2).random((16, 16, 16), dtype=np.float32)
    a = X.render_volume(vol, cams, 32, ds=0.025)
    b = X.render_volume(vol, cams, 32, ds=0.025, devices=[0, 1])
    assert np.array_equal(a, b)


def test_golden_vectors(X, scenes):
    """Committed oracle outputs (tests/golden/make_golden.py): pins GPU results across rounds."""
    from conftest import GOLDEN

    meta = json.loads((GOLDEN / "golden_v1.json").read_text())
    data = np.load(GOLDEN / "golden_v1.npz")
    for case in meta["cases"]:
        sc = X.Scene(str(scenes / case["object"]), str(scenes / case["deformation"]) if case["deformation"] else None)
        cams = X.cameras_from_angles([tuple(v) for v in case["views"]], case["R"], case["fov"])
        ref = data[case["key"]]
        for prec, tol in (("fp32", TOL_FP32), ("fp64", TOL_FP64)):
            img = X.render_scene(sc, cams, case["res"], integration=case["integration"], precision=prec, ds=case["ds"],
                                 flat_field=case["flat_field"], density_multiplier=case["density_multiplier"])
            assert np.abs(img.astype(np.float64) - ref).max() <= tol, (case["key"], prec)


def test_renderer_end_to_end_files(X, O, scenes, tmp_path):
    """XRayRenderer.render (xray_renderer.py:357-448 contract): PNG frames, transforms.json, object.json."""
    out_dir = tmp_path / "run" / "images"
    r = X.XRayRenderer()
    res = r.render({"input": str(scenes / "cube_w_hole.json"), "output_dir": str(out_dir), "resolution": 64, "num_images": 3,
                    "transforms_file": str(tmp_path / "run" / "transforms.json"), "flat_field": 0.1, "time_label": 2.5})
    assert res["success"] and res["num_images"] == 3
    tf = json.loads((tmp_path / "run" / "transforms.json").read_text())
    assert tf["w"] == tf["h"] == 64 and tf["cx"] == 32.0 and len(tf["frames"]) == 3
    assert tf["flat_field"] == math.exp(-0.1) and tf["camera_angle_x"] == 40.0 * math.pi / 180.0   # main.go:380-388
    f = 1 / math.tan((40.0 / 2) * math.pi / 180.0)
    assert tf["fl_x"] == f * 64.0 / 2.0
    assert tf["frames"][1]["file_path"] == "images/image_001.png" and tf["frames"][1]["time"] == 2.5
    eye, cm = O.camera_from_angles(90.0 + 120.0, 90.0, 4.0)
    assert np.array_equal(np.array(tf["frames"][1]["transform_matrix"]), cm)
    assert json.loads((tmp_path / "run" / "object.json").read_text())["type"] == "object_collection"
    # frame 1 decodes to the oracle image quantised as main.go:495-498 (allow 1 grey level for fp32 rounding)
    import zlib, struct
    data = (out_dir / "image_001.png").read_bytes()
    raw = zlib.decompress(data[data.index(b"IDAT") + 4:data.index(b"IEND") - 8])
    assert data[25] == 2  # opaque frame: 8-bit RGB, as Go's png.Encode writes an opaque *image.RGBA
    px = np.frombuffer(raw, dtype=np.uint8).reshape(64, 1 + 64 * 3)[:, 1:].reshape(64, 64, 3)
    osc = O.OracleScene(str(scenes / "cube_w_hole.json"), flat_field=0.1)
    ref, _ = osc.render_view(eye, cm, 64, 40.0, 4.0, osc.auto_ds(), "hierarchical")
    want = X.image_to_rgba8(ref)[..., :3]
    assert np.abs(px.astype(int) - want.astype(int)).max() <= 1
    assert (px == want).mean() > 0.999
    bad = r.render({"input": str(tmp_path / "missing.json"), "output_dir": str(out_dir)})
    assert bad["success"] is False and "render failed" in bad["error"]
