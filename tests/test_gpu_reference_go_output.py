"""The CUDA path against OUTPUT OF THE REFERENCE'S GO BINARY: the image stored in the reference's demo notebook
(tests/golden/extract_notebook_image.py; tests/test_reference_go_output.py shows the oracle reproduces it bit for bit
and nothing nearby does).  Three 300 x 300 projections of cube_w_hole, R = 5, fov = 45, ds = 0.1, hierarchical."""
import struct
import zlib

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

R_NB, FOV_NB, DS_NB, RES_NB = 5.0, 45.0, 0.1, 300


@pytest.fixture(scope="module")
def stored():
    return np.load(GOLDEN / "reference_go_cube_w_hole_3x300.npz")["image"]


def grey_strip(X, imgs):
    return np.hstack([X.image_to_rgba8(np.asarray(im, dtype=np.float64))[..., 0] for im in imgs])


@pytest.mark.parametrize("path", ["span", "march"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_library_reproduces_the_go_binary(X, scenes, stored, monkeypatch, path, precision):
    """Both precision modes through both kernel families (interval renderer; marching kernels).  fp64 mode differs from
    the Go arithmetic only in libm's exp (<= 1e-15 in I); fp32 mode by <= 2e-7: a pixel can change grey level only when
    I * 65535 sits that close to a multiple of 256 -- a handful of the 52 067 attenuated pixels at most."""
    if path == "march":
        monkeypatch.setenv("XRAY_NO_SPAN", "1")
    else:
        monkeypatch.delenv("XRAY_NO_SPAN", raising=False)
    sc = X.Scene(str(scenes / "cube_w_hole.json"))
    cams = X.cameras_from_angles(X.generate_camera_angles(3), R_NB, FOV_NB)
    imgs = X.render_scene(sc, cams, RES_NB, ds=DS_NB, integration="hierarchical", precision=precision)
    got = grey_strip(X, imgs)
    d = np.abs(got.astype(int) - stored.astype(int))
    assert d.max() <= 1
    assert int((d != 0).sum()) <= (2 if precision == "fp64" else 25), int((d != 0).sum())


def test_renderer_mirror_reproduces_the_notebook_run(X, scenes, stored, tmp_path):
    """The notebook's own call (`XRayRenderer.render`, xray_renderer.py:357-448) through this package's drop-in
    class, to PNG files, decoded and pasted side by side as the notebook does."""
    out_dir = tmp_path / "images"
    res = X.XRayRenderer().render({"input": str(scenes / "cube_w_hole.json"), "num_images": 3, "resolution": RES_NB,
                                   "output_dir": str(out_dir), "fname_pattern": "image_%03d.png",
                                   "transforms_file": str(tmp_path / "transforms.json"),
                                   "R": R_NB, "fov": FOV_NB, "ds": DS_NB})
    assert res["success"] and res["num_images"] == 3
    tiles = []
    for k in range(3):
        data = (out_dir / f"image_{k:03d}.png").read_bytes()
        w, h, depth, ctype = struct.unpack(">IIBB", data[16:26])
        assert (w, h, depth, ctype) == (RES_NB, RES_NB, 8, 2)  # opaque *image.RGBA -> 8-bit RGB (Go's png.Encode)
        raw = zlib.decompress(data[data.index(b"IDAT") + 4:data.index(b"IEND") - 8])
        px = np.frombuffer(raw, dtype=np.uint8).reshape(h, 1 + w * 3)
        assert not px[:, 0].any()  # filter type 0 on every row
        rgb = px[:, 1:].reshape(h, w, 3)
        assert np.array_equal(rgb[..., 0], rgb[..., 1]) and np.array_equal(rgb[..., 0], rgb[..., 2])
        tiles.append(rgb[..., 0])  # PIL's RGB -> L of a grey pixel is that grey
    d = np.abs(np.hstack(tiles).astype(int) - stored.astype(int))
    assert d.max() <= 1 and int((d != 0).sum()) <= 25


def test_device_resident_path_reproduces_the_go_binary(X, scenes, stored):
    torch = pytest.importorskip("torch")
    sc = X.Scene(str(scenes / "cube_w_hole.json"))
    cams = X.cameras_from_angles(X.generate_camera_angles(3), R_NB, FOV_NB)
    out = torch.zeros((3, RES_NB, RES_NB), dtype=torch.float64, device="cuda")
    X.render_scene_device(sc, cams, RES_NB, out, ds=DS_NB, precision="fp64", stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    d = np.abs(grey_strip(X, out.cpu().numpy()).astype(int) - stored.astype(int))
    assert d.max() <= 1 and int((d != 0).sum()) <= 2


def test_go_goldens_when_present_through_the_cuda_path(X, scenes):
    """Frames made by the reference CLI on a machine with Go (tools/make_go_goldens.sh -> tests/golden/go/), compared
    directly with fp64-mode renders through the C ABI.  None can be made in the build image, so this skips there."""
    import json

    from test_reference_go_output import png_decode_grey

    checked = 0
    for c in json.loads((GOLDEN / "go" / "cases.json").read_text())["cases"]:
        files = [GOLDEN / "go" / f"{c['name']}_{k:03d}.png" for k in range(len(c["azimuthal"]))]
        if not all(f.exists() for f in files):
            continue
        deform = str(scenes / (c["deformation"].rsplit(".", 1)[0] + ".json")) if c["deformation"] else None
        sc = X.Scene(str(scenes / (c["input"].rsplit(".", 1)[0] + ".json")), deform)
        cams = X.cameras_from_angles(list(zip(c["azimuthal"], c["polar"])), c["R"], c["fov"])
        imgs = X.render_scene(sc, cams, c["resolution"], precision="fp64", ds=c["ds"], integration=c["integration"],
                              flat_field=c["flat_field"], density_multiplier=c["density_multiplier"])
        ndiff = worst = 0
        for f, im in zip(files, imgs):
            d = np.abs(X.image_to_rgba8(im)[..., 0].astype(int) - png_decode_grey(f.read_bytes()).astype(int))
            ndiff += int((d != 0).sum())
            worst = max(worst, int(d.max()))
        allowed = max(2, c["resolution"] ** 2 * len(c["azimuthal"]) // 500) if c["libm"] else 2
        assert ndiff <= allowed and worst <= 1, (c["name"], ndiff, worst)
        checked += 1
    if not checked:
        pytest.skip("no Go-made goldens under tests/golden/go/")
