"""Parity of the CUDA path (through the C ABI) against the CPU oracle -- the first gate.

Tolerances are the north star's: fp32 mode per-pixel |dI| <= 1e-4, fp64 mode <= 1e-9 (helpers.py).
Reference-equivalent sample counts must equal the oracle's density() call counts exactly.
"""
import ctypes
import json
import os

import numpy as np
import pytest

from helpers import FOV, R, TOL_FP32, TOL_FP64, assert_parity, gpu_vs_oracle, oracle_images

pytestmark = pytest.mark.gpu

ANALYTIC = ["cube_w_hole", "balls", "box_w_pped", "pillar_array", "lattice"]


@pytest.mark.parametrize("integ", ["hierarchical", "simple"])
@pytest.mark.parametrize("name", ANALYTIC)
def test_bundled_scenes(X, O, scenes, name, integ):
    out, nref, _ = gpu_vs_oracle(X, O, str(scenes / f"{name}.json"), integ=integ, res=48)
    assert_parity(out, nref)


def test_gyroid_with_sigmoid_warp(X, O, scenes):  # BASELINE config 3 (coarser step so the oracle finishes in seconds)
    out, nref, _ = gpu_vs_oracle(X, O, str(scenes / "gyroid_example.json"), str(scenes / "deformation_sigmoid.json"), res=32,
                                 ds=0.004)
    assert_parity(out, nref)


def test_gyroid_auto_step_single_rows(X, O, scenes):
    """config 3 at its real step (ds = 4e-4, 8700 coarse steps): a few pixel rows against the oracle."""
    obj, d = str(scenes / "gyroid_example.json"), str(scenes / "deformation_sigmoid.json")
    sc, osc = X.Scene(obj, d), O.OracleScene(obj, d)
    ds = sc.auto_ds()
    res = 16
    views = [(100.0, 90.0)]
    cams = X.cameras_from_angles(views, R, FOV)
    eye, cm = O.camera_from_angles(*views[0], R)
    ref, n = osc.render_view(eye, cm, res, FOV, R, ds, "hierarchical", rows=(6, 10))
    for prec, tol in (("fp32", TOL_FP32), ("fp64", TOL_FP64)):
        img = X.render_scene(sc, cams, res, precision=prec, ds=ds)
        assert np.abs(img[0, 6:10].astype(np.float64) - ref[6:10]).max() <= tol


DEFORMS = [
    {"type": "linear", "strains": [0.0, 0.0, 0.0, 0.1, 0.1, 0.1]},
    {"type": "rigid", "displacements": [0.1, -0.2, 0.05]},
    {"type": "sigmoid", "amplitude": -0.15, "center": 0.1, "lengthscale": 0.1, "direction": "x"},
    {"type": "affine", "matrix": [[1.1, 0.2, 0.0], [0.0, 0.9, 0.1], [0.3, 0.0, 1.2]]},
    {"type": "gaussian", "amplitudes": [0.1, 0.0, -0.1], "sigmas": [0.3, 0.4, 0.5], "centers": [0.1, 0.0, -0.2]},
    {"type": "composed", "deformations": [{"type": "rigid", "displacements": [0.1, 0.0, 0.0]},
                                          {"type": "sigmoid", "amplitude": 0.2, "center": 0.0, "lengthscale": 0.2, "direction": "y"},
                                          {"type": "linear", "strains": [0.01, 0.02, 0.03, 0.0, 0.0, 0.05]}]},
]


@pytest.mark.parametrize("d", DEFORMS, ids=lambda d: d["type"])
@pytest.mark.parametrize("name", ["cube_w_hole", "pillar_array"])
def test_deformations(X, O, scenes, name, d):
    obj = json.loads((scenes / f"{name}.json").read_text())
    out, nref, _ = gpu_vs_oracle(X, O, obj, d, res=40, ff=0.1, dm=1.7)
    assert_parity(out, nref)


def test_every_primitive_and_negative_densities(X, O):
    obj = {"type": "object_collection", "objects": [
        {"type": "box", "center": [0.0, 0.0, 0.0], "sides": [1.4, 1.2, 1.0], "rho": 0.6},
        {"type": "cube", "center": [0.3, 0.3, 0.3], "side": 0.5, "rho": 0.3},
        {"type": "sphere", "center": [-0.2, 0.1, 0.0], "radius": 0.25, "rho": -1.0},
        {"type": "cylinder", "p0": [-0.6, -0.5, -0.4], "p1": [0.5, 0.4, 0.45], "radius": 0.08, "rho": -0.6},
        {"type": "parallelepiped", "origin": [-0.1, -0.1, -0.1], "v0": [0.7, 0.0, 0.0], "v1": [0.1, 0.6, 0.0],
         "v2": [0.1, 0.1, 0.5], "rho": 0.2},
        {"type": "gyroid", "center": [0.0, 0.0, 0.0], "scale": 0.15, "thickness": 0.3, "rho": 0.1},
    ]}
    for greedy in (False, True):
        o = dict(obj, greedy_dens_eval=greedy)
        out, nref, _ = gpu_vs_oracle(X, O, o, res=40, ds=0.02)
        assert_parity(out, nref)


@pytest.mark.parametrize("prim", [
    {"type": "sphere", "center": [0.1, 0.0, -0.1], "radius": 0.5, "rho": 2.0},
    {"type": "cylinder", "p0": [0.0, 0.0, -0.7], "p1": [0.2, 0.1, 0.7], "radius": 0.2, "rho": -0.5},
    {"type": "gyroid", "center": [0.0, 0.0, 0.0], "scale": 0.2, "thickness": 0.25, "rho": 0.5},
], ids=lambda p: p["type"])
def test_bare_primitive_roots_are_not_clamped(X, O, prim):
    """A top-level primitive returns rho as is (2.0, negative ...): only collections clamp (objects.go:431-436)."""
    out, nref, _ = gpu_vs_oracle(X, O, prim, res=32, ds=0.02 if prim["type"] != "gyroid" else 0.01, views=((90.0, 90.0),))
    assert_parity(out, nref)


def test_near_cancelling_collection_sum(X, O):
    """rho values whose partial sums cancel: zero-ness of the sum decides refinement (main.go:181)."""
    obj = {"type": "object_collection", "objects": [
        {"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 0.6, "rho": 0.1},
        {"type": "sphere", "center": [0.1, 0.0, 0.0], "radius": 0.5, "rho": 0.2},
        {"type": "sphere", "center": [0.0, 0.1, 0.0], "radius": 0.45, "rho": -0.3},
        {"type": "box", "center": [0.0, 0.0, 0.0], "sides": [0.5, 0.5, 0.5], "rho": 0.7},
        {"type": "box", "center": [0.1, 0.1, 0.0], "sides": [0.4, 0.4, 0.4], "rho": -0.7},
    ]}
    out, nref, _ = gpu_vs_oracle(X, O, obj, res=48, ds=0.02)
    assert_parity(out, nref)


def test_generic_interpreter_nested_scene(X, O, scenes):
    """A collection holding a tessellation, a voxel grid and primitives: takes the generic interpreter."""
    lat = json.loads((scenes / "pillar_array.json").read_text())
    rng = np.random.default_rng(4)
    vol = rng.random((6, 5, 7))
    lat.update(xmin=-0.5, xmax=0.5, ymin=-0.5, ymax=0.5, zmin=-0.6, zmax=0.6)
    obj = {"type": "object_collection", "objects": [
        {"type": "sphere", "center": [0.6, 0.0, 0.0], "radius": 0.3, "rho": 0.4},
        lat,
        {"type": "box", "center": [-0.6, 0.1, 0.0], "sides": [0.3, 0.5, 0.7], "rho": -0.2},
        {"type": "voxel_grid", "_array": vol * 0.2},
    ]}
    out, nref, _ = gpu_vs_oracle(X, O, obj, res=32, ds=0.02)
    assert_parity(out, nref)
    os.environ["XRAY_GENERIC_KERNEL"] = "1"  # the bundled shapes through the generic kernels must agree too
    try:
        out, nref, _ = gpu_vs_oracle(X, O, str(scenes / "lattice.json"), res=32)
        assert_parity(out, nref)
    finally:
        del os.environ["XRAY_GENERIC_KERNEL"]


def test_known_answers_through_the_gpu(X, O):
    """The reference's own physics fixtures (main_test.go:326-437) as seen through a central pixel."""
    import math

    sph = {"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 0.5, "rho": 1.0}
    sc = X.Scene(sph)
    cams = X.cameras_from_angles([(90.0, 90.0)], 5.0, 40.0)
    res = 4  # pixel (2,2) is the optical axis: (i/(res/2)-1, j/(res/2)-1) = (0,0)
    for prec in ("fp32", "fp64"):
        img = X.render_scene(sc, cams, res, integration="simple", precision=prec, ds=0.001)
        assert abs(float(img[0, 2, 2]) - math.exp(-1.0)) <= 2 * 0.001            # TestIntegrateSimple_SphereCenterRay
        img = X.render_scene(sc, cams, res, integration="hierarchical", precision=prec, ds=0.05)
        assert abs(float(img[0, 2, 2]) - math.exp(-1.0)) <= 0.05                # TestIntegrateHierarchical_SphereCenterRay
        img = X.render_scene(sc, cams, res, integration="simple", precision=prec, ds=0.001, density_multiplier=2.0)
        assert abs(float(img[0, 2, 2]) - math.exp(-2.0)) <= 2 * 0.001 * 2.0    # TestDensityMultiplierApplied
    far = X.Scene({"type": "sphere", "center": [0.0, 0.0, 10.0], "radius": 0.01, "rho": 1.0})
    for integ in ("simple", "hierarchical"):                                     # TestFlatFieldAppliedOnce
        img = X.render_scene(far, cams, res, integration=integ, precision="fp64", ds=0.01, flat_field=1.0)
        assert np.abs(img - math.exp(-1.0)).max() <= 1e-9
        img = X.render_scene(far, cams, res, integration=integ, precision="fp32", ds=0.01, flat_field=1.0)
        assert np.abs(img.astype(np.float64) - math.exp(-1.0)).max() <= 1e-7


def test_ragged_sizes_and_mixed_cameras(X, O, scenes):
    """Detector sizes that are not tile multiples (scalar store path) and cameras with different R / fov."""
    obj = str(scenes / "cube_w_hole.json")
    sc, osc = X.Scene(obj), O.OracleScene(obj)
    ds = sc.auto_ds()
    for res in (1, 7, 33, 50):
        cams = X.cameras_from_angles([(77.0, 80.0)], R, FOV)
        eye, cm = O.camera_from_angles(77.0, 80.0, R)
        ref, _ = osc.render_view(eye, cm, res, FOV, R, ds, "hierarchical")
        img = X.render_scene(sc, cams, res, ds=ds)
        assert img.shape == (1, res, res)
        assert np.abs(img[0].astype(np.float64) - ref).max() <= TOL_FP32
    cams = (X._lib.XRayCameraParams64 * 3)()
    specs = [(10.0, 90.0, 4.0, 40.0), (200.0, 60.0, 3.0, 50.0), (300.0, 120.0, 4.0, 30.0)]
    for k, (az, pol, rr, fov) in enumerate(specs):
        cams[k] = X.camera_from_angles(az, pol, rr, fov)
    img = X.render_scene(sc, cams, 24, precision="fp64", ds=ds)
    for k, (az, pol, rr, fov) in enumerate(specs):
        eye, cm = O.camera_from_angles(az, pol, rr)
        ref, _ = osc.render_view(eye, cm, 24, fov, rr, ds, "hierarchical")
        assert np.abs(img[k] - ref).max() <= TOL_FP64


def test_empty_scene_and_miss(X, O):
    """Rays that miss everything, and a collection with nothing positive: image is exp(-flat_field) exactly."""
    import math

    neg = {"type": "object_collection", "objects": [{"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 0.5, "rho": -1.0}]}
    cams = X.cameras_from_angles([(0.0, 90.0)], R, FOV)
    for prec in ("fp32", "fp64"):
        img = X.render_scene(X.Scene(neg), cams, 16, precision=prec, ds=0.02, flat_field=0.25)
        assert np.abs(img.astype(np.float64) - math.exp(-0.25)).max() <= (1e-7 if prec == "fp32" else 1e-15)
    empty = {"type": "object_collection", "objects": []}
    img = X.render_scene(X.Scene(empty), cams, 8, precision="fp64", ds=0.02)
    assert (img == 1.0).all()


def test_error_paths_on_gpu(X, scenes):
    sc = X.Scene(str(scenes / "cube_w_hole.json"))
    cams = X.cameras_from_angles([(0.0, 90.0)], R, FOV)
    with pytest.raises(X._lib.XRayError):
        X.render_scene(sc, cams, 8, integration=7)
    inf = X.Scene({"type": "object_collection", "objects": []})  # MinFeatureSize = +Inf -> no auto step
    with pytest.raises(X._lib.XRayError, match="ds"):
        X.render_scene(inf, cams, 8)
    L = X._lib.load()
    o = X._lib.make_opts()
    o.struct_size = 4
    out = np.zeros((1, 8, 8), dtype=np.float32)
    assert L.XRayRenderSceneCUDA(sc.handle, cams, 1, 8, ctypes.byref(o), out.ctypes.data_as(ctypes.c_void_p)) == 2


def _foam(n_cells, rad, jitter=0.0):
    """Explicit object_collection of a Kelvin foam (objects.go:588-637 strut table), n_cells^3 cells in [-0.8, 0.8]^3."""
    from test_oracle_known_answers import kelvin

    size = 1.6 / n_cells
    uc = kelvin(rad, size)["objects"]["objects"]
    rng = np.random.default_rng(8)
    objs = []
    for a in range(n_cells):
        for b in range(n_cells):
            for c in range(n_cells):
                off = np.array([a, b, c]) * size - 0.8
                for o in uc:
                    objs.append({"type": "cylinder", "p0": list(np.array(o["p0"]) + off + rng.normal(0, jitter, 3)),
                                 "p1": list(np.array(o["p1"]) + off + rng.normal(0, jitter, 3)), "radius": rad,
                                 "rho": float(rng.choice([1.0, 0.6, -0.4])) if jitter else 1.0})
    return objs


@pytest.mark.parametrize("greedy", [False, True])
def test_big_collections_use_cell_lists(X, O, greedy, kernel_path):
    """More than 63 children: no 64-bit child mask; per-cell ascending child lists merged across the warp keep
    the reference's summation / greedy order (objects.go:422-438).  Mixed-sign rho makes order matter."""
    objs = _foam(2, 0.03, jitter=0.01) + [{"type": "sphere", "center": [0.1, 0.0, -0.2], "radius": 0.25, "rho": 0.5},
                                           {"type": "box", "center": [-0.3, 0.2, 0.3], "sides": [0.3, 0.2, 0.4], "rho": -0.7}]
    obj = {"type": "object_collection", "objects": objs, "greedy_dens_eval": greedy}
    out, nref, _ = gpu_vs_oracle(X, O, obj, res=32, ds=0.01, views=((100.0, 80.0), (10.0, 95.0)))
    assert_parity(out, nref)
    # the lists really are in use: far fewer primitive tests than brute force
    assert out["fp32"][1]["primitive_tests"] < 0.05 * out["fp32"][1]["evaluated_samples"] * len(objs)


@pytest.mark.parametrize("n", [62, 63, 64, 65])
@pytest.mark.parametrize("wrap", ["flat", "tess"])
def test_collection_sizes_around_the_mask_width(X, O, n, wrap, kernel_path):
    """63 / 64 / 65 children: the marching kernels' 64-bit child masks keep bit 63 for "empty cell, skip distance", so a
    collection of exactly 64 must already take the cell lists (it took the masks in r1 and child 63 vanished wherever it
    was listed: found by soaking tests/test_gpu_fuzz.py over 1200 more seeds); the interval renderer takes up to 64."""
    rng = np.random.default_rng(6400 + n)
    objs = []
    for k in range(n):
        c = rng.uniform(-0.45, 0.45, 3)
        kind = k % 4
        rho = float(rng.choice([1.0, 0.6, -0.4, 0.3]))
        if kind == 0:
            objs.append({"type": "sphere", "center": list(c), "radius": float(rng.uniform(0.04, 0.12)), "rho": rho})
        elif kind == 1:
            objs.append({"type": "box", "center": list(c), "sides": list(rng.uniform(0.05, 0.25, 3)), "rho": rho})
        elif kind == 2:
            objs.append({"type": "cylinder", "p0": list(c), "p1": list(c + rng.uniform(-0.3, 0.3, 3)), "radius": float(rng.uniform(0.02, 0.08)), "rho": rho})
        else:
            m = np.eye(3) * rng.uniform(0.08, 0.2, 3) + rng.uniform(-0.04, 0.04, (3, 3))
            objs.append({"type": "parallelepiped", "origin": list(c), "v0": list(m[0]), "v1": list(m[1]), "v2": list(m[2]), "rho": rho})
    objs[-1] = {"type": "sphere", "center": [0.05, -0.1, 0.1], "radius": 0.3, "rho": 0.45}  # the LAST child is the big one
    if wrap == "flat":
        obj = {"type": "object_collection", "objects": objs, "greedy_dens_eval": bool(n % 2)}
    else:
        uc = {"objects": {"objects": objs}, "xmin": -0.5, "xmax": 0.5, "ymin": -0.5, "ymax": 0.5, "zmin": -0.5, "zmax": 0.5}
        obj = {"type": "tessellated_obj_coll", "uc": uc, "xmin": -0.9, "xmax": 0.9, "ymin": -0.8, "ymax": 0.8, "zmin": -0.7, "zmax": 0.7}
    for integ in ("hierarchical", "simple"):
        out, nref, _ = gpu_vs_oracle(X, O, obj, res=32, ds=0.01, integ=integ, views=((100.0, 80.0), (10.0, 95.0)))
        assert_parity(out, nref)


def test_big_unit_cell_collection(X, O):
    """A tessellated unit cell with > 63 struts (cell lists under the periodic fold + skip distances)."""
    uc_objs = _foam(2, 0.02)
    for o in uc_objs:  # move the foam into a [0, 1.6]^3 unit cell
        o["p0"] = [v + 0.8 for v in o["p0"]]
        o["p1"] = [v + 0.8 for v in o["p1"]]
    uc = {"objects": {"objects": uc_objs}, "xmin": 0.0, "xmax": 1.6, "ymin": 0.0, "ymax": 1.6, "zmin": 0.0, "zmax": 1.6}
    obj = {"type": "tessellated_obj_coll", "uc": uc, "xmin": -0.9, "xmax": 0.9, "ymin": -0.9, "ymax": 0.9, "zmin": -0.9, "zmax": 0.9}
    out, nref, _ = gpu_vs_oracle(X, O, obj, res=32, ds=0.01, views=((200.0, 70.0),))
    assert_parity(out, nref)


def test_skipping_changes_nothing(X, scenes, kernel_path):
    """Empty-space / in-wall skipping must be invisible: same images (to fp32 summation order), same
    reference-equivalent sample counts, fewer evaluated samples."""
    for name, deform in (("lattice", None), ("pillar_array", None), ("gyroid_example", "deformation_sigmoid")):
        sc = X.Scene(str(scenes / f"{name}.json"), str(scenes / f"{deform}.json") if deform else None)
        cams = X.cameras_from_angles([(33.0, 90.0), (140.0, 70.0)], R, FOV)
        ds = 0.004 if name == "gyroid_example" else -1.0
        a, sa = X.render_scene(sc, cams, 96, ds=ds, return_stats=True)
        os.environ["XRAY_NO_SKIP"] = "1"
        try:
            b, sb = X.render_scene(sc, cams, 96, ds=ds, return_stats=True)
        finally:
            del os.environ["XRAY_NO_SKIP"]
        assert np.abs(a.astype(np.float64) - b).max() <= 2e-6
        assert sa["ref_samples"] == sb["ref_samples"]
        if kernel_path == "march" or name == "gyroid_example":  # (the interval renderer evaluates no samples at all)
            assert sa["evaluated_samples"] < sb["evaluated_samples"]


def test_degenerate_cameras_are_refused(X, scenes):
    """polar = 0 makes LookAtV singular (main.go:236-237): the camera matrix is all zero (mgl64 Inv of a singular
    matrix) or NaN.  The reference would write garbage images; the plugin returns error 2 instead of marching."""
    sc = X.Scene(str(scenes / "cube_w_hole.json"))
    cam = X.camera_from_angles(30.0, 0.0, 4.0, 40.0)
    cams = (X._lib.XRayCameraParams64 * 1)(cam)
    with pytest.raises(X._lib.XRayError, match="degenerate"):
        X.render_scene(sc, cams, 8)
    good = X.camera_from_angles(30.0, 90.0, 4.0, 40.0)
    good.fov_y = 190.0
    with pytest.raises(X._lib.XRayError, match="degenerate"):
        X.render_scene(sc, (X._lib.XRayCameraParams64 * 1)(good), 8)


@pytest.mark.parametrize("Rcam,fov,promoted", [(40.0, 40.0, True), (40.0, 4.0, False), (4.0, 120.0, False), (1.9, 60.0, False)])
def test_far_wide_and_near_cameras(X, O, scenes, Rcam, fov, promoted, kernel_path):
    """fp32 mode is only sound while |o + d*R| keeps the fp32 position error under the guard-band budget; a distant
    camera must be rendered by the fp64 kernels instead (api.cu fp32_position_bound_ok).  Seen from outside: the
    fp32-mode image then agrees with the oracle to float rounding and nothing is skipped.  Wide and near cameras
    (eye inside the sample window) stay on the fp32 kernels and must hold the fp32 tolerance."""
    obj = str(scenes / "lattice.json")
    sc, osc = X.Scene(obj), O.OracleScene(obj)
    ds, res = 0.02, (192 if promoted else 24)
    views = [(100.0, 80.0)]
    cams = X.cameras_from_angles(views, Rcam, fov)
    ref, nref = oracle_images(O, osc, views, res, ds, "hierarchical", R=Rcam, fov=fov)
    img, st = X.render_scene(sc, cams, res, precision="fp32", ds=ds, return_stats=True)
    assert st["ref_samples"] == nref
    if promoted:
        assert np.abs(img.astype(np.float64) - ref).max() <= 1e-6
        assert st["fp64_fallbacks"] == 0  # the exact kernel has no guard band
    else:
        assert np.abs(img.astype(np.float64) - ref).max() <= TOL_FP32
        assert st["evaluated_samples"] < nref  # culled / skipped samples are not evaluated (span: intervals, not samples)


@pytest.mark.parametrize("name,deform,az", [("lattice", None, 90.0), ("lattice", None, 45.0), ("pillar_array", None, 0.0),
                                            ("lattice", "deformation_linear", 90.0), ("gyroid_example", "deformation_sigmoid", 90.0)])
def test_rays_inside_cell_face_planes(X, O, scenes, name, deform, az, kernel_path):
    """polar = 90 deg puts the central pixel row in the plane z = 0, which is a unit-cell face of the lattice and the
    pillar array: every sample of those rays sits on the fold discontinuity and the period is decided by rounding
    noise of the fp64 reference arithmetic.  fp32 mode must reproduce it through the exact-fold cold path
    (render_fast.cu exact_fold_cold); checked on the central row and its neighbours at full benchmark resolution."""
    obj = str(scenes / f"{name}.json")
    d = str(scenes / f"{deform}.json") if deform else None
    sc, osc = X.Scene(obj, d), O.OracleScene(obj, d)
    ds = sc.auto_ds() if name != "gyroid_example" else 0.004
    res = 256
    cams = X.cameras_from_angles([(az, 90.0)], R, FOV)
    eye, cm = O.camera_from_angles(az, 90.0, R)
    rows = (res // 2 - 1, res // 2 + 2)
    ref, _ = osc.render_view(eye, cm, res, FOV, R, ds, "hierarchical", rows=rows)
    img, st = X.render_scene(sc, cams, res, precision="fp32", ds=ds, return_stats=True)
    # oracle rows index i (camera x); the degenerate direction is j = res/2 (camera y = 0): check both cuts
    assert np.abs(img[0, rows[0]:rows[1]].astype(np.float64) - ref[rows[0]:rows[1]]).max() <= TOL_FP32
    refT, _ = osc.render_view(eye, cm, res, FOV, R, ds, "hierarchical")
    assert np.abs(img[0][:, res // 2 - 1:res // 2 + 2].astype(np.float64) - refT[:, res // 2 - 1:res // 2 + 2]).max() <= TOL_FP32
    if kernel_path == "march" or name == "gyroid_example":
        assert st["fp64_fallbacks"] > 0
    elif deform is None:
        # interval renderer: the face-plane rays are settled inside it (span_degenerate_axis), not handed to the marching kernels
        assert st["marched_tiles"] <= 2, st


# ---- one-primitive scenes: the lane-asynchronous kernel (render_fast.cu render_async_kernel) ----
GYROID_CELL = {"type": "tessellated_obj_coll", "xmin": -0.7, "xmax": 0.7, "ymin": -0.6, "ymax": 0.6, "zmin": -0.65, "zmax": 0.65,
               "uc": {"xmin": -1.0, "xmax": 1.0, "ymin": -1.0, "ymax": 1.0, "zmin": -1.0, "zmax": 1.0,
                      "objects": {"objects": [{"type": "gyroid", "center": [0.05, 0.0, -0.02], "scale": 0.12, "thickness": 0.25, "rho": 0.8}]}}}


@pytest.mark.parametrize("d", [None] + DEFORMS, ids=lambda d: d["type"] if d else "none")
def test_gyroid_cell_under_every_warp(X, O, d):
    """The gyroid skip bound uses the warp's Jacobian and curvature (second-order rule for none / rigid / linear / affine /
    sigmoid; first-order Lipschitz rule for gaussian and composed chains): each must stay within the tolerances and keep
    the oracle's sample count."""
    out, nref, _ = gpu_vs_oracle(X, O, GYROID_CELL, d, res=32, ds=0.002, views=((77.0, 90.0), (205.0, 62.0)))
    assert_parity(out, nref)


@pytest.mark.parametrize("integ", ["hierarchical", "simple"])
def test_unbounded_gyroid_fine_step(X, O, integ):
    """A bare gyroid fills the whole integration window: rays end inside a wall, and an in-wall skip must stop at the
    last lattice sample instead of adding weight for samples the reference never takes."""
    prim = {"type": "gyroid", "center": [0.0, 0.0, 0.0], "scale": 0.2, "thickness": 0.3, "rho": 0.5}
    out, nref, _ = gpu_vs_oracle(X, O, prim, res=24, ds=0.001, integ=integ, views=((90.0, 90.0), (10.0, 50.0)))
    assert_parity(out, nref)
    coll = {"type": "object_collection", "objects": [dict(prim, rho=1.5)]}  # one child: clamped to 1 unless greedy
    for greedy in (False, True):
        out, nref, _ = gpu_vs_oracle(X, O, dict(coll, greedy_dens_eval=greedy), res=24, ds=0.001, integ=integ, views=((120.0, 80.0),))
        assert_parity(out, nref)


def test_async_and_lockstep_kernels_agree(X, scenes):
    """XRAY_NO_ASYNC routes one-primitive scenes through the warp-synchronous kernel: same images up to fp32 summation
    order, same reference-equivalent sample counts."""
    cases = [("pillar_array", None, -1.0, "hierarchical"), ("pillar_array", "deformation_linear", -1.0, "simple"),
             ("gyroid_example", "deformation_sigmoid", 0.002, "hierarchical"), ("gyroid_example", None, 0.002, "simple")]
    cams = X.cameras_from_angles([(33.0, 90.0), (140.0, 70.0)], R, FOV)
    for name, deform, ds, integ in cases:
        sc = X.Scene(str(scenes / f"{name}.json"), str(scenes / f"{deform}.json") if deform else None)
        a, sa = X.render_scene(sc, cams, 96, ds=ds, integration=integ, return_stats=True)
        os.environ["XRAY_NO_ASYNC"] = "1"
        try:
            b, sb = X.render_scene(sc, cams, 96, ds=ds, integration=integ, return_stats=True)
        finally:
            del os.environ["XRAY_NO_ASYNC"]
        assert np.abs(a.astype(np.float64) - b).max() <= 2e-6, (name, deform)
        assert sa["ref_samples"] == sb["ref_samples"]


@pytest.mark.parametrize("prim", [
    {"type": "sphere", "center": [0.2, 0.25, 0.1], "radius": 0.18, "rho": 0.9},
    {"type": "box", "center": [0.25, 0.2, 0.0], "sides": [0.2, 0.3, 0.8], "rho": 1.0},
    {"type": "cylinder", "p0": [0.1, 0.1, -0.6], "p1": [0.4, 0.35, 0.6], "radius": 0.07, "rho": 0.7},
], ids=lambda p: p["type"])
def test_one_primitive_unit_cells(X, O, prim):
    """Tessellations of a single sphere / box / (tilted) cylinder: grid-driven empty-space skipping per lane."""
    obj = {"type": "tessellated_obj_coll", "xmin": -0.9, "xmax": 0.9, "ymin": -0.8, "ymax": 0.8, "zmin": -0.7, "zmax": 0.7,
           "uc": {"xmin": 0.0, "xmax": 0.5, "ymin": 0.0, "ymax": 0.5, "zmin": -1.0, "zmax": 1.0, "objects": {"objects": [prim]}}}
    for integ in ("hierarchical", "simple"):
        out, nref, _ = gpu_vs_oracle(X, O, obj, res=40, ds=0.01, integ=integ)
        assert_parity(out, nref)
