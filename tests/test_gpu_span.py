"""The interval ("span") renderer (csrc/render_span.cu): scenes of convex primitives are rendered by intersecting each ray with
each candidate child in fp64 and sweeping the interval end points over the reference's sample lattice.  These cases aim at
what is specific to it: tangent and face-parallel rays, rays inside cell-face planes, more intervals than its per-ray lists
hold (hand-over to the marching kernels), warps that keep rays straight, and that it really is the path taken."""
import numpy as np
import pytest

from helpers import FOV, R, TOL_FP32, TOL_FP64, assert_parity, gpu_vs_oracle

pytestmark = pytest.mark.gpu


def _check(X, O, obj, deform=None, *, views, res, ds=-1.0, integs=("hierarchical", "simple"), dm=1.0, ff=0.0, max_marched=None):
    worst = 0.0
    for integ in integs:
        out, nref, _ = gpu_vs_oracle(X, O, obj, deform, views=views, res=res, integ=integ, ds=ds, dm=dm, ff=ff)
        assert_parity(out, nref)
        for prec in ("fp32", "fp64"):
            st = out[prec][1]
            assert st["span_renderer"] and st["launches"] >= 3, "interval renderer (fast pass, settle pass) + hand-over pass expected"
            if max_marched is not None:
                assert st["marched_tiles"] <= max_marched, (st["marched_tiles"], hex(st["march_reasons"]), integ, prec)
            worst = max(worst, out[prec][0])
    return worst


def test_span_is_the_default_path_for_convex_scenes(X, O, scenes):
    for name in ("cube_w_hole", "lattice", "pillar_array", "balls", "box_w_pped"):
        _check(X, O, str(scenes / f"{name}.json"), views=((77.0, 83.0), (200.0, 101.0)), res=40, max_marched=2)  # (the central ray meets cap planes / sphere centres exactly on a lattice sample)
    # not eligible: gyroid (not convex), sigmoid warp (bends rays) -> one marching launch
    sc = X.Scene(str(scenes / "lattice.json"), str(scenes / "deformation_sigmoid.json"))
    cams = X.cameras_from_angles([(77.0, 83.0)], R, FOV)
    _, st = X.render_scene(sc, cams, 24, return_stats=True)
    assert not st["span_renderer"] and st["marched_tiles"] == 0


@pytest.mark.parametrize("warp", [
    {"type": "rigid", "displacements": [0.05, -0.11, 0.07]},
    {"type": "linear", "strains": [0.08, -0.05, 0.03, 0.02, -0.04, 0.06]},
    {"type": "affine", "matrix": [[0.9, 0.2, -0.1], [0.05, 1.1, 0.15], [-0.2, 0.1, 0.95]]},
    {"type": "composed", "deformations": [{"type": "linear", "strains": [0.1, 0.0, -0.05, 0.0, 0.03, 0.0]},
                                          {"type": "rigid", "displacements": [0.1, 0.0, -0.05]},
                                          {"type": "affine", "matrix": [[1.0, 0.1, 0.0], [0.0, 1.0, 0.1], [0.1, 0.0, 1.0]]}]},
])
def test_affine_warps_keep_rays_straight(X, O, scenes, warp):
    """deformations.go:87-92, 136-141, 173-175, 269-274: the warp is applied to the ray once instead of to every sample."""
    for name in ("lattice", "cube_w_hole", "box_w_pped"):
        _check(X, O, str(scenes / f"{name}.json"), warp, views=((90.0, 90.0), (141.0, 66.0)), res=36, dm=1.4, ff=0.02, max_marched=8)


def test_tangent_and_axis_parallel_rays(X, O):
    """The central ray (az 0, polar 90: the x axis) is tangent to a sphere, runs inside a box face, along a cylinder's axis
    and through the plane of its cap; the central row and column are parallel to box faces.  Whatever the interval renderer
    cannot decide it must hand over, and the image must still match."""
    obj = {"type": "object_collection", "objects": [
        {"type": "sphere", "center": [0.0, 0.3, 0.0], "radius": 0.3, "rho": 0.5},          # tangent to the x axis
        {"type": "box", "center": [0.4, -0.2, 0.0], "sides": [0.3, 0.4, 0.5], "rho": 0.4},  # face y = 0 contains it
        {"type": "cylinder", "p0": [-0.6, 0.0, 0.0], "p1": [-0.2, 0.0, 0.0], "radius": 0.1, "rho": 0.3},  # coaxial
        {"type": "cylinder", "p0": [0.0, -0.5, 0.0], "p1": [0.0, -0.5, 0.4], "radius": 0.2, "rho": -0.2},  # cap in z = 0
        {"type": "parallelepiped", "origin": [-0.5, 0.2, -0.3], "v0": [0.3, 0.0, 0.0], "v1": [0.0, 0.3, 0.0], "v2": [0.1, 0.0, 0.3],
         "rho": 0.6}]}
    for greedy in (False, True):
        obj["greedy_dens_eval"] = greedy
        _check(X, O, obj, views=((0.0, 90.0), (90.0, 90.0), (45.0, 90.0)), res=65, ds=0.013)


def test_more_intervals_than_the_lists_hold(X, O):
    """60 thin plates in a row: rays along the row cross more intervals than the per-ray lists of the interval renderer hold.
    The settle pass renders such rays in windows of the lattice (sweep state carried across windows); under a warp nothing
    is settled there, so those tiles go to the marching kernels.  Exact either way."""
    plates = [{"type": "box", "center": [-0.885 + 0.03 * k, 0.0, 0.0], "sides": [0.012, 0.8, 0.8], "rho": 0.05 + 0.01 * (k % 7)}
              for k in range(60)]
    obj = {"type": "object_collection", "objects": plates}
    for integ in ("hierarchical", "simple"):
        out, nref, _ = gpu_vs_oracle(X, O, obj, views=((0.0, 90.0), (60.0, 80.0)), res=48, ds=0.004, integ=integ)
        assert_parity(out, nref)
        assert out["fp32"][1]["marched_tiles"] == 0 and out["fp32"][1]["span_renderer"]
        out, nref, _ = gpu_vs_oracle(X, O, obj, {"type": "rigid", "displacements": [0.0, 0.01, 0.0]}, views=((0.0, 90.0), (60.0, 80.0)),
                                     res=48, ds=0.004, integ=integ)
        assert_parity(out, nref)
        assert out["fp32"][1]["marched_tiles"] > 0


def test_every_period_boundary_direction(X, O):
    """Tessellation whose cell faces pass through the origin in x, y and z, seen along each axis and along diagonals: rays
    inside face planes (period decided by 1e-16 quantities), rays through cell edges and corners (several faces crossed
    at once), negative periods."""
    uc = {"objects": {"objects": [{"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 0.12, "rho": 0.9},
                                  {"type": "cylinder", "p0": [0.0, 0.15, 0.15], "p1": [0.3, 0.15, 0.15], "radius": 0.04, "rho": 0.5},
                                  {"type": "box", "center": [0.15, 0.0, 0.3], "sides": [0.1, 0.1, 0.1], "rho": 0.7}]},
          "xmin": 0.0, "xmax": 0.3, "ymin": 0.0, "ymax": 0.3, "zmin": 0.0, "zmax": 0.3}
    obj = {"type": "tessellated_obj_coll", "uc": uc, "xmin": -0.75, "xmax": 0.75, "ymin": -0.6, "ymax": 0.6, "zmin": -0.45, "zmax": 0.9}
    views = ((0.0, 90.0), (90.0, 90.0), (45.0, 90.0), (180.0, 90.0), (30.0, 60.0), (270.0, 120.0))
    _check(X, O, obj, views=views, res=33, ds=0.011, max_marched=40)


def test_fine_lattice_and_window_edges(X, O):
    """A very small step (8700 coarse samples, like BASELINE config 3's ds) and objects that stick out of the sample window
    [R - 1.74, R + 1.74] (clipping at smin / smax, main.go:162-169)."""
    obj = {"type": "object_collection", "objects": [{"type": "cylinder", "p0": [-2.5, 0.0, 0.1], "p1": [2.5, 0.0, 0.1], "radius": 0.15, "rho": 0.2},
                                                    {"type": "sphere", "center": [0.0, 1.8, 0.0], "radius": 0.4, "rho": 0.5},
                                                    {"type": "box", "center": [0.0, 0.0, -0.3], "sides": [4.0, 0.2, 0.1], "rho": -0.1}]}
    _check(X, O, obj, views=((0.0, 90.0), (90.0, 85.0), (33.0, 90.0)), res=24, ds=0.0004, max_marched=12)


def test_rays_inside_bounding_planes_are_settled_not_marched(X, O, scenes):
    """Axis-aligned views, even image sizes (pixel res / 2 sits at 0 exactly, main.go:463): the central row and column run INSIDE
    the planes z = 0 / y = 0, which hold a face of each box and a cap of each cylinder here -- like the cap planes of BASELINE
    config 1's hole at azimuth 0 / 90 / ....  The settle pass decides such a constraint sample by sample with the reference's
    expressions (span_settle: one monotone step per plane, bisected) instead of handing the tiles to the marching kernels.
    Only the central pixel, where the planes of all four primitives meet, may still go there."""
    obj = {"type": "object_collection", "objects": [
        {"type": "box", "center": [0.0, 0.0, 0.25], "sides": [0.9, 0.7, 0.5], "rho": 0.6},             # face z = 0: the central row at polar 90
        {"type": "cylinder", "p0": [0.1, 0.0, -0.4], "p1": [0.1, 0.0, 0.0], "radius": 0.25, "rho": -0.3},  # cap in z = 0, axis along z
        {"type": "cylinder", "p0": [0.0, 0.0, 0.1], "p1": [0.0, 0.6, 0.1], "radius": 0.07, "rho": 0.2},   # cap in y = 0: central column at az 0 / 180
        {"type": "box", "center": [-0.3, 0.2, -0.2], "sides": [0.2, 0.4, 0.1], "rho": 0.3}]}            # face y = 0
    views = ((0.0, 90.0), (90.0, 90.0), (180.0, 90.0), (270.0, 90.0), (0.0, 0.0001), (45.0, 90.0))
    _check(X, O, obj, views=views, res=33, ds=0.0137, max_marched=0)
    _check(X, O, obj, views=views, res=64, ds=0.0137, max_marched=4)
    _check(X, O, str(scenes / "cube_w_hole.json"), views=((0.0, 90.0), (90.0, 90.0), (180.0, 90.0), (270.0, 90.0)), res=65, max_marched=0)
    _check(X, O, str(scenes / "cube_w_hole.json"), views=((0.0, 90.0), (90.0, 90.0)), res=128, max_marched=1)


def test_rays_along_edges_and_non_monotone_planes_are_handed_over(X, O):
    """Two undecided planes at once (the central ray runs along an edge of a box: the members can sit in the middle of the
    range) and a plane constraint whose rounded expression need not be monotone along the ray (a cap of a diagonal cylinder,
    a face of a sheared parallelepiped, both containing the eye): neither is bisected; the image must match all the same."""
    s2 = 0.5 ** 0.5
    obj = {"type": "object_collection", "objects": [
        {"type": "box", "center": [0.0, 0.2, 0.15], "sides": [0.8, 0.4, 0.3], "rho": 0.5},                # edge y = 0, z = 0 along x
        {"type": "cylinder", "p0": [0.0, 0.0, 0.0], "p1": [0.3 * s2, 0.0, 0.3 * s2], "radius": 0.2, "rho": 0.3},  # cap plane x + z = 0
        {"type": "parallelepiped", "origin": [0.0, 0.0, 0.0], "v0": [0.3, 0.0, -0.3], "v1": [0.0, -0.4, 0.0], "v2": [0.2, 0.0, 0.25],
         "rho": 0.4}]}
    views = ((0.0, 90.0), (180.0, 90.0), (90.0, 45.0), (270.0, 135.0), (90.0, 135.0))
    for res in (33, 48):
        _check(X, O, obj, views=views, res=res, ds=0.0093)
    _, st = X.render_scene(X.Scene(obj), X.cameras_from_angles([(0.0, 90.0)], R, FOV), 48, ds=0.0093, return_stats=True)
    assert st["marched_tiles"] >= 1 and st["march_reasons"] & 2


def test_rays_along_cell_edges_of_a_tessellation(X, O, scenes):
    """Even image sizes put the central ray on a cell EDGE of the lattice at axis-aligned views (two coordinates inside face
    planes at once); the settle pass bisects each axis' period step on its own."""
    uc = {"objects": {"objects": [{"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 0.12, "rho": 0.9},
                                  {"type": "cylinder", "p0": [0.0, 0.0, 0.0], "p1": [0.3, 0.0, 0.0], "radius": 0.04, "rho": 0.5},
                                  {"type": "cylinder", "p0": [0.0, 0.0, 0.0], "p1": [0.0, 0.3, 0.0], "radius": 0.05, "rho": 0.4},
                                  {"type": "cylinder", "p0": [0.0, 0.0, 0.0], "p1": [0.0, 0.0, 0.3], "radius": 0.03, "rho": 0.3},
                                  {"type": "box", "center": [0.15, 0.0, 0.3], "sides": [0.1, 0.1, 0.1], "rho": 0.7}]},
          "xmin": 0.0, "xmax": 0.3, "ymin": 0.0, "ymax": 0.3, "zmin": 0.0, "zmax": 0.3}
    obj = {"type": "tessellated_obj_coll", "uc": uc, "xmin": -0.75, "xmax": 0.75, "ymin": -0.6, "ymax": 0.6, "zmin": -0.45, "zmax": 0.9}
    views = ((0.0, 90.0), (90.0, 90.0), (180.0, 90.0), (270.0, 90.0), (0.0, 0.0001), (0.0, 179.9999))
    _check(X, O, obj, views=views, res=32, ds=0.011, max_marched=16)
    _check(X, O, obj, views=views, res=33, ds=0.011, max_marched=0)
    _check(X, O, str(scenes / "lattice.json"), views=((0.0, 90.0), (90.0, 90.0)), res=64, max_marched=2)


@pytest.mark.parametrize("bins", [True, False])
@pytest.mark.parametrize("tilt", [1e-4, 1e-6, 3e-8, 1e-10])
def test_rays_alongside_cell_faces(X, O, scenes, monkeypatch, bins, tilt):
    """Views a hair (1e-4 ... 1e-10 degrees) off the axes: whole pixel rows run ALONGSIDE a cell face of the tessellation --
    within 1e-6 ... 1e-12 of it for the whole length, on one side first and the other later -- without being inside it to
    1e-9.  The fp32 candidate walk (tiles without a usable screen-space bin; XRAY_SPAN_NO_BINS forces it everywhere) must
    still find the children on the exact ray's side of the face, and rays nearly parallel to a strut's axis must survive the
    fp32 pre-filter."""
    if not bins:
        monkeypatch.setenv("XRAY_SPAN_NO_BINS", "1")
    uc = {"objects": {"objects": [{"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 0.12, "rho": 0.9},
                                  {"type": "cylinder", "p0": [0.0, 0.0, 0.0], "p1": [0.3, 0.0, 0.0], "radius": 0.04, "rho": 0.5},
                                  {"type": "cylinder", "p0": [0.0, 0.0, 0.0], "p1": [0.0, 0.3, 0.0], "radius": 0.05, "rho": 0.4},
                                  {"type": "cylinder", "p0": [0.0, 0.0, 0.0], "p1": [0.0, 0.0, 0.3], "radius": 0.03, "rho": 0.3},
                                  {"type": "box", "center": [0.15, 0.0, 0.3], "sides": [0.1, 0.1, 0.1], "rho": 0.7}]},
          "xmin": 0.0, "xmax": 0.3, "ymin": 0.0, "ymax": 0.3, "zmin": 0.0, "zmax": 0.3}
    obj = {"type": "tessellated_obj_coll", "uc": uc, "xmin": -0.75, "xmax": 0.75, "ymin": -0.6, "ymax": 0.6, "zmin": -0.45, "zmax": 0.9}
    views = ((tilt, 90.0), (90.0 - tilt, 90.0), (180.0, 90.0 + tilt), (270.0 + tilt, 90.0 - tilt), (0.0, tilt), (45.0, 180.0 - tilt))
    _check(X, O, obj, views=views, res=32, ds=0.011)
    _check(X, O, str(scenes / "lattice.json"), views=views[:3], res=32)
    _check(X, O, str(scenes / "pillar_array.json"), views=(views[0], views[4]), res=32)


def test_bins_that_overflow_walk_in_the_settle_pass(X, O):
    """A fine tessellation on a coarse detector: every 8 x 16 pixel tile sees far more than 128 (period, child) instances, so
    every bin overflows.  The binned fast pass has no grid walk compiled in: it leaves those tiles to the settle pass, which
    walks the candidate grid for them (and renders rays with more intervals than the lists hold in windows only where a
    bin exists -- here they are handed over).  Exact either way."""
    uc = {"objects": {"objects": [{"type": "sphere", "center": [0.06, 0.06, 0.06], "radius": 0.035, "rho": 0.9},
                                  {"type": "cylinder", "p0": [0.0, 0.06, 0.06], "p1": [0.12, 0.06, 0.06], "radius": 0.015, "rho": 0.5},
                                  {"type": "cylinder", "p0": [0.06, 0.0, 0.03], "p1": [0.06, 0.12, 0.03], "radius": 0.012, "rho": 0.4}]},
          "xmin": 0.0, "xmax": 0.12, "ymin": 0.0, "ymax": 0.12, "zmin": 0.0, "zmax": 0.12}
    obj = {"type": "tessellated_obj_coll", "uc": uc, "xmin": -0.72, "xmax": 0.72, "ymin": -0.6, "ymax": 0.6, "zmin": -0.48, "zmax": 0.6}
    views = ((17.0, 80.0), (90.0, 90.0), (45.0, 60.0))
    for integ in ("hierarchical", "simple"):
        out, nref, _ = gpu_vs_oracle(X, O, obj, views=views, res=32, ds=0.004, integ=integ)
        assert_parity(out, nref)
        st = out["fp32"][1]
        assert st["span_renderer"] and st["launches"] >= 4
        assert st["marched_tiles"] < 0.5 * len(views) * 4 * 4 * 2, st  # most warp tiles are rendered by the settle pass, not marched


def test_span_matches_marching_kernels_at_benchmark_resolution(X, scenes, monkeypatch):
    """Full 1024^2 lattice view and a 1024^2 crop-equivalent of the pillar array: interval renderer versus the marching
    kernels, which the 1e-4 gate of the rest of the suite already ties to the oracle; same reference-equivalent sample count."""
    for name, res in (("lattice", 1024), ("pillar_array", 1024)):
        sc = X.Scene(str(scenes / f"{name}.json"))
        cams = X.cameras_from_angles([(90.0, 90.0), (123.0, 90.0)], R, FOV)
        a, sa = X.render_scene(sc, cams, res, return_stats=True)
        monkeypatch.setenv("XRAY_NO_SPAN", "1")
        b, sb = X.render_scene(sc, cams, res, return_stats=True)
        monkeypatch.delenv("XRAY_NO_SPAN")
        assert np.abs(a.astype(np.float64) - b).max() <= 2e-6
        assert sa["ref_samples"] == sb["ref_samples"]
        assert sa["span_renderer"] and not sb["span_renderer"] and sa["launches"] > sb["launches"]
        assert sa["marched_tiles"] <= 0.002 * 2 * (res // 4) * (res // 8), sa  # warp tiles of 4 x 8 pixels


def test_clipping_flags_match_the_reference_probes(X, O, monkeypatch):
    """main.go:162-169: integrate_hierarchical evaluates density() at smin and smax of every ray and warns when it is > 0.
    The library reports the same two facts in the stats (bits of stats[7]), for the interval renderer and the marching kernels."""
    rod = {"type": "cylinder", "p0": [-2.5, 0.0, 0.1], "p1": [2.5, 0.0, 0.1], "radius": 0.15, "rho": 0.2}       # sticks out both ends
    half = {"type": "cylinder", "p0": [0.0, 0.0, 0.1], "p1": [2.5, 0.0, 0.1], "radius": 0.15, "rho": 0.2}       # only on the +x side
    ball = {"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 0.5, "rho": 1.0}
    R_, res, ds = 4.0, 16, 0.02
    for obj, view, want in ((rod, (0.0, 90.0), (True, True)), (half, (0.0, 90.0), (True, False)), (half, (180.0, 90.0), (False, True)),
                            (ball, (0.0, 90.0), (False, False)), (rod, (90.0, 90.0), (False, False))):
        sc = X.Scene({"type": "object_collection", "objects": [obj]})
        cams = X.cameras_from_angles([view], R_, FOV)
        for nospan in (False, True):
            if nospan:
                monkeypatch.setenv("XRAY_NO_SPAN", "1")
            else:
                monkeypatch.delenv("XRAY_NO_SPAN", raising=False)
            _, st = X.render_scene(sc, cams, res, ds=ds, return_stats=True)
            assert (st["clipping_smin"], st["clipping_smax"]) == want, (obj, view, st)
            assert st["span_renderer"] == (not nospan)
        monkeypatch.delenv("XRAY_NO_SPAN", raising=False)
        # the oracle's own density() at the two ends of every ray
        osc = O.OracleScene({"type": "object_collection", "objects": [obj]})
        eye, cm = O.camera_from_angles(view[0], view[1], R_)
        f = 1.0 / np.tan(np.deg2rad(FOV) / 2.0)
        got = [False, False]
        for i in range(res):
            for j in range(res):
                v = cm @ np.array([i / (res / 2) - 1, j / (res / 2) - 1, -f, 1.0])
                d = v[:3] / v[3] - eye
                d = d * (1.0 / np.sqrt(d @ d))
                for e, s in enumerate((R_ - 1.74, R_ + 1.74)):
                    p = eye + d * s
                    got[e] = got[e] or osc.density(*p) > 0
        assert tuple(got) == want


# ---- randomised scenes aimed at the interval renderer's special cases ---------------------------------------------------
def _snap(rng, v, step=0.05):
    """Coordinates on a coarse grid: faces, caps and centres then coincide with planes through the origin, with cell faces
    and with each other -- the situations where a ray runs inside a bounding plane or meets a surface exactly on a lattice sample."""
    return float(np.round(v / step) * step) if rng.random() < 0.6 else float(v)


def _convex_prim(rng, lo, hi, snap=True):
    ext = hi - lo
    c = lo + rng.uniform(0.15, 0.85, 3) * ext
    s = float(ext.min())
    f = (lambda v: _snap(rng, v)) if snap else float
    t = rng.choice(["sphere", "box", "cube", "cylinder", "parallelepiped"])
    rho = float(rng.choice([1.0, 0.7, 0.35, -0.5, -1.0, 0.15]))
    if t == "sphere":
        return {"type": t, "center": [f(v) for v in c], "radius": float(rng.uniform(0.1, 0.45) * s), "rho": rho}
    if t == "box":
        return {"type": t, "center": [f(v) for v in c], "sides": [f(v) + 0.05 for v in rng.uniform(0.1, 0.7, 3) * ext], "rho": rho}
    if t == "cube":
        return {"type": t, "center": [f(v) for v in c], "side": f(rng.uniform(0.15, 0.6) * s) + 0.05, "rho": rho}
    if t == "cylinder":
        axis = rng.integers(0, 4)
        d = rng.uniform(-0.45, 0.45, 3) * ext
        if axis < 3:  # axis-parallel struts: caps parallel to cell faces and image rows / columns
            d = np.zeros(3)
            d[axis] = rng.uniform(0.2, 0.6) * ext[axis] * rng.choice([-1, 1])
        p0 = np.array([f(v) for v in c])
        return {"type": t, "p0": list(p0), "p1": [f(v) for v in p0 + d + 1e-3], "radius": float(rng.uniform(0.04, 0.2) * s), "rho": rho}
    m = np.eye(3) * rng.uniform(0.2, 0.6, 3) * ext + rng.uniform(-0.1, 0.1, (3, 3)) * s
    return {"type": t, "origin": [f(v) for v in c - 0.3 * ext], "v0": list(m[0]), "v1": list(m[1]), "v2": list(m[2]), "rho": rho}


def _span_scene(rng):
    kind = rng.choice(["flat", "tess", "tess", "bare"])
    if kind == "bare":
        return _convex_prim(rng, np.array([-0.6] * 3), np.array([0.6] * 3))
    if kind == "flat":
        n = int(rng.integers(2, 12))
        return {"type": "object_collection", "greedy_dens_eval": bool(rng.random() < 0.4),
                "objects": [_convex_prim(rng, np.array([-0.7] * 3), np.array([0.7] * 3)) for _ in range(n)]}
    cell = np.array([rng.choice([0.25, 0.3, 0.4, 0.5, 0.8]) for _ in range(3)])
    lo = np.array([rng.choice([0.0, -0.1, -cell[a] / 2, 0.05]) for a in range(3)])
    objs = []
    for _ in range(int(rng.integers(1, 7))):
        p = _convex_prim(rng, lo, lo + cell)
        if rng.random() < 0.8:
            p["rho"] = abs(p["rho"])
        objs.append(p)
    uc = {"objects": {"objects": objs}, "xmin": lo[0], "xmax": lo[0] + cell[0], "ymin": lo[1], "ymax": lo[1] + cell[1],
          "zmin": lo[2], "zmax": lo[2] + cell[2]}
    b = [rng.choice([0.6, 0.75, 0.8, 0.9]) for _ in range(3)]
    return {"type": "tessellated_obj_coll", "uc": uc, "xmin": -b[0], "xmax": b[0], "ymin": -b[1], "ymax": b[1], "zmin": -b[2], "zmax": b[2]}


def _span_warp(rng):
    t = rng.choice(["none", "none", "none", "rigid", "linear", "affine"])
    if t == "none":
        return None
    if t == "rigid":
        return {"type": t, "displacements": [_snap(rng, v) for v in rng.uniform(-0.2, 0.2, 3)]}
    if t == "linear":
        return {"type": t, "strains": list(rng.uniform(-0.1, 0.1, 6))}
    return {"type": t, "matrix": (np.eye(3) + rng.uniform(-0.15, 0.15, (3, 3))).tolist()}


@pytest.mark.parametrize("seed", range(300))
def test_random_convex_scene_special_views(X, O, seed):
    """Axis-aligned and diagonal views at polar = 90 (rays inside cell-face / cap / box-face planes, surfaces met exactly on
    lattice samples, rays through cell edges), snapped geometry, both integrators, odd detector sizes, both precisions."""
    rng = np.random.default_rng(9000 + seed)
    obj = _span_scene(rng)
    deform = _span_warp(rng)
    integ = "hierarchical" if rng.random() < 0.7 else "simple"
    res = int(rng.choice([17, 32, 33, 48]))
    ds = float(rng.choice([0.03, 0.02, 0.0125, 0.01, 0.005]))
    az = float(rng.choice([0.0, 90.0, 180.0, 270.0, 45.0, 135.0, 30.0, rng.uniform(0, 360)]))
    pol = float(rng.choice([90.0, 90.0, 90.0, 60.0, rng.uniform(40, 140)]))
    out, nref, _ = gpu_vs_oracle(X, O, obj, deform, views=((az, pol),), res=res, integ=integ, ds=ds, ff=float(rng.choice([0.0, 0.1])),
                                 dm=float(rng.choice([1.0, 0.5, 2.0])))
    assert_parity(out, nref)  # (a scene without any positive density is not eligible and takes the marching kernels: fine)


@pytest.mark.parametrize("seed", range(200))
def test_random_convex_scene_near_special_views(X, O, seed, monkeypatch):
    """The same scene generator seen a hair off the special directions (tilts of 1e-3 ... 1e-11 degrees, also straight down the
    z axis), even detector sizes (central row and column through the origin), half of the cases with the fp32 candidate walk
    forced (no screen-space bins): rays alongside faces, nearly parallel to axes of cylinders, nearly inside cap planes."""
    rng = np.random.default_rng(77000 + seed)
    if rng.random() < 0.5:
        monkeypatch.setenv("XRAY_SPAN_NO_BINS", "1")
    obj = _span_scene(rng)
    deform = _span_warp(rng) if rng.random() < 0.3 else None
    integ = "hierarchical" if rng.random() < 0.6 else "simple"
    res = int(rng.choice([16, 32, 48]))
    ds = float(rng.choice([0.03, 0.02, 0.0125, 0.01, 0.005]))
    tilt = float(10.0 ** rng.uniform(-11, -3)) * (1 if rng.random() < 0.5 else -1)
    az = float(rng.choice([0.0, 90.0, 180.0, 270.0, 45.0, 135.0])) + (tilt if rng.random() < 0.7 else 0.0)
    pol = float(rng.choice([90.0, 90.0, 0.0, 180.0, 45.0])) + (tilt if rng.random() < 0.7 else 0.0)
    pol = min(max(pol, 1e-12), 180.0 - 1e-12)
    out, nref, _ = gpu_vs_oracle(X, O, obj, deform, views=((az, pol),), res=res, integ=integ, ds=ds, ff=float(rng.choice([0.0, 0.1])),
                                 dm=float(rng.choice([1.0, 0.5, 2.0])))
    assert_parity(out, nref)
