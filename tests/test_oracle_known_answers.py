"""Pins the CPU oracle against every known-answer case of the reference's own unit tests
(main_test.go, objects/objects_test.go, deformations/deformations_test.go).  The reference has no
golden images; these exact cases are all it pins."""
import math

import numpy as np
import pytest


def sphere(rho=1.0, r=0.5, c=(0.0, 0.0, 0.0)):
    return {"type": "sphere", "center": list(c), "radius": r, "rho": rho}


# ---- main_test.go: integrator physics ---------------------------------------------------
def test_simple_sphere_center_ray(O):  # TestIntegrateSimple_SphereCenterRay, main_test.go:326-339
    s = O.OracleScene(sphere())
    got, _ = s.integrate("simple", (0, 0, -5), (0, 0, 1), 0.001, 4.5, 5.5)
    assert abs(got - math.exp(-1.0)) <= 2 * 0.001


def test_simple_slab(O):  # TestIntegrateSimple_SlabAttenuation, main_test.go:343-361
    s = O.OracleScene({"type": "box", "center": [0, 0, 0], "sides": [10, 10, 1], "rho": 2.0})
    got, _ = s.integrate("simple", (0, 0, -5), (0, 0, 1), 0.001, 0, 10)
    assert abs(got - math.exp(-2.0)) <= 2 * 0.001 * 2.0


def test_hierarchical_sphere_center_ray(O):  # main_test.go:365-378
    s = O.OracleScene(sphere())
    got, _ = s.integrate("hierarchical", (0, 0, -5), (0, 0, 1), 0.05, 4.5, 5.5)
    assert abs(got - math.exp(-1.0)) <= 0.05


def test_simple_vs_hierarchical_agreement(O):  # main_test.go:382-394
    s = O.OracleScene(sphere())
    a, _ = s.integrate("simple", (0, 0, -5), (0, 0, 1), 0.001, 4.5, 5.5)
    b, _ = s.integrate("hierarchical", (0, 0, -5), (0, 0, 1), 0.05, 4.5, 5.5)
    assert abs(a - b) <= 0.01


@pytest.mark.parametrize("integ", ["simple", "hierarchical"])
def test_flat_field_applied_once(O, integ):  # main_test.go:399-420
    s = O.OracleScene(sphere(c=(0, 0, 10), r=0.01), flat_field=1.0)
    got, _ = s.integrate(integ, (0, 0, -5), (0, 0, 1), 0.01, 0, 1)
    assert abs(got - math.exp(-1.0)) <= 1e-9


def test_density_multiplier(O):  # main_test.go:424-437
    s = O.OracleScene(sphere(), density_multiplier=2.0)
    got, _ = s.integrate("simple", (0, 0, -5), (0, 0, 1), 0.001, 4.5, 5.5)
    assert abs(got - math.exp(-2.0)) <= 2 * 0.001 * 2.0


@pytest.mark.parametrize("ds,tol", [(0.01, 5e-2), (0.001, 5e-3), (0.0001, 5e-4)])
def test_simple_ds_convergence(O, ds, tol):  # main_test.go:443-465
    s = O.OracleScene(sphere())
    got, _ = s.integrate("simple", (0, 0, -5), (0, 0, 1), ds, 4.5, 5.5)
    assert abs(got - math.exp(-1.0)) <= tol


@pytest.mark.parametrize("DS", [0.1, 0.05, 0.02])
def test_hierarchical_boundary_accuracy(O, DS):  # main_test.go:476-497
    s = O.OracleScene(sphere())
    ref, _ = s.integrate("simple", (0, 0, -5), (0, 0, 1), 0.0001, 4.5, 5.5)
    got, _ = s.integrate("hierarchical", (0, 0, -5), (0, 0, 1), DS, 4.5, 5.5)
    assert abs(got - ref) <= DS


def test_hierarchical_off_center(O):  # main_test.go:501-522
    s = O.OracleScene(sphere())
    d = 0.3
    want = math.exp(-2 * math.sqrt(0.25 - d * d))
    ref, _ = s.integrate("simple", (d, 0, -5), (0, 0, 1), 0.0001, 4.5, 5.5)
    got, _ = s.integrate("hierarchical", (d, 0, -5), (0, 0, 1), 0.05, 4.5, 5.5)
    assert abs(ref - want) <= 1e-3
    assert abs(got - ref) <= 0.05


# fp64 values of a plain-Python restatement of main.go:144-199
APPENDIX_B = [
    ("simple", 1e-3, (0, 0, -5), 1.0, 0.3682475046136626),
    ("simple", 1e-4, (0, 0, -5), 1.0, 0.3678794411714768),
    ("hierarchical", 0.1, (0, 0, -5), 1.0, 0.3642189795715233),
    ("hierarchical", 0.05, (0, 0, -5), 1.0, 0.3660446348040152),
    ("hierarchical", 0.02, (0, 0, -5), 1.0, 0.36714441755772087),
    ("hierarchical", 0.05, (0.3, 0, -5), 1.0, 0.4470879265593563),
    ("simple", 1e-3, (0, 0, -5), 2.0, 0.13560622465418948),
]


@pytest.mark.parametrize("integ,ds,origin,dm,want", APPENDIX_B)
def test_appendix_b_values(O, integ, ds, origin, dm, want):
    s = O.OracleScene(sphere(), density_multiplier=dm)
    got, _ = s.integrate(integ, origin, (0, 0, 1), ds, 4.5, 5.5)
    assert abs(got - want) <= 2e-16


def test_lattice_step_counts(O):  # SURVEY.md Appendix B lattice facts (fp64 repeated addition)
    lo, hi = 4 - 1.74, 4 + 1.74
    assert [O.step_count("hierarchical", ds, lo, hi) for ds in (0.03, 0.02, 0.005, 0.0004)] == [115, 174, 696, 8700]
    assert O.step_count("simple", 0.03, lo, hi) == 116
    assert O.step_count("simple", 2 / 1024 / 5, lo, hi) == 8909
    assert O.step_count("simple", 2 / 32 / 5, lo, hi) == 279


# ---- main_test.go: camera / angles -----------------------------------------------------
@pytest.mark.parametrize("az,pol,R", [(0, 90, 4), (90, 90, 4), (45, 60, 2.5), (200, 120, 5), (90, 10, 4)])
def test_camera_eye_and_invariants(O, az, pol, R):  # TestComputeCameraFromAngles pins eye to 1e-6
    eye, cam = O.camera_from_angles(az, pol, R)
    th, ph = math.radians(az), math.radians(pol)
    want = np.array([R * math.cos(th) * math.sin(ph), R * math.sin(th) * math.sin(ph), R * math.cos(ph)])
    assert np.abs(eye - want).max() <= 1e-6
    # invariants that stand in for the un-vendored mathgl
    assert np.abs(cam[:3, 3] - eye).max() <= 1e-12            # camera * (0,0,0,1) = eye
    rot = cam[:3, :3]
    assert np.abs(rot.T @ rot - np.eye(3)).max() <= 1e-12      # orthonormal
    f = -eye / R
    assert np.abs(rot[:, 2] - (-f)).max() <= 1e-12             # third column = -forward
    assert np.allclose(cam[3], [0, 0, 0, 1], atol=1e-15)
    s = np.array([-math.sin(th), math.cos(th), 0.0])            # closed form for sin(phi) > 0
    assert np.abs(rot[:, 0] - s).max() <= 1e-12


def test_generate_camera_angles(O):  # main_test.go:197-283: az_i = 90 + i*360/N, modulo sharding
    a = O.generate_camera_angles(8)
    assert [x[0] for x in a] == [90.0 + i * 45.0 for i in range(8)]
    assert all(x[1] == 90.0 for x in a)
    b = O.generate_camera_angles(8, job_num=1, jobs_modulo=3)
    assert [x[0] for x in b] == [90.0 + i * 45.0 for i in (1, 4, 7)]


# ---- objects_test.go --------------------------------------------------------------------
def test_sphere_geometry(O):  # objects_test.go:14-42
    s = O.OracleScene(sphere(rho=2.0))
    assert [s.object_density(0, 0, z) for z in (0, 0.4, 0.5, 0.6)] == [2.0, 2.0, 0.0, 0.0]
    s2 = O.OracleScene(sphere(rho=5.0, r=0.3, c=(1, 2, 3)))
    assert s2.object_density(1, 2, 3) == 5.0 and s2.object_density(1, 2, 3.31) == 0.0


def test_box_geometry(O):  # objects_test.go:44-67
    b = O.OracleScene({"type": "box", "center": [0, 0, 0], "sides": [2, 1, 0.5], "rho": 1.0})
    cases = [((0, 0, 0), 1), ((0.9, 0, 0), 1), ((1.0, 0, 0), 0), ((0, 0.4, 0), 1), ((0, 0.5, 0), 0), ((0, 0, 0.24), 1),
             ((0, 0, 0.25), 0)]
    for p, want in cases:
        assert b.object_density(*p) == want


def test_cylinder_geometry(O):  # objects_test.go:69-93
    c = O.OracleScene({"type": "cylinder", "p0": [0, 0, -1], "p1": [0, 0, 1], "radius": 0.5, "rho": 1.0})
    cases = [((0, 0, 0), 1), ((0.4, 0, 0), 1), ((0.5, 0, 0), 0), ((0, 0, -1.0), 1), ((0, 0, 1.0), 1), ((0, 0, -1.001), 0),
             ((0, 0, 1.001), 0), ((0.4, 0, 1.0), 1)]
    for p, want in cases:
        assert c.object_density(*p) == want


def test_cylinder_default_rho(O):  # objects.go:326-330
    c = O.OracleScene({"type": "cylinder", "p0": [0, 0, -1], "p1": [0, 0, 1], "radius": 0.5})
    assert c.object_density(0, 0, 0) == 1.0


def test_collection_clamping_and_greedy(O):  # objects_test.go:95-112
    objs = [sphere(rho=0.8), sphere(rho=0.8)]
    assert O.OracleScene({"type": "object_collection", "objects": objs}).object_density(0, 0, 0) == 1.0
    assert O.OracleScene({"type": "object_collection", "objects": objs, "greedy_dens_eval": True}).object_density(0, 0, 0) == 0.8


def vox(arr):
    return {"type": "voxel_grid", "_array": np.asarray(arr, dtype=np.float64)}


def test_voxel_layout(O):  # objects_test.go:119-155: idx = z*NX*NY + x*NY + y
    N = 3
    rho = np.zeros(N * N * N)
    rho[0 * N * N + 2 * N + 0] = 1.0
    vg = O.OracleScene(vox(rho.reshape(N, N, N)))
    assert abs(vg.object_density(1.0, -1.0, -1.0) - 1.0) <= 1e-9
    assert vg.object_density(-1.0, 1.0, -1.0) <= 1e-9
    rho2 = np.zeros(N * N * N)
    rho2[1 * N * N] = 1.0
    vg2 = O.OracleScene(vox(rho2.reshape(N, N, N)))
    assert abs(vg2.object_density(-1.0, -1.0, 0.0) - 1.0) <= 1e-9
    assert vg2.object_density(-1.0, 0.0, -1.0) <= 1e-9


def test_voxel_outside_bounds(O):  # objects_test.go:209-238
    vg = O.OracleScene(vox(np.ones((2, 2, 2))))
    for p in [(1.001, 0, 0), (-1.001, 0, 0), (0, 1.001, 0), (0, -1.001, 0), (0, 0, 1.001), (0, 0, -1.001), (2, 2, 2)]:
        assert vg.object_density(*p) == 0.0
    assert vg.object_density(1.0, 1.0, 1.0) == 1.0 and vg.object_density(-1.0, -1.0, -1.0) == 1.0


def test_voxel_trilinear(O):  # objects_test.go:240-268
    vg = O.OracleScene(vox(np.ones((2, 2, 2))))
    for p in [(0, 0, 0), (0.5, -0.5, 0.3), (1, 1, 1), (-1, -1, -1)]:
        assert abs(vg.object_density(*p) - 1.0) <= 1e-9
    r = np.zeros(8)
    r[0] = 1.0
    vg2 = O.OracleScene(vox(r.reshape(2, 2, 2)))
    assert abs(vg2.object_density(-1, -1, -1) - 1.0) <= 1e-9
    assert abs(vg2.object_density(0.0, -1, -1) - 0.5) <= 1e-9


def test_voxel_non_cubic(O):  # objects_test.go:311-349
    nx, ny, nz = 2, 4, 3
    rho = np.zeros(nx * ny * nz)
    rho[2 * nx * ny + 1 * ny + 3] = 1.0
    vg = O.OracleScene(vox(rho.reshape(nz, nx, ny)))
    assert abs(vg.object_density(1, 1, 1) - 1.0) <= 1e-9
    assert vg.object_density(-1, 1, 1) <= 1e-9 and vg.object_density(1, -1, 1) <= 1e-9
    rho2 = np.zeros(nx * ny * nz)
    rho2[1] = 1.0
    vg2 = O.OracleScene(vox(rho2.reshape(nz, nx, ny)))
    wy = 1.0 / (ny - 1) * 2.0 - 1.0
    assert abs(vg2.object_density(-1.0, wy, -1.0) - 1.0) <= 1e-9
    assert vg2.object_density(wy, -1.0, -1.0) <= 1e-9


def test_voxel_non_cubic_axis_separation(O):  # objects_test.go:433-461
    nx, ny, nz = 16, 32, 16
    k, i, j = np.meshgrid(np.arange(nz), np.arange(nx), np.arange(ny), indexing="ij")
    x, y, z = i / (nx - 1) * 2 - 1, j / (ny - 1) * 2 - 1, k / (nz - 1) * 2 - 1
    rho = (((x - 0.5) ** 2 + y ** 2 + z ** 2) < 0.3 ** 2).astype(np.float64)
    vg = O.OracleScene(vox(rho))
    assert vg.object_density(0.5, 0, 0) >= 0.8 and vg.object_density(0, 0.5, 0) <= 0.1


def test_voxel_raw_uint8(O, tmp_path):  # objects_test.go:282-303: u8 -> b/255
    p = tmp_path / "u8.raw"
    p.write_bytes(bytes([255, 128, 0]))
    arr, dims = O.voxel_grid_from_raw(str(p), [1, 1, 3], "uint8")
    assert dims == (1, 1, 3)
    assert np.abs(arr.ravel() - np.array([1.0, 128.0 / 255.0, 0.0])).max() <= 1e-10


def kelvin(rad, scale):  # objects.go:588-637 MakeKelvin strut table
    P = [
        (.25, 0, .5, .5, 0, .75), (.25, 1, .5, .5, 1, .75), (.25, 0, .5, .5, 0, .25), (.25, 1, .5, .5, 1, .25),
        (.25, 0, .5, 0, .25, .5), (.5, 0, .75, .75, 0, .5), (.5, 1, .75, .75, 1, .5), (.5, 0, .75, .5, .25, 1),
        (.75, 0, .5, .5, 0, .25), (.75, 1, .5, .5, 1, .25), (.75, 0, .5, 1, .25, .5), (.5, 0, .25, .5, .25, 0),
        (1, .5, .75, .75, .5, 1), (1, .75, .5, .75, 1, .5), (1, .5, .25, .75, .5, 0), (.25, 1, .5, 0, .75, .5),
        (.5, 1, .75, .5, .75, 1), (.5, 1, .25, .5, .75, 0), (0, .25, .5, 0, .5, .75), (1, .25, .5, 1, .5, .75),
        (0, .25, .5, 0, .5, .25), (1, .25, .5, 1, .5, .25), (0, .5, .75, .25, .5, 1), (0, .5, .75, 0, .75, .5),
        (1, .5, .75, 1, .75, .5), (0, .75, .5, 0, .5, .25), (1, .75, .5, 1, .5, .25), (0, .5, .25, .25, .5, 0),
        (.25, .5, 0, .5, .75, 0), (.25, .5, 1, .5, .75, 1), (.25, .5, 0, .5, .25, 0), (.25, .5, 1, .5, .25, 1),
        (.5, .75, 0, .75, .5, 0), (.5, .75, 1, .75, .5, 1), (.75, .5, 0, .5, .25, 0), (.75, .5, 1, .5, .25, 1),
    ]
    objs = [{"type": "cylinder", "p0": [a * scale, b * scale, c * scale], "p1": [d * scale, e * scale, f * scale],
             "radius": rad, "rho": 1.0} for a, b, c, d, e, f in P]
    return {"objects": {"type": "object_collection", "objects": objs}, "xmin": 0.0, "xmax": scale, "ymin": 0.0,
            "ymax": scale, "zmin": 0.0, "zmax": scale}


def tess(uc, b=2.0):
    return {"type": "tessellated_obj_coll", "uc": uc, "xmin": -b, "xmax": b, "ymin": -b, "ymax": b, "zmin": -b, "zmax": b}


def test_tessellation_x_axis_cylinder(O):  # objects_test.go:356-385
    uc = {"objects": {"objects": [{"type": "cylinder", "p0": [0, 0.5, 0.5], "p1": [1, 0.5, 0.5], "radius": 0.1, "rho": 1.0}]},
          "xmin": 0, "xmax": 1, "ymin": 0, "ymax": 1, "zmin": 0, "zmax": 1}
    lat = O.OracleScene(tess(uc))
    for x in (-1.9, -1.5, -1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 1.9):
        assert lat.object_density(x, 0.5, 0.5) != 0
    assert lat.object_density(1.0 - 1e-9, 0.5, 0.5) != 0 and lat.object_density(1.0 + 1e-9, 0.5, 0.5) != 0


def test_tessellation_kelvin_faces(O):  # objects_test.go:387-428
    lat = O.OracleScene(tess(kelvin(0.1, 1.0)))
    a, b = lat.object_density(0.5, 0.1875, 0.9375), lat.object_density(0.5, 0.1875, 1.0625)
    assert a != 0 and b != 0 and abs(a - b) <= 1e-9
    assert lat.object_density(0.375, 0.01, 0.625) != 0 and lat.object_density(0.375, -0.01, 0.625) != 0


# ---- deformations_test.go ---------------------------------------------------------------
def test_affine_is_linear_map(O):  # deformations_test.go:11-32
    M = [[1.1, 0.2, 0.0], [0.0, 0.9, 0.1], [0.3, 0.0, 1.2]]
    s = O.OracleScene(sphere(), {"type": "affine", "matrix": M})
    p = (0.3, -0.4, 0.5)
    got = s.deform(*p)
    want = tuple(sum(M[r][c] * p[c] for c in range(3)) for r in range(3))
    assert np.abs(np.array(got) - np.array(want)).max() <= 1e-12
    ident = O.OracleScene(sphere(), {"type": "affine", "matrix": [[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]})
    assert ident.deform(*p) == p  # identity leaves the point untouched (no "+ p" displacement term)


def test_gaussian_is_displacement(O):  # deformations_test.go:34-56
    d = {"type": "gaussian", "amplitudes": [0.1, 0.2, 0.3], "sigmas": [0.5, 0.5, 0.5], "centers": [0.0, 0.0, 0.0]}
    s = O.OracleScene(sphere(), d)
    assert np.abs(np.array(s.deform(0, 0, 0)) - np.array([0.1, 0.2, 0.3])).max() <= 1e-12
    far = s.deform(10, 10, 10)
    assert np.abs(np.array(far) - 10).max() <= 1e-9


# ---- auto step (main.go:350-353 + MinFeatureSize) ---------------------------------------
def test_auto_ds_of_bundled_scenes(O, scenes):
    want = {"cube_w_hole": 0.03, "balls": 0.03, "box_w_pped": 0.0207846, "pillar_array": 0.02, "lattice": 0.005,
            "gyroid_example": 0.0004}
    for name, ds in want.items():
        got = O.OracleScene(str(scenes / f"{name}.json")).auto_ds()
        assert abs(got - ds) <= 1e-7, (name, got)


def test_borrowed_fp32_volume_equals_copied_fp64_volume(O):
    """The `_array_f32` hook (full-size 1024^3 tests) must give bit-identical densities to the fp64 copy path."""
    rng = np.random.default_rng(3)
    vol = rng.random((9, 7, 11), dtype=np.float32)
    a = O.OracleScene({"type": "voxel_grid", "_array": vol.astype(np.float64)})
    b = O.OracleScene({"type": "voxel_grid", "_array_f32": vol})
    for x, y, z in rng.uniform(-1.05, 1.05, size=(400, 3)):
        assert a.density(x, y, z) == b.density(x, y, z)
    assert a.min_feature_size() == b.min_feature_size()
