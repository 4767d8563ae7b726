"""Cross-checks the C++ oracle against an independent plain-Python (IEEE double, no FMA) restatement
of the same reference functions: objects.go Density methods, deformations.go Apply methods and the
two integrators of main.go.  Pure-Python loops, so small cases only.  Agreement must be bit exact
except where libm transcendentals differ (gyroid / sigmoid / gaussian: 1e-12)."""
import json
import math
import random

import numpy as np
import pytest


# ---- independent restatement ------------------------------------------------------------
def py_density(o, x, y, z):
    t = o["type"]
    if t == "sphere":  # objects.go:63-72
        c = o["center"]
        x, y, z = x - c[0], y - c[1], z - c[2]
        return o["rho"] if x * x + y * y + z * z < o["radius"] * o["radius"] else 0.0
    if t in ("box", "cube"):  # objects.go:171-179, :115-121
        c = o["center"]
        s = o["sides"] if t == "box" else [o["side"]] * 3
        return o["rho"] if (abs(x - c[0]) < 0.5 * s[0] and abs(y - c[1]) < 0.5 * s[1] and abs(z - c[2]) < 0.5 * s[2]) else 0.0
    if t == "cylinder":  # objects.go:334-350
        p0, p1 = o["p0"], o["p1"]
        v = [p1[i] - p0[i] for i in range(3)]
        w = [x - p0[0], y - p0[1], z - p0[2]]
        c = (w[0] * v[0] + w[1] * v[1] + w[2] * v[2]) / (v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
        if c < 0.0 or c > 1.0:
            return 0.0
        e = [w[i] - v[i] * c for i in range(3)]
        d = math.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
        return o.get("rho", 1.0) if d < o["radius"] else 0.0
    if t == "gyroid":  # objects.go:1014-1032
        c, s = o["center"], o["scale"]
        x, y, z = (x - c[0]) / s, (y - c[1]) / s, (z - c[2]) / s
        g = math.sin(x) * math.cos(y) + math.sin(y) * math.cos(z) + math.sin(z) * math.cos(x)
        return o["rho"] if abs(g) < o["thickness"] else 0.0
    if t == "object_collection":  # objects.go:422-438
        dens = 0.0
        for c in o["objects"]:
            rho = py_density(c, x, y, z)
            if o.get("greedy_dens_eval", False) and rho > 0.0:
                return rho
            dens += rho
        return min(max(dens, 0.0), 1.0) if not (dens < 0.0) else 0.0
    if t == "tessellated_obj_coll":  # objects.go:568-582, :458-464
        if x < o["xmin"] or x > o["xmax"] or y < o["ymin"] or y > o["ymax"] or z < o["zmin"] or z > o["zmax"]:
            return 0.0
        uc = o["uc"]
        dx = uc["xmax"] - uc["xmin"]
        x = x - dx * math.floor((x - uc["xmin"]) / dx)
        dy = uc["ymax"] - uc["ymin"]
        y = y - dy * math.floor((y - uc["ymin"]) / dy)
        dz = uc["zmax"] - uc["zmin"]
        z = z - dz * math.floor((z - uc["zmin"]) / dz)
        if x < uc["xmin"] or x > uc["xmax"] or y < uc["ymin"] or y > uc["ymax"] or z < uc["zmin"] or z > uc["zmax"]:
            return 0.0
        coll = dict(uc["objects"], type="object_collection", greedy_dens_eval=True)
        return py_density(coll, x, y, z)
    raise ValueError(t)


def py_deform(d, x, y, z):
    t = d["type"]
    if t == "linear":  # deformations.go:136-141
        s = d["strains"]
        return (x + s[0] * x + s[5] * y + s[4] * z, y + s[5] * x + s[1] * y + s[3] * z, z + s[4] * x + s[3] * y + s[2] * z)
    if t == "rigid":
        u = d["displacements"]
        return (x + u[0], y + u[1], z + u[2])
    if t == "sigmoid":  # deformations.go:210-222
        A, c, L = d["amplitude"], d["center"], d["lengthscale"]
        q = {"x": x, "y": y, "z": z}[d["direction"]]
        q = q + A / (1 + math.exp(-(q - c) / L))
        return (q, y, z) if d["direction"] == "x" else ((x, q, z) if d["direction"] == "y" else (x, y, q))
    if t == "affine":
        M = d["matrix"]
        return tuple(M[r][0] * x + M[r][1] * y + M[r][2] * z for r in range(3))
    if t == "gaussian":  # deformations.go:29-38
        A, S, C = d["amplitudes"], d["sigmas"], d["centers"]
        x0, y0, z0 = x - C[0], y - C[1], z - C[2]
        r2 = x0 * x0 + y0 * y0 + z0 * z0
        return (x + A[0] * math.exp(-r2 / (2 * S[0] * S[0])), y + A[1] * math.exp(-r2 / (2 * S[1] * S[1])),
                z + A[2] * math.exp(-r2 / (2 * S[2] * S[2])))
    if t == "composed":
        for s in d["deformations"]:
            x, y, z = py_deform(s, x, y, z)
        return (x, y, z)
    raise ValueError(t)


def py_integrate(dens, integ, o, d, ds, smin, smax, ff=0.0):
    ln = 1.0 / math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    d = [d[0] * ln, d[1] * ln, d[2] * ln]
    T = ff
    if integ == "simple":  # main.go:144-154
        s = smin
        while s < smax:
            T += dens(o[0] + d[0] * s, o[1] + d[1] * s, o[2] + d[2] * s) * ds
            s += ds
        return math.exp(-T)
    right, left, dsf, prev = smin + ds, smin, ds / 10.0, 0.0  # main.go:159-199
    while right <= smax:
        rho = dens(o[0] + d[0] * right, o[1] + d[1] * right, o[2] + d[2] * right)
        if (rho == 0) != (prev == 0):
            left += dsf
            while left < right:
                T += dens(o[0] + d[0] * left, o[1] + d[1] * left, o[2] + d[2] * left) * dsf
                left += dsf
            T += rho * dsf
        else:
            T += rho * ds
        prev = rho
        left = right
        right += ds
    return math.exp(-T)


# ---- tests ------------------------------------------------------------------------------
SCENE_FILES = ["cube_w_hole", "balls", "box_w_pped", "pillar_array", "lattice", "gyroid_example"]


@pytest.mark.parametrize("name", SCENE_FILES)
def test_density_matches_python(O, scenes, name):
    obj = json.loads((scenes / f"{name}.json").read_text())
    if name == "box_w_pped":
        obj["objects"] = [c for c in obj["objects"] if c["type"] != "parallelepiped"]  # pped needs mgl64 Mat3.Inv
    osc = O.OracleScene(obj)
    rng = random.Random(5)
    tol = 0.0
    bad = 0
    for _ in range(4000):
        p = [rng.uniform(-1.05, 1.05) for _ in range(3)]
        a, b = osc.object_density(*p), py_density(obj, *p)
        if a != b:
            bad += 1
    # gyroid: sin/cos of glibc vs CPython's libm are the same library here; any mismatch is a bug
    assert bad == 0, f"{bad} mismatches (tol {tol})"


def test_parallelepiped_inverse_invariant(O):
    """Mat3.Inv comes from un-vendored mathgl: pin M * Minv = I and a hand-checkable membership case."""
    v0, v1, v2 = [0.7, 0.0, 0.0], [0.1, 0.6, 0.0], [0.1, 0.1, 0.5]
    m = np.array([v0, v1, v2], dtype=np.float64)  # rows = columns of M (column-major storage)
    out = np.zeros(9)
    import ctypes
    dp = ctypes.POINTER(ctypes.c_double)
    O.lib().oracle_mat3_inv(np.ascontiguousarray(m.ravel()).ctypes.data_as(dp), out.ctypes.data_as(dp))
    M = m.T
    Minv = out.reshape(3, 3).T
    assert np.abs(M @ Minv - np.eye(3)).max() <= 1e-14
    pp = {"type": "parallelepiped", "origin": [-0.1, -0.1, -0.1], "v0": v0, "v1": v1, "v2": v2, "rho": -1.0}
    osc = O.OracleScene(pp)
    inside = np.array([-0.1, -0.1, -0.1]) + 0.5 * np.array(v0) + 0.5 * np.array(v1) + 0.5 * np.array(v2)
    assert osc.object_density(*inside) == -1.0
    outside = np.array([-0.1, -0.1, -0.1]) + 1.01 * np.array(v0) + 0.5 * np.array(v1) + 0.5 * np.array(v2)
    assert osc.object_density(*outside) == 0.0
    assert osc.object_density(-0.1, -0.1, -0.1) == 0.0  # q = 0 is outside (strict)


DEFORMS = [
    {"type": "linear", "strains": [0.0, 0.0, 0.0, 0.1, 0.1, 0.1]},
    {"type": "linear", "strains": [0.05, -0.02, 0.01, 0.1, -0.03, 0.07]},
    {"type": "rigid", "displacements": [0.1, -0.2, 0.05]},
    {"type": "sigmoid", "amplitude": 0.2, "center": 0.0, "lengthscale": 0.2, "direction": "z"},
    {"type": "sigmoid", "amplitude": -0.1, "center": 0.3, "lengthscale": 0.05, "direction": "x"},
    {"type": "affine", "matrix": [[1.1, 0.2, 0.0], [0.0, 0.9, 0.1], [0.3, 0.0, 1.2]]},
    {"type": "gaussian", "amplitudes": [0.1, 0.0, -0.1], "sigmas": [0.3, 0.4, 0.5], "centers": [0.1, 0.0, -0.2]},
    {"type": "composed", "deformations": [{"type": "rigid", "displacements": [0.1, 0.0, 0.0]},
                                          {"type": "sigmoid", "amplitude": 0.2, "center": 0.0, "lengthscale": 0.2, "direction": "y"},
                                          {"type": "linear", "strains": [0.01, 0.02, 0.03, 0.0, 0.0, 0.05]}]},
]


@pytest.mark.parametrize("d", DEFORMS, ids=lambda d: d["type"])
def test_deformation_matches_python(O, d):
    osc = O.OracleScene({"type": "sphere", "center": [0, 0, 0], "radius": 0.5, "rho": 1.0}, d)
    rng = random.Random(11)
    for _ in range(500):
        p = [rng.uniform(-1.2, 1.2) for _ in range(3)]
        got, want = osc.deform(*p), py_deform(d, *p)
        assert got == want


@pytest.mark.parametrize("integ", ["simple", "hierarchical"])
@pytest.mark.parametrize("name,ds", [("cube_w_hole", 0.03), ("pillar_array", 0.02), ("lattice", 0.02)])
def test_integrators_match_python(O, scenes, integ, name, ds):
    obj = json.loads((scenes / f"{name}.json").read_text())
    d = {"type": "sigmoid", "amplitude": 0.2, "center": 0.0, "lengthscale": 0.2, "direction": "z"}
    osc = O.OracleScene(obj, d, flat_field=0.05, density_multiplier=1.3)

    def dens(x, y, z):
        x, y, z = py_deform(d, x, y, z)
        return py_density(obj, x, y, z) * 1.3

    rng = random.Random(3)
    for _ in range(12):
        o = [4.0 * math.cos(rng.uniform(0, 6.28)), 4.0 * math.sin(rng.uniform(0, 6.28)), rng.uniform(-0.5, 0.5)]
        tgt = [rng.uniform(-0.6, 0.6) for _ in range(3)]
        dr = [tgt[i] - o[i] for i in range(3)]
        got, _ = osc.integrate(integ, o, dr, ds, 4 - 1.74, 4 + 1.74)
        want = py_integrate(dens, integ, o, dr, ds, 4 - 1.74, 4 + 1.74, ff=0.05)
        assert abs(got - want) <= 1e-15


def test_render_view_pixel_mapping(O, scenes):
    """main.go:457-465: pixel (i, j) -> ((i/(res/2)-1, j/(res/2)-1, -f)) through the camera matrix."""
    obj = json.loads((scenes / "cube_w_hole.json").read_text())
    osc = O.OracleScene(obj)
    res, Rr, fov, ds = 8, 4.0, 40.0, 0.03
    eye, cam = O.camera_from_angles(120.0, 80.0, Rr)
    img, n = osc.render_view(eye, cam, res, fov, Rr, ds, "hierarchical")
    f = 1 / math.tan((fov / 2) * math.pi / 180.0)
    for (i, j) in [(0, 0), (3, 5), (4, 4), (7, 2)]:
        v = np.array([i / (res / 2) - 1, j / (res / 2) - 1, -f, 1.0])
        w = cam @ v
        vx = w[:3] * (1 / w[3])
        dr = vx - eye
        want = py_integrate(lambda x, y, z: py_density(obj, x, y, z), "hierarchical", list(eye), list(dr), ds, Rr - 1.74, Rr + 1.74)
        assert abs(img[i, j] - want) <= 1e-12
