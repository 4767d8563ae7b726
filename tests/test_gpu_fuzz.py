"""Randomised scenes against the oracle: random primitives (all types, mixed-sign rho), greedy and summed
collections, tessellations (also nested inside collections), voxel grids as children, random warps, both
integrators, odd detector sizes.  Every case must meet the same gates as the bundled scenes."""
import numpy as np
import pytest

from helpers import assert_parity, gpu_vs_oracle

pytestmark = pytest.mark.gpu


def rnd_prim(rng, scale=1.0, allow_gyroid=True):
    kinds = ["sphere", "box", "cube", "cylinder", "parallelepiped"] + (["gyroid"] if allow_gyroid else [])
    t = rng.choice(kinds)
    c = list(rng.uniform(-0.5, 0.5, 3) * scale)
    rho = float(rng.choice([1.0, 0.7, 0.35, -0.5, -1.0, 0.15]))
    if t == "sphere":
        return {"type": t, "center": c, "radius": float(rng.uniform(0.08, 0.4) * scale), "rho": rho}
    if t == "box":
        return {"type": t, "center": c, "sides": list(rng.uniform(0.1, 0.9, 3) * scale), "rho": rho}
    if t == "cube":
        return {"type": t, "center": c, "side": float(rng.uniform(0.1, 0.8) * scale), "rho": rho}
    if t == "cylinder":
        p1 = list(np.array(c) + rng.uniform(-0.6, 0.6, 3) * scale)
        return {"type": t, "p0": c, "p1": p1, "radius": float(rng.uniform(0.03, 0.2) * scale), "rho": rho}
    if t == "parallelepiped":
        m = np.eye(3) * rng.uniform(0.2, 0.6, 3) + rng.uniform(-0.15, 0.15, (3, 3))
        return {"type": t, "origin": c, "v0": list(m[0] * scale), "v1": list(m[1] * scale), "v2": list(m[2] * scale), "rho": rho}
    return {"type": "gyroid", "center": c, "scale": float(rng.uniform(0.08, 0.25)), "thickness": float(rng.uniform(0.1, 0.6)),
            "rho": abs(rho)}


def rnd_tess(rng):
    cell = rng.uniform(0.3, 0.7, 3)
    lo = rng.uniform(-0.2, 0.1, 3)
    objs = []
    for _ in range(int(rng.integers(1, 6))):
        p = rnd_prim(rng, scale=float(cell.min()), allow_gyroid=rng.random() < 0.2)
        for key in ("center", "p0", "p1", "origin"):
            if key in p:
                p[key] = list(np.array(p[key]) + lo + 0.5 * cell)
        p["rho"] = abs(p["rho"]) if rng.random() < 0.8 else p["rho"]
        objs.append(p)
    uc = {"objects": {"objects": objs}, "xmin": lo[0], "xmax": lo[0] + cell[0], "ymin": lo[1], "ymax": lo[1] + cell[1],
          "zmin": lo[2], "zmax": lo[2] + cell[2]}
    b = rng.uniform(0.5, 0.95, 3)
    return {"type": "tessellated_obj_coll", "uc": uc, "xmin": -b[0], "xmax": b[0], "ymin": -b[1], "ymax": b[1], "zmin": -b[2], "zmax": b[2]}


def rnd_deform(rng):
    t = rng.choice(["none", "none", "sigmoid", "linear", "rigid", "affine", "gaussian", "composed"])
    if t == "none":
        return None
    if t == "sigmoid":
        return {"type": t, "amplitude": float(rng.uniform(-0.25, 0.25)), "center": float(rng.uniform(-0.3, 0.3)),
                "lengthscale": float(rng.uniform(0.05, 0.4)), "direction": str(rng.choice(["x", "y", "z"]))}
    if t == "linear":
        return {"type": t, "strains": list(rng.uniform(-0.12, 0.12, 6))}
    if t == "rigid":
        return {"type": t, "displacements": list(rng.uniform(-0.2, 0.2, 3))}
    if t == "affine":
        return {"type": t, "matrix": (np.eye(3) + rng.uniform(-0.2, 0.2, (3, 3))).tolist()}
    if t == "gaussian":
        return {"type": t, "amplitudes": list(rng.uniform(-0.15, 0.15, 3)), "sigmas": list(rng.uniform(0.2, 0.6, 3)),
                "centers": list(rng.uniform(-0.3, 0.3, 3))}
    return {"type": "composed", "deformations": [d for d in (rnd_deform(rng), rnd_deform(rng)) if d and d["type"] != "composed"] or
            [{"type": "rigid", "displacements": [0.05, 0.0, -0.05]}]}


def rnd_scene(rng):
    shape = rng.choice(["prim", "flat", "flat_big", "tess", "nested"])
    if shape == "prim":
        return rnd_prim(rng)
    if shape == "flat":
        return {"type": "object_collection", "greedy_dens_eval": bool(rng.random() < 0.3),
                "objects": [rnd_prim(rng) for _ in range(int(rng.integers(1, 9)))]}
    if shape == "flat_big":
        return {"type": "object_collection", "greedy_dens_eval": bool(rng.random() < 0.3),
                "objects": [rnd_prim(rng, scale=0.5, allow_gyroid=False) for _ in range(int(rng.integers(64, 90)))]}
    if shape == "tess":
        return rnd_tess(rng)
    kids = [rnd_prim(rng) for _ in range(int(rng.integers(1, 4)))] + [rnd_tess(rng)]
    if rng.random() < 0.5:
        kids.append({"type": "voxel_grid", "_array": rng.random((int(rng.integers(2, 9)), int(rng.integers(2, 9)), int(rng.integers(2, 9)))) * 0.5})
    if rng.random() < 0.3:
        kids.append(rnd_tess(rng))
    rng.shuffle(kids)
    return {"type": "object_collection", "greedy_dens_eval": bool(rng.random() < 0.3), "objects": kids}


@pytest.mark.parametrize("seed", range(150))
def test_random_scene(X, O, seed):
    rng = np.random.default_rng(1000 + seed)
    obj = rnd_scene(rng)
    deform = rnd_deform(rng)
    integ = "hierarchical" if rng.random() < 0.7 else "simple"
    res = int(rng.choice([17, 24, 32]))
    ds = float(rng.choice([0.03, 0.017, 0.01]))
    views = ((float(rng.uniform(0, 360)), float(rng.uniform(35, 145))),)
    out, nref, _ = gpu_vs_oracle(X, O, obj, deform, views=views, res=res, integ=integ, ds=ds, ff=float(rng.choice([0.0, 0.2])),
                                 dm=float(rng.choice([1.0, 0.5, 2.0])))
    assert_parity(out, nref)


def rnd_one_primitive_scene(rng):
    """Scenes the lane-asynchronous kernel takes: exactly one sphere / box / cylinder / gyroid, bare, as the only child of a
    collection, or as the only child of a tessellated unit cell."""
    t = rng.choice(["gyroid", "gyroid", "sphere", "box", "cylinder"])
    wrap = rng.choice(["bare", "coll", "tess", "tess"])
    cell = rng.uniform(0.35, 0.9, 3) if wrap == "tess" else np.ones(3)
    lo = rng.uniform(-0.3, 0.0, 3) if wrap == "tess" else -0.5 * np.ones(3)
    c = list(lo + 0.5 * cell + rng.uniform(-0.1, 0.1, 3) * cell)
    rho = float(rng.choice([1.0, 0.6, 1.7, -0.4]))
    s = float(cell.min())
    if t == "gyroid":
        p = {"type": t, "center": c, "scale": float(rng.uniform(0.06, 0.2)), "thickness": float(rng.uniform(0.1, 0.7)), "rho": abs(rho)}
    elif t == "sphere":
        p = {"type": t, "center": c, "radius": float(rng.uniform(0.1, 0.45) * s), "rho": rho}
    elif t == "box":
        p = {"type": t, "center": c, "sides": list(rng.uniform(0.2, 0.8, 3) * cell), "rho": rho}
    else:
        p = {"type": t, "p0": list(np.array(c) - rng.uniform(-0.4, 0.4, 3) * cell), "p1": list(np.array(c) + rng.uniform(-0.4, 0.4, 3) * cell),
             "radius": float(rng.uniform(0.05, 0.25) * s), "rho": rho}
    if wrap == "bare":
        return p
    if wrap == "coll":
        return {"type": "object_collection", "greedy_dens_eval": bool(rng.random() < 0.5), "objects": [p]}
    b = rng.uniform(0.5, 0.95, 3)
    uc = {"objects": {"objects": [p]}, "xmin": lo[0], "xmax": lo[0] + cell[0], "ymin": lo[1], "ymax": lo[1] + cell[1],
          "zmin": lo[2], "zmax": lo[2] + cell[2]}
    return {"type": "tessellated_obj_coll", "uc": uc, "xmin": -b[0], "xmax": b[0], "ymin": -b[1], "ymax": b[1], "zmin": -b[2], "zmax": b[2]}


@pytest.mark.parametrize("seed", range(80))
def test_random_one_primitive_scene(X, O, seed):
    """The per-lane march, the second-order gyroid skip bound under every warp type, refined-interval skipping and the
    deferred guard-band settling, at steps down to 1e-3 (3480 coarse samples per ray)."""
    rng = np.random.default_rng(5000 + seed)
    obj = rnd_one_primitive_scene(rng)
    deform = rnd_deform(rng)
    integ = "hierarchical" if rng.random() < 0.75 else "simple"
    res = int(rng.choice([17, 24, 32]))
    ds = float(rng.choice([0.02, 0.007, 0.003, 0.001]))
    views = ((float(rng.uniform(0, 360)), float(rng.uniform(35, 145))),)
    out, nref, _ = gpu_vs_oracle(X, O, obj, deform, views=views, res=res, integ=integ, ds=ds, ff=float(rng.choice([0.0, 0.2])),
                                 dm=float(rng.choice([1.0, 0.5, 2.0])))
    assert_parity(out, nref)
