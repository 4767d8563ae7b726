"""Development aid: one scene, span path vs marching kernels, list the differing pixels per view."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import xray_projection_render_b200 as X  # noqa: E402

uc = {"objects": {"objects": [{"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 0.12, "rho": 0.9},
                              {"type": "cylinder", "p0": [0.0, 0.0, 0.0], "p1": [0.3, 0.0, 0.0], "radius": 0.04, "rho": 0.5},
                              {"type": "cylinder", "p0": [0.0, 0.0, 0.0], "p1": [0.0, 0.3, 0.0], "radius": 0.05, "rho": 0.4},
                              {"type": "cylinder", "p0": [0.0, 0.0, 0.0], "p1": [0.0, 0.0, 0.3], "radius": 0.03, "rho": 0.3},
                              {"type": "box", "center": [0.15, 0.0, 0.3], "sides": [0.1, 0.1, 0.1], "rho": 0.7}]},
      "xmin": 0.0, "xmax": 0.3, "ymin": 0.0, "ymax": 0.3, "zmin": 0.0, "zmax": 0.3}
obj = {"type": "tessellated_obj_coll", "uc": uc, "xmin": -0.75, "xmax": 0.75, "ymin": -0.6, "ymax": 0.6, "zmin": -0.45, "zmax": 0.9}
views = ((0.0, 90.0), (90.0, 90.0), (180.0, 90.0), (270.0, 90.0), (0.0, 0.0001), (0.0, 179.9999))
if os.environ.get("DBG_VIEWS"):
    views = tuple(views[int(k)] for k in os.environ["DBG_VIEWS"].split(","))
res = int(sys.argv[1]) if len(sys.argv) > 1 else 32
keep = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else range(len(uc["objects"]["objects"]))
uc["objects"]["objects"] = [uc["objects"]["objects"][k] for k in keep]
for integ in os.environ.get("DBG_INTEG", "simple,hierarchical").split(","):
    for v in views:
        cams = X.cameras_from_angles([v], 4.0, 40.0)
        os.environ.pop("XRAY_NO_SPAN", None)
        a, sa = X.render_scene(X.Scene(obj), cams, res, integration=integ, precision="fp64", ds=0.011, return_stats=True)
        os.environ["XRAY_NO_SPAN"] = "1"
        b, sb = X.render_scene(X.Scene(obj), cams, res, integration=integ, precision="fp64", ds=0.011, return_stats=True)
        d = np.abs(a.astype(np.float64) - b)
        bad = np.argwhere(d > 1e-9)
        print(integ, v, "max", d.max(), "nbad", len(bad), "marched", sa["marched_tiles"], hex(sa["march_reasons"]), "nref", sa["ref_samples"], sb["ref_samples"])
        for q in bad[:12]:
            print("   ", tuple(q), a[tuple(q)], b[tuple(q)])
if os.environ.get("XRAY_DEBUG_FB_CAUSE") == "77":
    os.environ.pop("XRAY_NO_SPAN", None)
    for v in views[4:]:
        cams = X.cameras_from_angles([v], 4.0, 40.0)
        a, sa = X.render_scene(X.Scene(obj), cams, res, integration="simple", precision="fp64", ds=0.011, return_stats=True)
        np.set_printoptions(linewidth=250, precision=2, suppress=True)
        print(v)
        print(np.where(a[0] < 0, a[0], 0).astype(int))
