"""Correctness + timing of large explicit collections (n > 64 children)."""
import sys, time, json
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import xray_projection_render_b200 as X
from oracle import oracle as O
from helpers import gpu_vs_oracle, assert_parity
from test_oracle_known_answers import kelvin

def foam(n_cells, rad):
    """Explicit object_collection of a Kelvin foam, n_cells^3 unit cells filling [-0.8, 0.8]^3."""
    size = 1.6 / n_cells
    uc = kelvin(rad, size)["objects"]["objects"]
    objs = []
    for a in range(n_cells):
        for b in range(n_cells):
            for c in range(n_cells):
                off = np.array([a, b, c]) * size - 0.8
                for o in uc:
                    objs.append({"type": "cylinder", "p0": list(np.array(o["p0"]) + off), "p1": list(np.array(o["p1"]) + off),
                                 "radius": rad, "rho": 1.0})
    return {"type": "object_collection", "objects": objs}

for n_cells, res in ((2, 24), (4, 16)):
    obj = foam(n_cells, 0.02)
    t0 = time.time()
    out, nref, _ = gpu_vs_oracle(X, O, obj, res=res, ds=0.01, views=((100.0, 80.0),))
    assert_parity(out, nref)
    print(f"foam {n_cells}^3: {len(obj['objects'])} cylinders parity ok fp32={out['fp32'][0]:.2e} fp64={out['fp64'][0]:.2e} "
          f"prim_tests={out['fp32'][1]['primitive_tests']} eval={out['fp32'][1]['evaluated_samples']} ({time.time()-t0:.1f}s)")
    sc = X.Scene(obj)
    cams = X.cameras_from_angles(X.generate_camera_angles(4), 4.0, 40.0)
    X.render_scene(sc, cams, 512)
    t0 = time.time(); X.render_scene(sc, cams, 512); dt = time.time() - t0
    print(f"   4 views at 512^2, auto ds={sc.auto_ds():.4f}: {dt*1e3:.1f} ms")
