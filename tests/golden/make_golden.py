"""Generate tests/golden/golden_v1.{npz,json}: small oracle renders of the bundled scenes.

The reference is Go and cannot run in this image (no Go toolchain), so these vectors come from the
CPU oracle (oracle/xray_oracle.cpp) after it was pinned against the reference's own known-answer
tests.  They freeze today's oracle output so that any later drift of the oracle OR of the GPU path
is caught.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

SCENES = ROOT / "tests" / "scenes"
OUT = Path(__file__).resolve().parent

CASES = [
    # key, object, deformation, integration, ds (<=0 auto), flat_field, density_multiplier, res, views
    ("cube_hier", "cube_w_hole.json", None, "hierarchical", -1, 0.0, 1.0, 32, [(90.0, 90.0), (201.0, 64.0)]),
    ("cube_simple_ff", "cube_w_hole.json", None, "simple", 0.01, 0.2, 1.5, 24, [(33.0, 90.0)]),
    ("balls_hier", "balls.json", None, "hierarchical", -1, 0.0, 1.0, 32, [(135.0, 90.0)]),
    ("pped_hier", "box_w_pped.json", None, "hierarchical", -1, 0.0, 1.0, 28, [(250.0, 75.0)]),
    ("pillar_hier", "pillar_array.json", None, "hierarchical", -1, 0.0, 1.0, 32, [(90.0, 90.0), (117.0, 50.0)]),
    ("lattice_hier", "lattice.json", None, "hierarchical", -1, 0.0, 1.0, 24, [(90.0, 90.0), (300.0, 100.0)]),
    ("gyroid_sigmoid", "gyroid_example.json", "deformation_sigmoid.json", "hierarchical", 0.004, 0.0, 1.0, 20, [(90.0, 90.0)]),
    ("cube_linear", "cube_w_hole.json", "deformation_linear.json", "hierarchical", -1, 0.0, 1.0, 24, [(10.0, 110.0)]),
]
R, FOV = 4.0, 40.0

arrays, meta = {}, {"R": R, "fov": FOV, "cases": []}
for key, obj, deform, integ, ds, ff, dm, res, views in CASES:
    osc = O.OracleScene(str(SCENES / obj), str(SCENES / deform) if deform else None, flat_field=ff, density_multiplier=dm)
    if ds <= 0:
        ds = osc.auto_ds()
    imgs = []
    nsamp = 0
    for az, pol in views:
        eye, cm = O.camera_from_angles(az, pol, R)
        im, n = osc.render_view(eye, cm, res, FOV, R, ds, integ)
        imgs.append(im)
        nsamp += n
    arrays[key] = np.stack(imgs)
    meta["cases"].append({"key": key, "object": obj, "deformation": deform, "integration": integ, "ds": ds, "flat_field": ff,
                          "density_multiplier": dm, "res": res, "views": views, "R": R, "fov": FOV, "ref_samples": nsamp})
np.savez_compressed(OUT / "golden_v1.npz", **arrays)
(OUT / "golden_v1.json").write_text(json.dumps(meta, indent=1))
print("wrote", sum(a.nbytes for a in arrays.values()), "bytes of images for", len(CASES), "cases")
