"""Development aid (dev build, XRAY_DEBUG_FB_CAUSE=77): list the pixels the interval renderer hands over and the reason codes."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ["XRAY_CUDA_LIB"] = str(ROOT / "xray_projection_render_b200" / "lib_dev" / "libcuda_render.so")
os.environ["XRAY_DEBUG_FB_CAUSE"] = "77"
import numpy as np  # noqa: E402

import xray_projection_render_b200 as X  # noqa: E402

SC = ROOT / "tests" / "scenes"
for name, res, views in (("cube_w_hole", 40, [(77.0, 83.0)]), ("cube_w_hole", 40, [(200.0, 101.0)]), ("lattice", 1024, [(91.0, 90.0)]),
                         ("lattice", 1024, [(135.0, 90.0)]), ("pillar_array", 1024, [(90.0, 90.0)])):
    sc = X.Scene(str(SC / f"{name}.json"))
    cams = X.cameras_from_angles(views, 4.0, 40.0)
    for integ in ("hierarchical", "simple"):
        img = X.render_scene(sc, cams, res, integration=integ)
        bad = np.argwhere(img[0] < 0)
        codes = sorted(set((-img[0][img[0] < 0]).astype(int).tolist()))
        print(name, res, views[0], integ, "bad pixels", len(bad), "codes", codes, "first", bad[:12].tolist(), flush=True)
