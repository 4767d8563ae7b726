"""Development aid: split the fp64 fallbacks of a scene by cause (COUNT kernel variants, XRAY_DEBUG_FB_CAUSE)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import xray_projection_render_b200 as X  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "lattice.json"
deform = sys.argv[2] if len(sys.argv) > 2 else None
res = int(sys.argv[3]) if len(sys.argv) > 3 else 512
sc = X.Scene(str(ROOT / "tests" / "scenes" / name), str(ROOT / "tests" / "scenes" / deform) if deform else None)
views = [(90.0 + 360.0 / 16 * i + 3.3, 90.0) for i in range(16)]
cams = X.cameras_from_angles(views, 4.0, 40.0)
for cause in (0, 1, 2, 4):
    os.environ["XRAY_DEBUG_FB_CAUSE"] = str(cause)
    _, st = X.render_scene(sc, cams, res, return_stats=True)
    print(name, "cause", cause, "fallbacks", st["fp64_fallbacks"], "evaluated", st["evaluated_samples"],
          "ratio %.4f%%" % (100.0 * st["fp64_fallbacks"] / max(1, st["evaluated_samples"])))
