"""Development aid: per-ray averages of the interval renderer's counters (candidates that reach the exact stage, intervals)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import xray_projection_render_b200 as X  # noqa: E402

SC = ROOT / "tests" / "scenes"
for name, res, nv in (("lattice", 1024, 8), ("pillar_array", 2048, 2), ("box_w_pped", 1024, 4), ("cube_w_hole", 512, 4)):
    sc = X.Scene(str(SC / f"{name}.json"))
    cams = X.cameras_from_angles([(91.0 + 13 * k, 90.0) for k in range(nv)], 4.0, 40.0)
    _, st = X.render_scene(sc, cams, res, return_stats=True)
    rays = max(1, st["rays"])
    print(f"{name:14s}: rays {rays} candidates/ray {st['primitive_tests'] / rays:.2f} intervals/ray {st['evaluated_samples'] / rays:.2f} "
          f"ref samples/ray {st['ref_samples'] / rays:.0f} marched_tiles {st['marched_tiles']}")
