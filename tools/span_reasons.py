"""Development aid: how many tiles the interval renderer hands over, and why (stats[6], stats[7])."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import xray_projection_render_b200 as X  # noqa: E402

SC = ROOT / "tests" / "scenes"
for name, res, views in (("cube_w_hole", 40, [(77.0, 83.0)]), ("cube_w_hole", 40, [(200.0, 101.0)]), ("cube_w_hole", 512, [(90.0, 90.0)]),
                         ("lattice", 1024, [(90.0, 90.0)]), ("lattice", 1024, [(91.0, 90.0)]), ("lattice", 1024, [(95.0, 90.0)]),
                         ("lattice", 1024, [(135.0, 90.0)]), ("lattice", 1024, [(95.0, 80.0)]), ("pillar_array", 4096, [(91.0, 90.0)]),
                         ("pillar_array", 1024, [(90.0, 90.0)]), ("box_w_pped", 1024, [(91.0, 90.0)]), ("balls", 1024, [(91.0, 90.0)])):
    sc = X.Scene(str(SC / f"{name}.json"))
    cams = X.cameras_from_angles(views, 4.0, 40.0)
    for integ in ("hierarchical", "simple"):
        _, st = X.render_scene(sc, cams, res, integration=integ, return_stats=True)
        print(f"{name:14s} res {res:5d} view {views[0]} {integ:12s}: marched_tiles {st['marched_tiles']:5d} of {((res + 7) // 8) * ((res + 15) // 16)} reasons {st['march_reasons']:#x} "
              f"fallbacks {st['fp64_fallbacks']}", flush=True)
