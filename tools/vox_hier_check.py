"""Development aid: voxel volume through the extended entry, simple vs hierarchical, fp32 vs fp64."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import xray_projection_render_b200 as X  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
vol = bench.synthetic_volume(n)
cams = X.cameras_from_angles([(10.0 + 45 * i, 90.0) for i in range(4)], 4.0, 40.0)
ds = 2.0 / n / 5.0
for integ, prec in (("simple", "fp32"), ("hierarchical", "fp32"), ("simple", "fp64")):
    for rep in range(2):
        t0 = time.perf_counter()
        img, st = X.render_volume(vol, cams, res, integration=integ, precision=prec, ds=ds, return_stats=True)
        dt = time.perf_counter() - t0
    print(integ, prec, "%.1f ms" % (dt * 1e3), "ref Gs/s %.1f" % (st["ref_samples"] / dt / 1e9), "eval", st["evaluated_samples"], "fb", st["fp64_fallbacks"], flush=True)
