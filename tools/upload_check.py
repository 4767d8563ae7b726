"""Development aid: legacy voxel call with a PAGEABLE 4 GiB volume, staged vs plain upload."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import xray_projection_render_b200 as X  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
vol = np.random.default_rng(0).random((n, n, n), dtype=np.float32)
cams32 = X.to_legacy(X.cameras_from_angles([(10.0, 90.0)], 4.0, 40.0))
ds = float(np.float32(2.0 / n / 5.0))
for mode in ("staged", "plain", "staged", "plain"):
    if mode == "plain":
        os.environ["XRAY_NO_STAGED_UPLOAD"] = "1"
    else:
        os.environ.pop("XRAY_NO_STAGED_UPLOAD", None)
    t0 = time.perf_counter()
    X.render_volume_legacy(vol, cams32, 256, ds)
    print(mode, "%.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
