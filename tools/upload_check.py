"""Development aid: legacy voxel call with a PAGEABLE volume, staged vs plain upload, optionally with torch loaded."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
if len(sys.argv) > 2 and sys.argv[2] == "torch":
    import torch  # noqa: F401
    torch.zeros(1, device="cuda")
import xray_projection_render_b200 as X  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
res = int(sys.argv[3]) if len(sys.argv) > 3 else 256
nviews = int(sys.argv[4]) if len(sys.argv) > 4 else 1
vol = np.random.default_rng(0).random((n, n, n), dtype=np.float32)
cams32 = X.to_legacy(X.cameras_from_angles([(10.0 + 5 * i, 90.0) for i in range(nviews)], 4.0, 40.0))
ds = float(np.float32(2.0 / n / 5.0))
out = np.empty((nviews, res, res), dtype=np.float32)
for mode in ("staged", "plain", "staged", "staged"):
    if mode == "plain":
        os.environ["XRAY_NO_STAGED_UPLOAD"] = "1"
    else:
        os.environ.pop("XRAY_NO_STAGED_UPLOAD", None)
    t0 = time.perf_counter()
    X.render_volume_legacy(vol, cams32, res, ds, out=out)
    print(mode, "%.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
