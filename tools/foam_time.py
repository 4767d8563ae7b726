"""Development aid: a flat object_collection of a Kelvin foam (36 n^3 struts: far more than 64 children) at 1024^2."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import xray_projection_render_b200 as X  # noqa: E402
from test_gpu_parity import _foam  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
res, nv = 1024, 8
obj = {"type": "object_collection", "objects": _foam(n, 0.02)}
sc = X.Scene(obj)
cams = X.cameras_from_angles([(91.0 + 11 * k, 80.0) for k in range(nv)], 4.0, 40.0)
out = torch.empty((nv, res, res), dtype=torch.float32, device="cuda")
ds = sc.auto_ds()
for integ in ("hierarchical", "simple"):
    for _ in range(2):
        X.render_scene_device(sc, cams, res, out, ds=ds, integration=integ)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        X.render_scene_device(sc, cams, res, out, ds=ds, integration=integ)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3 / nv
    print(f"FOAM {n}^3 cells = {len(obj['objects'])} struts, ds {ds:.5f}, {integ}: {ms:.3f} ms/view ~{res * res * (3.48 / ds) / ms / 1e6:.0f} Gsamples/s (coarse)", flush=True)
