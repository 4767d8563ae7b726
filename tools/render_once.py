"""Development aid for ncu captures: render N views of a bundled scene into a device buffer, a few times.
python tools/render_once.py lattice.json 1024 8 [deformation.json] [reps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import xray_projection_render_b200 as X  # noqa: E402

SC = ROOT / "tests" / "scenes"
obj, res, nv = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
deform = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
sc = X.Scene(str(SC / obj), str(SC / deform) if deform else None)
cams = X.cameras_from_angles([(90.0 + k, 90.0) for k in range(nv)], 4.0, 40.0)
out = torch.empty((nv, res, res), dtype=torch.float32, device="cuda")
for _ in range(reps):
    X.render_scene_device(sc, cams, res, out, ds=sc.auto_ds())
torch.cuda.synchronize()
print("done", float(out.mean()))
