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
cams = X.cameras_from_angles([(90.0 + k, 90.0) for k in range(nv)], 4.0, 40.0)
out = torch.empty((nv, res, res), dtype=torch.float32, device="cuda")
if obj.startswith("voxel"):  # voxel1024 / voxel256 ...: the synthetic BASELINE volume, device resident
    import bench

    n = int(obj[5:])
    vol = torch.from_numpy(bench.synthetic_volume(n)).cuda()
    for _ in range(reps):
        X.render_volume_device(vol, (n, n, n), cams, res, out, ds=2.0 / n / 5.0, integration=__import__('os').environ.get('INTEG', 'simple'))
else:
    sc = X.Scene(str(SC / obj), str(SC / deform) if deform else None)
    for _ in range(reps):
        X.render_scene_device(sc, cams, res, out, ds=sc.auto_ds())
torch.cuda.synchronize()
print("done", float(out.mean()))
