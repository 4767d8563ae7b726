"""Development aid: device-resident volume render timings for the integrator / precision combinations.
python tools/vol_time.py [N=512] [res=1024] [views=8]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
import xray_projection_render_b200 as X  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nv = int(sys.argv[3]) if len(sys.argv) > 3 else 8
vol = torch.from_numpy(bench.synthetic_volume(n)).cuda()
cams = X.cameras_from_angles([(90.0 + 7 * k, 90.0) for k in range(nv)], 4.0, 40.0)
ds = 2.0 / n / 5.0
for integ in ("simple", "hierarchical"):
    for prec in ("fp32", "fp64"):
        out = torch.empty((nv, res, res), dtype=torch.float32, device="cuda")
        import ctypes
        stats = (ctypes.c_uint64 * X._lib.XRAY_NUM_STATS)()
        for _ in range(2):
            X.render_volume_device(vol, (n, n, n), cams, res, out, ds=ds, integration=integ, precision=prec, stats=stats)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            X.render_volume_device(vol, (n, n, n), cams, res, out, ds=ds, integration=integ, precision=prec)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3 / nv
        coarse = res * res * (3.48 / ds)
        print(f"VOL {n}^3 res {res} {integ:12s} {prec}: {ms:8.3f} ms/view ~{coarse / ms / 1e6:8.1f} Gsamples/s (coarse)  ref_samples/view {stats[0] / nv:.3e} fallbacks {stats[2]}", flush=True)
