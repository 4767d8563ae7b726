"""Development aid: time the analytic BASELINE configs through the device-output entry point (kernels only).
python tools/span_time.py [lattice pillar ...]"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import xray_projection_render_b200 as X  # noqa: E402
from oracle import oracle as O  # noqa: E402

SC = ROOT / "tests" / "scenes"
CFG = {"lattice": ("lattice.json", None, 1024, 32), "pillar": ("pillar_array.json", None, 4096, 4), "cube": ("cube_w_hole.json", None, 512, 8),
       "box_w_pped": ("box_w_pped.json", None, 1024, 32), "balls": ("balls.json", None, 1024, 32),
       "lattice_linear": ("lattice.json", "deformation_linear.json", 1024, 32)}
names = [a for a in sys.argv[1:] if a in CFG] or ["lattice", "pillar"]
for name in names:
    obj, deform, res, nv = CFG[name]
    sc = X.Scene(str(SC / obj), str(SC / deform) if deform else None)
    cams = X.cameras_from_angles([(90.0 + k, 90.0) for k in range(nv)], 4.0, 40.0)
    out = torch.empty((nv, res, res), dtype=torch.float32, device="cuda")
    ds = sc.auto_ds()
    for _ in range(2):
        X.render_scene_device(sc, cams, res, out, ds=ds)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        X.render_scene_device(sc, cams, res, out, ds=ds, stream=torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    K = O.step_count("hierarchical", ds, 4.0 - 1.74, 4.0 + 1.74)
    tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("XRAY_"))
    print(f"TIME {name:14s} {ms / nv:.4f} ms/view  ~{nv * res * res * K / ms / 1e6:.0f} Gsamples/s (coarse)  [{tag}]", flush=True)
