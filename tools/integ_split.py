"""Development aid: what the refinement replay costs -- the same scene and step with the simple and the hierarchical
integrator (kernel-resident timing, CUDA events)."""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import xray_projection_render_b200 as X  # noqa: E402

for name, deform in (("lattice", None), ("pillar_array", None), ("gyroid_example", "deformation_sigmoid")):
    sc = X.Scene(str(ROOT / "tests/scenes" / f"{name}.json"), str(ROOT / "tests/scenes" / f"{deform}.json") if deform else None)
    res, nv = (1024, 24) if name != "pillar_array" else (4096, 4)
    cams = X.cameras_from_angles(X.generate_camera_angles(nv), 4.0, 40.0)
    out = torch.empty((nv, res, res), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream()
    for integ in ("simple", "hierarchical"):
        stats = (ctypes.c_uint64 * X._lib.XRAY_NUM_STATS)()
        X.render_scene_device(sc, cams, res, out, integration=integ, precision="fp32", stream=st.cuda_stream, stats=stats)
        for _ in range(2):
            X.render_scene_device(sc, cams, res, out, integration=integ, precision="fp32", stream=st.cuda_stream)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(3):
            X.render_scene_device(sc, cams, res, out, integration=integ, precision="fp32", stream=st.cuda_stream)
        b.record(st)
        torch.cuda.synchronize()
        rays = stats[4]
        print(f"{name:15s} {integ:12s} {a.elapsed_time(b) / 3:8.2f} ms  ref/ray {stats[0] / rays:7.1f} evaluated/ray {stats[1] / rays:6.1f} "
              f"prim tests/ray {stats[3] / rays:6.1f} fp64/ray {stats[2] / rays:5.2f}")
