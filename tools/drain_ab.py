"""Development aid: end-to-end time through XRayRenderSceneCUDA with PAGEABLE host images (what a cgo / ctypes caller
hands over), A/B over the host-side drain settings (XRAY_NO_STREAM_COPY, XRAY_DRAIN_THREADS).
python tools/drain_ab.py [pillar lattice]"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402,F401  (device context)

import xray_projection_render_b200 as X  # noqa: E402

SC = ROOT / "tests" / "scenes"
CFG = {"pillar": ("pillar_array.json", 4096, 8), "lattice": ("lattice.json", 1024, 128)}
SETTINGS = [("memcpy, default threads", {"XRAY_NO_STREAM_COPY": "1"}), ("stream, default threads", {}),
            ("stream, 4 threads", {"XRAY_DRAIN_THREADS": "4"}), ("stream, 12 threads", {"XRAY_DRAIN_THREADS": "12"}),
            ("stream, 16 threads", {"XRAY_DRAIN_THREADS": "16"}), ("memcpy, 16 threads", {"XRAY_NO_STREAM_COPY": "1", "XRAY_DRAIN_THREADS": "16"})]
names = [a for a in sys.argv[1:] if a in CFG] or list(CFG)
print("host cores", os.cpu_count(), flush=True)
for name in names:
    obj, res, nv = CFG[name]
    sc = X.Scene(str(SC / obj), None)
    cams = X.cameras_from_angles([(90.0 + k, 90.0) for k in range(nv)], 4.0, 40.0)
    ds = sc.auto_ds()
    out = np.zeros((nv, res, res), dtype=np.float32)  # pageable, pages touched
    ref = None
    for label, env in SETTINGS:
        for k in ("XRAY_NO_STREAM_COPY", "XRAY_DRAIN_THREADS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        best = 1e9
        for rep in range(5):
            out[:, ::64, ::64] = -1.0
            t0 = time.perf_counter()
            X.render_scene(sc, cams, res, ds=ds, out=out)
            best = min(best, time.perf_counter() - t0)
        if ref is None:
            ref = out.copy()
        same = bool(np.array_equal(ref, out))
        print(f"DRAIN {name:8s} {label:26s} {best * 1e3:8.2f} ms  {out.nbytes / best / 1e9:6.1f} GB/s of images  identical={same}", flush=True)
