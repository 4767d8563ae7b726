"""Development aid: soak tests/test_gpu_volume.py::test_hierarchical_volume_kernel_random.  python tools/vol_soak.py FIRST LAST"""
import os
import sys

import numpy as np
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import xray_projection_render_b200 as X  # noqa: E402
from oracle import oracle as O  # noqa: E402
import test_gpu_volume as T  # noqa: E402

first, last = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(first, last):
    try:
        T.test_hierarchical_volume_kernel_random(X, O, seed)
        if os.environ.get("SOAK_FP64"):
            T.test_fp64_volume_kernel_random(X, O, seed)
        if os.environ.get("SOAK_SIMPLE"):
            rng = np.random.default_rng(880000 + seed)
            shape = tuple(int(v) for v in rng.choice([1, 2, 3, 5, 9, 20, 31], 3))
            vol = (rng.random(shape) - (0.5 if seed % 2 else 0.0)).astype(np.float32)
            vol[rng.random(shape) < rng.choice([0.0, 0.6, 0.95])] = 0.0
            views = [(float(rng.choice([0.0, 90.0, 45.0, rng.uniform(0, 360)])), float(rng.choice([90.0, rng.uniform(30, 150)]))) for _ in range(2)]
            ds = float(rng.choice([0.05, 0.02, 0.007]))
            res = int(rng.choice([9, 16, 33]))
            cams = X.cameras_from_angles(views, 4.0, 40.0)
            osc = O.OracleScene({"type": "voxel_grid", "_array": vol.astype(np.float64)}, flat_field=0.05, density_multiplier=0.8)
            ref = np.stack([osc.render_view(*O.camera_from_angles(az, pol, 4.0), res, 40.0, 4.0, ds, "simple")[0] for az, pol in views])
            img = X.render_volume(vol, cams, res, integration="simple", precision="fp32", ds=ds, flat_field=0.05, density_multiplier=0.8)
            assert np.abs(img.astype(np.float64) - ref).max() <= 1e-4, ("simple", shape, float(np.abs(img.astype(np.float64) - ref).max()))
            img = X.render_volume(vol.astype(np.float64), cams, res, integration="hierarchical", precision="fp64", ds=ds, flat_field=0.05, density_multiplier=0.8)
            ref = np.stack([osc.render_view(*O.camera_from_angles(az, pol, 4.0), res, 40.0, 4.0, ds, "hierarchical")[0] for az, pol in views])
            assert np.abs(img.astype(np.float64) - ref).max() <= 1e-9, ("fp64 hier", shape, float(np.abs(img.astype(np.float64) - ref).max()))
    except AssertionError as e:
        bad += 1
        import traceback
        tb = traceback.extract_tb(e.__traceback__)[-1]
        print("FAIL", seed, tb.lineno, tb.line, str(e)[:200].replace("\n", " "), flush=True)
print("done", first, last, "failures", bad)
