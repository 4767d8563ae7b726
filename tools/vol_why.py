import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import xray_projection_render_b200 as X
from oracle import oracle as O
from helpers import R, FOV
for seed in (int(a) for a in sys.argv[1:]):
    rng = np.random.default_rng(4400 + seed)
    shape = tuple(int(v) for v in rng.choice([1, 2, 3, 7, 16, 24, 33], 3))
    vol = rng.random(shape, dtype=np.float32)
    if seed % 2:
        vol -= np.float32(0.5)
    vol[rng.random(shape) < rng.choice([0.0, 0.5, 0.9, 0.99])] = 0.0
    if seed % 3 == 0:
        vol = np.round(vol * 4).astype(np.float32)
    views = [(float(rng.choice([0.0, 90.0, 45.0, rng.uniform(0, 360)])), float(rng.choice([90.0, rng.uniform(30, 150)]))) for _ in range(2)]
    ds = float(rng.choice([0.05, 0.02, 0.011]))
    dm = float(rng.choice([1.0, 0.6, -0.7, 0.0]))
    res = int(rng.choice([16, 21, 32])); ff = float(rng.choice([0.0, 0.1]))
    print("seed", seed, "shape", shape, "nonzero", int((vol != 0).sum()), "views", views, "ds", ds, "dm", dm, "res", res, "ff", ff)
    cams = X.cameras_from_angles(views, R, FOV)
    osc = O.OracleScene({"type": "voxel_grid", "_array": vol.astype(np.float64)}, flat_field=ff, density_multiplier=dm)
    for integ in ("simple", "hierarchical"):
        ref, nref = [], 0
        for az, pol in views:
            im, k = osc.render_view(*O.camera_from_angles(az, pol, R), res, FOV, R, ds, integ)
            ref.append(im); nref += k
        ref = np.stack(ref)
        for gen in (False, True):
            if gen: os.environ["XRAY_VOLUME_GENERIC"] = "1"
            else: os.environ.pop("XRAY_VOLUME_GENERIC", None)
            img, st = X.render_volume(vol, cams, res, integration=integ, precision="fp32", ds=ds, flat_field=ff, density_multiplier=dm, return_stats=True)
            d = np.abs(img.astype(np.float64) - ref)
            print(f"   {integ:12s} generic={gen}: err {d.max():.3e} nbad {(d > 1e-4).sum()} refs {st['ref_samples']}/{nref} range [{ref.min():.3g},{ref.max():.3g}]")
            if d.max() > 1e-4 and not gen:
                for q in np.argwhere(d > 1e-4)[:6]:
                    print("      ", tuple(q), img[tuple(q)], ref[tuple(q)])
