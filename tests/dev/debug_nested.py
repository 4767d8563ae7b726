import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT/"tests"))
import xray_projection_render_b200 as X
from oracle import oracle as O
from helpers import gpu_vs_oracle
lat = json.loads((ROOT/"tests/scenes/pillar_array.json").read_text())
lat.update(xmin=-0.5, xmax=0.5, ymin=-0.5, ymax=0.5, zmin=-0.6, zmax=0.6)
vol = np.random.default_rng(4).random((6,5,7))
sph = {"type": "sphere", "center": [0.6, 0.0, 0.0], "radius": 0.3, "rho": 0.4}
box = {"type": "box", "center": [-0.6, 0.1, 0.0], "sides": [0.3, 0.5, 0.7], "rho": -0.2}
vx = {"type": "voxel_grid", "_array": vol * 0.2}
cases = {"sph+tess": [sph, lat], "tess+sph": [lat, sph], "sph+vox": [sph, vx], "sph+box": [sph, box], "tess only coll": [lat], "vox only coll":[vx],
         "all": [sph, lat, box, vx], "sph,tess,box": [sph, lat, box], "sph,box,vox":[sph,box,vx]}
for name, objs in cases.items():
    obj = {"type": "object_collection", "objects": objs}
    out, nref, ref = gpu_vs_oracle(X, O, obj, res=32, ds=0.02)
    e32, st32, img = out["fp32"]; e64, st64, _ = out["fp64"]
    print(f"{name:16s} fp32 err={e32:.3e} nan={int(np.isnan(img).sum())} ref_samples={st32['ref_samples']}/{nref} | fp64 err={e64:.3e} {st64['ref_samples']}")
