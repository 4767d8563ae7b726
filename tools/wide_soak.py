"""Development aid: randomised scenes with randomised CAMERAS (distance, field of view), step sizes down to 0.0015 and detector
sizes 5 ... 70, several views per call, against the oracle; both kernel paths.  python tools/wide_soak.py FIRST LAST"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import xray_projection_render_b200 as X  # noqa: E402
from oracle import oracle as O  # noqa: E402
import test_gpu_fuzz as T  # noqa: E402
import test_gpu_span as S  # noqa: E402
from helpers import oracle_images  # noqa: E402

first, last = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(first, last):
    rng = np.random.default_rng(555000 + seed)
    kind = rng.choice(["fuzz", "span", "one"])
    obj = T.rnd_scene(rng) if kind == "fuzz" else (S._span_scene(rng) if kind == "span" else T.rnd_one_primitive_scene(rng))
    def sparsify(o):  # voxel children with exact zeros (and sometimes mixed signs): zero-ness matters to the hierarchical integrator
        if isinstance(o, dict):
            if o.get("type") == "voxel_grid" and "_array" in o:
                arr = np.array(o["_array"], dtype=np.float64)
                if rng.random() < 0.5:
                    arr -= 0.25
                arr[rng.random(arr.shape) < rng.choice([0.3, 0.7, 0.95])] = 0.0
                o["_array"] = arr
            for v in list(o.values()):
                sparsify(v)
        elif isinstance(o, list):
            for v in o:
                sparsify(v)
    if os.environ.get("SOAK_NESTED"):
        kids = [T.rnd_prim(rng, allow_gyroid=False) for _ in range(int(rng.integers(0, 3)))]
        kids.append({"type": "voxel_grid", "_array": rng.random((int(rng.integers(1, 12)), int(rng.integers(1, 12)), int(rng.integers(1, 12)))) * 0.6})
        if rng.random() < 0.4:
            kids.append(T.rnd_tess(rng))
        rng.shuffle(kids)
        obj = {"type": "object_collection", "greedy_dens_eval": bool(rng.random() < 0.3), "objects": kids}
    sparsify(obj)
    deform = T.rnd_deform(rng) if rng.random() < 0.5 else None
    integ = "hierarchical" if rng.random() < 0.7 else "simple"
    res = int(rng.integers(5, 71))
    Rcam = float(rng.choice([4.0, 2.6, 3.0, 6.0, 9.0]))
    fov = float(rng.choice([40.0, 12.0, 25.0, 60.0, 85.0]))
    ds = float(10.0 ** rng.uniform(np.log10(0.0015), np.log10(0.05)))
    nv = int(rng.integers(1, 4))
    views = tuple((float(rng.choice([0.0, 90.0, 45.0, rng.uniform(0, 360)])), float(rng.choice([90.0, rng.uniform(20, 160)]))) for _ in range(nv))
    ff = float(rng.choice([0.0, 0.2]))
    dm = float(rng.choice([1.0, 0.5, 2.0]))
    try:
        sc = X.Scene(obj, deform)
        osc = O.OracleScene(obj, deform, flat_field=ff, density_multiplier=dm)
        ref, nref = oracle_images(O, osc, views, res, ds, integ, R=Rcam, fov=fov)
        cams = X.cameras_from_angles(views, Rcam, fov)
    except Exception as e:  # scene the front end refuses: not what this soak is about
        print("SKIP", seed, type(e).__name__, str(e)[:100], flush=True)
        continue
    for no_span in (False, True):
        if no_span:
            os.environ["XRAY_NO_SPAN"] = "1"
        else:
            os.environ.pop("XRAY_NO_SPAN", None)
        for prec, tol in (("fp32", 1e-4), ("fp64", 1e-9)):
            try:
                img, st = X.render_scene(sc, cams, res, integration=integ, precision=prec, ds=ds, flat_field=ff, density_multiplier=dm,
                                         return_stats=True)
            except Exception as e:
                print("ERR ", seed, kind, prec, type(e).__name__, str(e)[:160], flush=True)
                bad += 1
                continue
            err = float(np.abs(img.astype(np.float64) - ref).max())
            if err > tol or st["ref_samples"] != nref:
                bad += 1
                print(f"FAIL seed {seed} {kind} no_span={no_span} {prec} err {err:.3e} refs {st['ref_samples']}/{nref} R {Rcam} fov {fov} ds {ds:.5f} res {res} "
                      f"{integ} deform {deform and deform['type']} views {views}", flush=True)
print("done", first, last, "failures", bad)
