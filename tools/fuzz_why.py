"""Development aid: details of failing seeds of tests/test_gpu_fuzz.py::test_random_scene.  python tools/fuzz_why.py SEED..."""
import json
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
from helpers import gpu_vs_oracle  # noqa: E402


def summarize(o, depth=0):
    if isinstance(o, dict):
        t = o.get("type", "uc" if "uc" in o else "?")
        if "objects" in o and isinstance(o["objects"], list):
            return f"{t}[{', '.join(summarize(k) for k in o['objects'])}]"
        if "uc" in o:
            return f"tess({summarize(o['uc']['objects'])})"
        if "objects" in o:
            return summarize(o["objects"])
        return t
    return "?"


for seed in (int(a) for a in sys.argv[1:]):
    rng = np.random.default_rng(1000 + seed)
    obj = T.rnd_scene(rng)
    deform = T.rnd_deform(rng)
    integ = "hierarchical" if rng.random() < 0.7 else "simple"
    res = int(rng.choice([17, 24, 32]))
    ds = float(rng.choice([0.03, 0.017, 0.01]))
    views = ((float(rng.uniform(0, 360)), float(rng.uniform(35, 145))),)
    ff = float(rng.choice([0.0, 0.2]))
    dm = float(rng.choice([1.0, 0.5, 2.0]))
    print(f"seed {seed}: {summarize(obj)} deform {deform and deform['type']} {integ} res {res} ds {ds} views {views} ff {ff} dm {dm}")
    for no_span in (False, True):
        for knob in ({}, {"XRAY_GENERIC_KERNEL": "1"}, {"XRAY_NO_SKIP": "1"}):
            for k in ("XRAY_NO_SPAN", "XRAY_GENERIC_KERNEL", "XRAY_NO_SKIP"):
                os.environ.pop(k, None)
            if no_span:
                os.environ["XRAY_NO_SPAN"] = "1"
            os.environ.update(knob)
            out, nref, ref = gpu_vs_oracle(X, O, obj, deform, views=views, res=res, integ=integ, ds=ds, ff=ff, dm=dm)
            line = f"   no_span={no_span} {knob}:"
            for prec in ("fp32", "fp64"):
                err, st, img = out[prec]
                d = np.abs(img.astype(np.float64) - ref)
                nb = int((d > (1e-4 if prec == "fp32" else 1e-9)).sum())
                line += f" {prec} err {err:.3e} nbad {nb} refs {st['ref_samples']}/{nref} span {st['span_renderer']} fb {st['fp64_fallbacks']} launches {st['launches']};"
            print(line, flush=True)
    if os.environ.get("DUMP"):
        json.dump({"obj": obj, "deform": deform}, open(f"gpurun_out/fuzz_{seed}.json", "w"), default=lambda a: a.tolist())
