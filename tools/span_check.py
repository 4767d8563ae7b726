"""Development aid: interval ("span") renderer versus the marching kernels and the oracle on the bundled scenes,
plus a quick timing of the BASELINE analytic configs.  Run on the GPU box: python tools/span_check.py [--time]"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import xray_projection_render_b200 as X  # noqa: E402
from oracle import oracle as O  # noqa: E402

SC = ROOT / "tests" / "scenes"
R, FOV = 4.0, 40.0


def run(obj, deform, views, res, integ, prec, span, ds=-1.0, dm=1.0, ff=0.0):
    if span:
        os.environ.pop("XRAY_NO_SPAN", None)
    else:
        os.environ["XRAY_NO_SPAN"] = "1"
    sc = X.Scene(obj, deform)
    cams = X.cameras_from_angles(views, R, FOV)
    return X.render_scene(sc, cams, res, integration=integ, precision=prec, ds=ds if ds > 0 else sc.auto_ds(), flat_field=ff,
                          density_multiplier=dm, return_stats=True)


def main():
    views = [(90.0, 90.0), (131.0, 70.0), (17.0, 100.0)]
    cases = [("cube_w_hole.json", None), ("lattice.json", None), ("pillar_array.json", None), ("balls.json", None),
             ("box_w_pped.json", None), ("lattice.json", "deformation_linear.json"), ("cube_w_hole.json", "deformation_linear.json")]
    worst = 0.0
    for obj, deform in cases:
        of, df = str(SC / obj), (str(SC / deform) if deform else None)
        osc = O.OracleScene(of, df, flat_field=0.01, density_multiplier=1.3)
        sc = X.Scene(of, df)
        ds = sc.auto_ds()
        for integ in ("hierarchical", "simple"):
            for res in (48, 37):
                ref, nref = [], 0
                for az, pol in views:
                    eye, cm = O.camera_from_angles(az, pol, R)
                    im, k = osc.render_view(eye, cm, res, FOV, R, ds, integ)
                    ref.append(im)
                    nref += k
                ref = np.stack(ref)
                for prec in ("fp32", "fp64"):
                    a, sa = run(of, df, views, res, integ, prec, True, ds, 1.3, 0.01)
                    b, sb = run(of, df, views, res, integ, prec, False, ds, 1.3, 0.01)
                    ea = float(np.abs(a.astype(np.float64) - ref).max())
                    eb = float(np.abs(b.astype(np.float64) - ref).max())
                    ok = ea <= (1e-4 if prec == "fp32" else 1e-9) and sa["ref_samples"] == nref
                    worst = max(worst, ea)
                    print(f"{'OK ' if ok else 'BAD'} {obj:18s} {str(deform):26s} {integ:12s} res {res} {prec}: span err {ea:.2e} (march {eb:.2e}) "
                          f"ref_samples {sa['ref_samples']} vs oracle {nref} marched_tiles {sa['marched_tiles']} launches {sa['launches']} "
                          f"intervals {sa['evaluated_samples']} cands {sa['primitive_tests']}", flush=True)
                    if not ok:
                        bad = np.argwhere(np.abs(a.astype(np.float64) - ref) > 1e-6)
                        print("   first bad pixels (view,i,j):", bad[:8].tolist(), " n_bad", len(bad))
    print("worst", worst)
    if "--time" in sys.argv:
        import torch

        for name, obj, deform, res, nv in (("lattice", "lattice.json", None, 1024, 24), ("pillar", "pillar_array.json", None, 4096, 4),
                                           ("cube", "cube_w_hole.json", None, 512, 8), ("box_w_pped", "box_w_pped.json", None, 1024, 24),
                                           ("balls", "balls.json", None, 1024, 24), ("lattice_linear", "lattice.json", "deformation_linear.json", 1024, 24)):
            sc = X.Scene(str(SC / obj), str(SC / deform) if deform else None)
            vs = [(90.0 + 360.0 * k / 360, 90.0) for k in range(nv)]
            cams = X.cameras_from_angles(vs, R, FOV)
            out = torch.empty((nv, res, res), dtype=torch.float32, device="cuda")
            for span in (True, False):
                if span:
                    os.environ.pop("XRAY_NO_SPAN", None)
                else:
                    os.environ["XRAY_NO_SPAN"] = "1"
                for _ in range(2):
                    X.render_scene_device(sc, cams, res, out, ds=sc.auto_ds())
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    X.render_scene_device(sc, cams, res, out, ds=sc.auto_ds(), stream=torch.cuda.current_stream().cuda_stream)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                K = O.step_count("hierarchical", sc.auto_ds(), R - 1.74, R + 1.74)
                print(f"TIME {name:14s} span={span}: {ms / nv:.3f} ms/view, ~{nv * res * res * K / ms / 1e6:.0f} Gsamples/s (coarse only)", flush=True)


if __name__ == "__main__":
    main()
