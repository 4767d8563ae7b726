"""Quick GPU parity + timing sweep (development aid; the judged checks live in tests/ -m gpu)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import xray_projection_render_b200 as X  # noqa: E402
from oracle import oracle as O  # noqa: E402

EX = ROOT / "tests" / "scenes"


def check(name, obj, deform=None, res=64, views=((90.0, 90.0), (131.0, 70.0)), integ="hierarchical", ds=-1.0, ff=0.0, dm=1.0):
    sc = X.Scene(str(EX / obj), str(EX / deform) if deform else None)
    osc = O.OracleScene(str(EX / obj), str(EX / deform) if deform else None, flat_field=ff, density_multiplier=dm)
    if ds <= 0:
        ds = sc.auto_ds()
    cams = X.cameras_from_angles(views, 4.0, 40.0)
    ref = np.zeros((len(views), res, res))
    nref = 0
    for v, (az, pol) in enumerate(views):
        eye, cm = O.camera_from_angles(az, pol, 4.0)
        ref[v], n = osc.render_view(eye, cm, res, 40.0, 4.0, ds, integ)
        nref += n
    out = {}
    for prec in ("fp64", "fp32"):
        t0 = time.time()
        img, st = X.render_scene(sc, cams, res, integration=integ, precision=prec, ds=ds, flat_field=ff,
                                 density_multiplier=dm, return_stats=True)
        dt = time.time() - t0
        err = np.abs(img.astype(np.float64) - ref)
        out[prec] = err.max()
        print(f"{name:28s} {integ[:4]} {prec} res={res} maxerr={err.max():.3e} n>1e-4={int((err > 1e-4).sum())} "
              f"ref_samples gpu/oracle={st['ref_samples']}/{nref} eval={st['evaluated_samples']} fb={st['fp64_fallbacks']} "
              f"prim={st['primitive_tests']} t={dt*1e3:.1f}ms")
    return out


def check_volume(shape=(32, 32, 32), res=32, rough=True):
    nx, ny, nz = shape
    rng = np.random.default_rng(1234)
    if rough:
        vol = rng.random((nz, nx, ny), dtype=np.float32)
    else:
        vol = np.zeros((nz, nx, ny), dtype=np.float32)
        vol[nz // 4:3 * nz // 4, nx // 4:3 * nx // 4, ny // 3:2 * ny // 3] = 1.0
    ds = 2.0 / min(shape) / 5.0
    views = ((0.0, 90.0), (37.0, 60.0))
    cams = X.cameras_from_angles(views, 4.0, 40.0)
    cams32 = X.to_legacy(cams)
    camsw = X.from_legacy(cams32)
    dsf = float(np.float32(ds))
    osc = O.OracleScene({"type": "voxel_grid", "_array": vol.astype(np.float64)})
    ref = np.zeros((len(views), res, res))
    refw = np.zeros((len(views), res, res))
    for v, (az, pol) in enumerate(views):
        eye, cm = O.camera_from_angles(az, pol, 4.0)
        ref[v], _ = osc.render_view(eye, cm, res, 40.0, 4.0, ds, "simple")
        e32 = np.array(list(camsw[v].eye)); m32 = np.array(list(camsw[v].view)).reshape(4, 4)
        refw[v], _ = osc.render_view(e32, m32, res, float(camsw[v].fov_y), float(camsw[v].R), dsf, "simple")
    for prec in ("fp64", "fp32"):
        img, st = X.render_volume(vol, cams, res, precision=prec, ds=ds, return_stats=True)
        err = np.abs(img.astype(np.float64) - ref)
        print(f"volume{shape} Ex {prec} maxerr={err.max():.3e} eval={st['evaluated_samples']} fb={st['fp64_fallbacks']}")
    img = X.render_volume_legacy(vol, cams32, res, dsf)
    err = np.abs(img.astype(np.float64) - refw)
    print(f"volume{shape} legacy maxerr={err.max():.3e} (vs oracle fed the fp32-rounded cameras/ds)")
    import os
    os.environ["XRAY_VOLUME_GENERIC"] = "1"
    img = X.render_volume_legacy(vol, cams32, res, dsf)
    del os.environ["XRAY_VOLUME_GENERIC"]
    err = np.abs(img.astype(np.float64) - refw)
    print(f"volume{shape} legacy(generic interpreter) maxerr={err.max():.3e}")


def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    check("cube_w_hole", "cube_w_hole.json", res=res)
    check("cube_w_hole simple", "cube_w_hole.json", res=res, integ="simple")
    check("balls", "balls.json", res=res)
    check("box_w_pped", "box_w_pped.json", res=res)
    check("pillar_array", "pillar_array.json", res=res)
    check("lattice", "lattice.json", res=res)
    check("gyroid+sigmoid", "gyroid_example.json", "deformation_sigmoid.json", res=min(res, 32), ds=0.004)
    check_volume((32, 32, 32), 32, True)
    check_volume((24, 32, 16), 32, False)
    check("cube+linear ff dm", "cube_w_hole.json", "deformation_linear.json", res=res, ff=0.1, dm=1.7)


if __name__ == "__main__":
    main()
