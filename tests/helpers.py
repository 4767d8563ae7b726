"""Shared helpers for the parity tests."""
import numpy as np

R, FOV = 4.0, 40.0
TOL_FP32 = 1e-4   # north star: per-pixel absolute transmission error, fp32 mode
TOL_FP64 = 1e-9   # north star: fp64 mode


def oracle_images(O, osc, views, res, ds, integ, R=R, fov=FOV):
    imgs, n = [], 0
    for az, pol in views:
        eye, cm = O.camera_from_angles(az, pol, R)
        im, k = osc.render_view(eye, cm, res, fov, R, ds, integ)
        imgs.append(im)
        n += k
    return np.stack(imgs), n


def gpu_vs_oracle(X, O, obj, deform=None, *, views=((90.0, 90.0), (131.0, 70.0)), res=40, integ="hierarchical", ds=-1.0,
                  ff=0.0, dm=1.0, precisions=("fp32", "fp64")):
    """Render with the GPU library and with the oracle; return {precision: (max|dI|, stats)} and the oracle sample count."""
    sc = X.Scene(obj, deform)
    osc = O.OracleScene(obj, deform, flat_field=ff, density_multiplier=dm)
    if ds <= 0:
        ds = sc.auto_ds()
        assert ds == osc.auto_ds()
    ref, nref = oracle_images(O, osc, views, res, ds, integ)
    cams = X.cameras_from_angles(views, R, FOV)
    out = {}
    for prec in precisions:
        img, st = X.render_scene(sc, cams, res, integration=integ, precision=prec, ds=ds, flat_field=ff,
                                 density_multiplier=dm, return_stats=True)
        out[prec] = (float(np.abs(img.astype(np.float64) - ref).max()), st, img)
    return out, nref, ref


def assert_parity(out, nref):
    if "fp32" in out:
        err, st, _ = out["fp32"]
        assert err <= TOL_FP32, f"fp32 mode max|dI| {err:.3e} > {TOL_FP32}"
        assert st["ref_samples"] == nref, "reference-equivalent sample count differs from the oracle's density() calls"
    if "fp64" in out:
        err, st, _ = out["fp64"]
        assert err <= TOL_FP64, f"fp64 mode max|dI| {err:.3e} > {TOL_FP64}"
        assert st["ref_samples"] == nref
