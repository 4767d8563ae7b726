"""Parity at BASELINE.json's FULL sizes ("missing" item 4): index-space coordinates reach 1023 in fp32 in the
voxel kernel, so the 1e-4 bound has to be shown where it is hardest, not only on 96x128x64 grids.

  * config 4: synthetic 1024^3 fp32 volume (bench.synthetic_volume), 2048^2 detector, legacy symbol, simple integrator,
    ds = 2/1024/5 -- three of the 1440 views, > 2000 pixels each: random ones plus rays grazing the r = 0.9 mask (a jump of
    up to 1.0 in density) and rays that clip the cube's edges (in/out-of-volume test at |x| = 1);
  * a 512^3 uniform-random field (worst case roughness: every voxel differs from its neighbours);
  * config 1: the full 512^2 image of cube_w_hole, hierarchical, both precisions.

The oracle borrows the fp32 volume (OracleScene `_array_f32`), so the 1024^3 case needs 4 GiB of host memory, not 12.
"""
import os

import numpy as np
import pytest

from helpers import FOV, R, TOL_FP32, TOL_FP64, assert_parity

pytestmark = pytest.mark.gpu


def _ray_dirs(cam, res):
    """World-space unit directions of every pixel of a view (main.go:457-465), numpy fp64; for PICKING pixels only."""
    eye = np.array(list(cam.eye))
    M = np.array(list(cam.view)).reshape(4, 4)
    f = 1.0 / np.tan(np.deg2rad(float(cam.fov_y)) / 2.0)
    a = np.arange(res) / (res / 2.0) - 1.0
    px, py = np.meshgrid(a, a, indexing="ij")
    v = np.stack([px, py, np.full_like(px, -f), np.ones_like(px)], axis=-1) @ M.T
    w = v[..., :3] / v[..., 3:4] - eye
    return eye, w / np.linalg.norm(w, axis=-1, keepdims=True)


def _pick_pixels(cam, res, rng, n_random, n_mask, n_edge):
    eye, d = _ray_dirs(cam, res)
    # closest approach to the origin: |eye - (eye.d) d|
    t0 = -(d @ eye)
    b = np.linalg.norm(eye + d * t0[..., None], axis=-1)
    graze = np.argsort(np.abs(b - 0.9), axis=None)[: 4 * n_mask]
    graze = rng.choice(graze, size=n_mask, replace=False)
    # chord inside [-1,1]^3 (slab test); shortest non-empty chords clip an edge or a corner of the cube
    with np.errstate(divide="ignore", invalid="ignore"):
        t1, t2 = (-1.0 - eye) / d, (1.0 - eye) / d
    tn, tf = np.minimum(t1, t2).max(-1), np.maximum(t1, t2).min(-1)
    chord = np.where(tf > tn, tf - tn, np.inf)
    edge = np.argsort(chord, axis=None)[: 4 * n_edge]
    edge = rng.choice(edge, size=n_edge, replace=False)
    rnd = rng.integers(0, res * res, size=n_random)
    flat = np.unique(np.concatenate([graze, edge, rnd]))
    return np.stack([flat // res, flat % res], axis=1).astype(np.int32)


def _check_volume_views(X, O, vol, views, res, ds, n_random, n_mask, n_edge, seed):
    nz, nx, ny = vol.shape
    cams = X.cameras_from_angles(views, R, FOV)
    cams32 = X.to_legacy(cams)
    ds32 = float(np.float32(ds))
    img = X.render_volume_legacy(vol, cams32, res, ds32)
    osc = O.OracleScene({"type": "voxel_grid", "_array_f32": vol})
    rng = np.random.default_rng(seed)
    worst = 0.0
    for v, c in enumerate(X.from_legacy(cams32)):
        ij = _pick_pixels(c, res, rng, n_random, n_mask, n_edge)
        want, _ = osc.render_pixels(np.array(list(c.eye)), np.array(list(c.view)).reshape(4, 4), res, float(c.fov_y), float(c.R),
                                    ds32, ij, "simple")
        got = img[v][ij[:, 0], ij[:, 1]].astype(np.float64)
        err = np.abs(got - want)
        worst = max(worst, float(err.max()))
        assert err.max() <= TOL_FP32, f"view {views[v]}: max|dI| {err.max():.3e} at pixel {ij[err.argmax()]}"
        assert want.min() < 0.9 and want.max() > 0.999  # both attenuated and free rays were checked
    return worst


@pytest.mark.timeout(1500)
def test_config4_full_size_1024_cubed_2048_detector(X, O):
    import bench

    if os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") < 20 * 2 ** 30:
        pytest.skip("needs ~10 GiB of host memory")
    n, res, total_views = 1024, 2048, 1440
    vol = bench.synthetic_volume(n)
    views = [(v * 360.0 / total_views + 90.0, 90.0) for v in (0, 487, 1201)]  # main.go:242-257
    worst = _check_volume_views(X, O, vol, views, res, 2.0 / n / 5.0, n_random=1700, n_mask=300, n_edge=200, seed=4)
    print(f"config 4 full size: max|dI| = {worst:.3e} over 3 views")


@pytest.mark.timeout(900)
def test_rough_512_cubed_field(X, O):
    rng = np.random.default_rng(1234)
    vol = rng.random((512, 512, 512), dtype=np.float32)
    worst = _check_volume_views(X, O, vol, [(17.0, 90.0), (203.0, 64.0)], 1024, 2.0 / 512 / 5.0, n_random=1500, n_mask=0, n_edge=300,
                                seed=5)
    print(f"512^3 rough field: max|dI| = {worst:.3e}")


@pytest.mark.timeout(900)
def test_config1_full_512_image(X, O, scenes):
    """BASELINE configs[0]: examples/cube_w_hole.yaml, one 512x512 projection, default (hierarchical) integrator, auto ds."""
    from helpers import gpu_vs_oracle

    out, nref, ref = gpu_vs_oracle(X, O, str(scenes / "cube_w_hole.json"), views=[(90.0, 90.0)], res=512)
    assert_parity(out, nref)
    assert ref.min() < 0.6 and ref.max() == 1.0
    print(f"config 1 full image: fp32 {out['fp32'][0]:.2e}, fp64 {out['fp64'][0]:.2e}, {nref} reference samples")
