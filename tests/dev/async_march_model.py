"""CPU model of the lane-asynchronous march for a tessellated gyroid (render_fast.cu render_async_kernel).

Two uses (test infrastructure, like the oracle it checks against; needs only numpy):
  * `python tests/dev/async_march_model.py check`  -- runs the kernel's per-lane state machine (fp32 evaluation, guard band,
    second-order skip bound, immediate fine replay, fp64 oracle for guard-band samples) on a small image and compares
    transmissions and reference-equivalent sample counts with the oracle.  This is how the skip bound and the state
    machine were validated before any GPU time was spent.
  * `python tests/dev/async_march_model.py iters`  -- warp iteration counts of the lockstep march (every lane at the same
    lattice index, skip = warp minimum) against the asynchronous one (max over lanes), with the first-order Lipschitz
    rule and with the second-order rule.  BASELINE config 3 (gyroid + sigmoid, ds = 4e-4): 340 -> 69.
"""
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

f32 = np.float32
R, FOV = 4.0, 40.0


def lattice(ds, smin, smax):
    s = [smin]
    right = smin + ds
    while right <= smax:
        s.append(right)
        right += ds
    s = np.array(s)
    dsf = ds / 10.0
    nfine = np.zeros(len(s), dtype=np.int64)
    for k in range(len(s) - 1):
        left, n = s[k] + dsf, 0
        while left < s[k + 1]:
            n += 1
            left += dsf
        nfine[k] = n
    return s, nfine


class Scene:
    """gyroid_example.json-like: outer box +-ob, unit cell +-1 (never folds), one gyroid, optional sigmoid-z / linear warp."""

    def __init__(self, scale=0.1, thick=0.2, ob=0.8, rho=1.0, sigmoid=None, linear=None):
        self.scale, self.thick, self.ob, self.rho, self.sigmoid, self.linear = scale, thick, ob, rho, sigmoid, linear
        self.obj = {"type": "tessellated_obj_coll", "xmin": -ob, "xmax": ob, "ymin": -ob, "ymax": ob, "zmin": -ob, "zmax": ob,
                    "uc": {"xmin": -1.0, "xmax": 1.0, "ymin": -1.0, "ymax": 1.0, "zmin": -1.0, "zmax": 1.0,
                           "objects": {"objects": [{"type": "gyroid", "center": [0.0, 0.0, 0.0], "scale": scale, "thickness": thick, "rho": rho}]}}}
        self.deform = None
        lip, jac2, curv = 1.0, 1.0, 0.0
        if sigmoid:
            A, c, L = sigmoid
            self.deform = {"type": "sigmoid", "amplitude": A, "center": c, "lengthscale": L, "direction": "z"}
            lip = jac2 = 1.0 + abs(A / (4 * L))
            curv = abs(A) * 0.0962251 / (L * L)
        if linear:
            self.deform = {"type": "linear", "strains": list(linear)}
            e = linear
            M = np.array([[1 + e[0], e[5], e[4]], [e[5], 1 + e[1], e[3]], [e[4], e[3], 1 + e[2]]])
            self.M = M
            lip = np.abs(M).sum(axis=1).max()
            jac2 = math.sqrt((M * M).sum())
        self.lip = max(1.0, lip)
        self.M2 = 1.01 * (2 * (jac2 / scale) ** 2 + 3 * curv / scale)
        self.g1eps = 1e-3 * jac2
        ep = 1e-6 * lip + 3e-7
        amax = 3.0 / scale
        self.tol = 2.0 * (3.0 * (ep / scale + 3 * 6e-8 * amax) + 6.0 * (1e-6 + 4 * 6e-8 * (1 + 1e-2 * amax)) + 12 * 6e-8)
        self.osc = O.OracleScene(self.obj, self.deform)

    def warp(self, x, y, z, d):
        """fp32 warp + e = J d."""
        e = d.copy()
        if self.sigmoid:
            A, c, L = (f32(v) for v in self.sigmoid)
            with np.errstate(over="ignore"):
                E = np.exp((z - c) * f32(-1.0 / float(L)), dtype=f32)
            sg = f32(1) / (f32(1) + E)
            z = z + A / (f32(1) + E)
            e[..., 2] = d[..., 2] * (f32(1) + (A / L) * sg * (f32(1) - sg))
        if self.linear:
            M = self.M.astype(f32)
            p = np.stack([x, y, z], -1) @ M.T
            x, y, z = p[..., 0], p[..., 1], p[..., 2]
            e = d @ M.T
        return x, y, z, e


def rays(az, pol, res, ij):
    eye, cm = O.camera_from_angles(az, pol, R)
    F = 1.0 / math.tan(math.radians(FOV) / 2)
    i, j = ij[:, 0], ij[:, 1]
    v = np.stack([i / (res / 2) - 1, j / (res / 2) - 1, -F * np.ones(len(i)), np.ones(len(i))], -1) @ cm.T
    v = v[:, :3] / v[:, 3:4] - eye
    d = v * (1.0 / np.sqrt((v * v).sum(1)))[:, None]
    return eye, d


def march(sc: Scene, az, pol, res, ds, rule="second", lockstep_warps=None):
    """Per-lane asynchronous march of every pixel; returns (image, reference-equivalent samples, iterations per lane)."""
    smin, smax = R - 1.74, R + 1.74
    s_tab, nfine = lattice(ds, smin, smax)
    n_steps = len(s_tab) - 1
    t_tab = (s_tab - R).astype(f32)
    dsf = f32(ds / 10.0)
    wC, wF = f32(ds), dsf
    ij = np.stack(np.meshgrid(np.arange(res), np.arange(res), indexing="ij"), -1).reshape(-1, 2)
    eye, d64 = rays(az, pol, res, ij)
    pc = (eye + d64 * R).astype(f32)
    pd = d64.astype(f32)
    n = len(ij)
    m2s = f32(0.999 / (ds * sc.lip))
    k = np.zeros(n, dtype=np.int64)
    k1 = np.full(n, n_steps)
    jf = np.zeros(n, dtype=np.int64)
    nf = np.zeros(n, dtype=np.int64)
    kf = np.zeros(n, dtype=np.int64)
    T = np.zeros(n)
    prev = np.zeros(n, dtype=f32)
    n_fine = np.zeros(n, dtype=np.int64)
    iters = np.zeros(n, dtype=np.int64)
    fallbacks = 0
    while True:
        fine = jf < nf
        act = fine | (k < k1)
        if not act.any():
            break
        iters += act
        base = np.where(act, np.where(fine, kf, k + 1), 0)
        nsub = np.where(fine, jf + 1, 0)
        t = t_tab[base] + np.where(fine, nsub.astype(f32) * dsf, f32(0))
        t = t.astype(f32)
        x, y, z = (pd[:, a] * t + pc[:, a] for a in range(3))
        x, y, z, e = sc.warp(x.astype(f32), y.astype(f32), z.astype(f32), pd)
        m = np.maximum(np.abs(x) - f32(sc.ob), np.maximum(np.abs(y) - f32(sc.ob), np.abs(z) - f32(sc.ob)))
        inside = m <= 0
        edge = act & (np.abs(m) < f32(4e-6))
        q = [(v * f32(1.0 / sc.scale)).astype(f32) for v in (x, y, z)]
        sx, cx, sy, cy, sz, cz = np.sin(q[0]), np.cos(q[0]), np.sin(q[1]), np.cos(q[1]), np.sin(q[2]), np.cos(q[2])
        g = sx * cy + sy * cz + sz * cx
        tt = np.abs(g) - f32(sc.thick)
        test = act & inside
        near = (np.abs(tt) < f32(sc.tol)) & test
        rho = np.where(test & (tt < 0), f32(sc.rho), f32(0))
        mm = np.maximum(np.abs(tt) - f32(sc.tol), 0)
        if rule == "second":
            gx, gy, gz = cx * cy - sz * sx, cy * cz - sx * sy, cz * cx - sy * sz
            G1 = (np.abs(gx * e[:, 0] + gy * e[:, 1] + gz * e[:, 2]) + f32(sc.g1eps)) * f32(1.0 / sc.scale)
            delta = 2 * mm / (G1 + np.sqrt(2 * f32(sc.M2) * mm + G1 * G1))
            clr = delta * f32(sc.lip)
        else:
            clr = mm * f32(sc.scale / 3.03)
        clear = np.where(test, np.maximum(np.minimum(clr, -m) - f32(1e-5), 0) * m2s, 0)
        clear = np.where(act & ~inside, np.maximum(m - f32(1.6e-5), 0) * m2s, clear)
        unc = near | edge
        for idx in np.nonzero(unc)[0]:  # guard band: the reference decides
            s = s_tab[base[idx]]
            for _ in range(nsub[idx]):
                s += ds / 10.0
            p = eye + d64[idx] * s
            rho[idx] = sc.osc.density(*p)
            fallbacks += 1
        # bookkeeping: refined sub-steps skip with the same clearance (clear coarse steps = 10 * clear sub-steps)
        af = np.minimum(np.where(unc, 0, clear * f32(10.0)), (nf - jf).astype(f32)).astype(np.int64)
        advf = np.where(fine & (af >= 2), af, 1)
        T += np.where(fine, rho.astype(np.float64) * wF * advf, 0)
        n_fine += np.where(fine, advf, 0)
        jf = jf + np.where(fine, advf, 0)
        coarse = act & ~fine
        trans = coarse & ((rho == 0) != (prev == 0))
        kf = np.where(trans, k, kf)
        nf = np.where(trans, nfine[np.minimum(k, n_steps)], nf)
        jf = np.where(trans, 0, jf)
        w = np.where(trans, wF, wC)
        T += np.where(coarse, rho.astype(np.float64) * w, 0)
        prev = np.where(coarse, rho, prev)
        a = np.where(unc, 0, clear)
        a = np.minimum(a, (k1 - k).astype(f32))
        nskip = a.astype(np.int64)
        adv = np.where(coarse & (nskip >= 2), nskip, 1)
        T += np.where(coarse & (nskip >= 2) & (rho != 0), rho.astype(np.float64) * wC * (adv - 1), 0)
        k = k + np.where(coarse, adv, 0)
    img = np.exp(-T).reshape(res, res)
    return img, int(n * n_steps + n_fine.sum()), iters.reshape(res, res), fallbacks


def check():
    worst = 0.0
    for name, sc, ds in (("sigmoid", Scene(sigmoid=(0.2, 0.0, 0.2)), 0.002), ("none", Scene(), 0.002),
                         ("linear", Scene(linear=(0.02, -0.03, 0.01, 0.05, 0.04, -0.06)), 0.002),
                         ("steep sigmoid", Scene(scale=0.15, thick=0.3, sigmoid=(-0.3, 0.1, 0.05)), 0.003)):
        for az, pol in ((90.0, 90.0), (131.0, 70.0)):
            res = 24
            img, nref, iters, fb = march(sc, az, pol, res, ds)
            eye, cm = O.camera_from_angles(az, pol, R)
            ref, n = sc.osc.render_view(eye, cm, res, FOV, R, ds, "hierarchical")
            err = float(np.abs(img - ref).max())
            worst = max(worst, err)
            print(f"{name:14s} az={az:5.1f} pol={pol:4.1f}: max|dI| = {err:.2e}  samples {nref} vs oracle {n}  "
                  f"iterations/lane mean {iters.mean():.0f} max {iters.max()}  fp64 re-evaluations {fb}")
            assert nref == n and err <= 1e-4
    print("ok, worst", worst)


def iters():
    sc = Scene(sigmoid=(0.2, 0.0, 0.2))
    ds = 0.1 * sc.scale * sc.thick / 5
    for rule in ("first", "second"):
        _, _, it, _ = march(sc, 120.0, 90.0, 64, ds, rule=rule)
        # a warp is a 4 x 8 pixel tile: asynchronous cost = max over its lanes
        tiles = it.reshape(16, 4, 8, 8).transpose(0, 2, 1, 3).reshape(-1, 32)
        print(f"{rule}-order rule, asynchronous: iterations per warp = {tiles.max(1).mean():.0f} (mean lane {it.mean():.0f})")


if __name__ == "__main__":
    {"check": check, "iters": iters}[sys.argv[1] if len(sys.argv) > 1 else "check"]()
