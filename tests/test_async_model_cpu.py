"""CPU check of the ALGORITHM behind render_async_kernel (render_fast.cu): a numpy model of the per-lane state machine --
fp32 evaluation, guard band, first- and second-order gyroid skip bounds, refinement replay, reference decides guard-band
samples -- must reproduce the oracle's reference-equivalent sample counts exactly and its images far inside the fp32
tolerance.  The CUDA kernel itself is checked on the GPU (test_gpu_parity.py); this pins the mathematics of the skip
bounds (|h(t+tau) - h(t)| <= |h'| tau + M2 tau^2 / 2 with the warp's Jacobian and curvature) without one."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "dev"))

from helpers import FOV, R, TOL_FP32  # noqa: E402


@pytest.mark.parametrize("rule", ["first", "second"])
@pytest.mark.parametrize("warp", ["none", "sigmoid", "steep_sigmoid", "linear"])
def test_async_march_model_matches_oracle(O, warp, rule):
    import async_march_model as M

    sc = {"none": lambda: M.Scene(), "sigmoid": lambda: M.Scene(sigmoid=(0.2, 0.0, 0.2)),
          "steep_sigmoid": lambda: M.Scene(scale=0.15, thick=0.3, sigmoid=(-0.3, 0.1, 0.05)),
          "linear": lambda: M.Scene(linear=(0.02, -0.03, 0.01, 0.05, 0.04, -0.06))}[warp]()
    ds, res = 0.004, 16
    for az, pol in ((131.0, 70.0),):
        img, nref, iters, _ = M.march(sc, az, pol, res, ds, rule=rule)
        eye, cm = O.camera_from_angles(az, pol, R)
        ref, n = sc.osc.render_view(eye, cm, res, FOV, R, ds, "hierarchical")
        assert nref == n
        assert np.abs(img - ref).max() <= TOL_FP32 / 100
        assert ref.min() < 0.95  # the view crosses the gyroid
    if rule == "second":  # the second-order bound must not need more evaluations than the Lipschitz rule
        _, _, it1, _ = M.march(sc, 131.0, 70.0, res, ds, rule="first")
        assert iters.sum() <= it1.sum()
