"""Development aid: soak the randomised interval-renderer tests over many more seeds than the suite runs.
python tools/span_soak.py FIRST LAST"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import xray_projection_render_b200 as X  # noqa: E402
from oracle import oracle as O  # noqa: E402
import test_gpu_span as T  # noqa: E402


class MP:
    def setenv(self, k, v):
        os.environ[k] = v


first, last = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(first, last):
    for fn in (T.test_random_convex_scene_near_special_views, T.test_random_convex_scene_special_views):
        os.environ.pop("XRAY_SPAN_NO_BINS", None)
        try:
            if fn is T.test_random_convex_scene_near_special_views:
                fn(X, O, seed, MP())
            else:
                fn(X, O, seed)
        except AssertionError as e:
            bad += 1
            print("FAIL", fn.__name__, seed, str(e)[:200], flush=True)
print("done", first, last, "failures", bad)
