"""Development aid: soak tests/test_gpu_fuzz.py over many more seeds than the suite runs, both kernel paths.
python tools/fuzz_soak.py FIRST LAST"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import xray_projection_render_b200 as X  # noqa: E402
from oracle import oracle as O  # noqa: E402
import test_gpu_fuzz as T  # noqa: E402

first, last = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(first, last):
    for fn in (T.test_random_scene, T.test_random_one_primitive_scene):
        for no_span in (False, True):
            if no_span:
                os.environ["XRAY_NO_SPAN"] = "1"
            else:
                os.environ.pop("XRAY_NO_SPAN", None)
            try:
                fn(X, O, seed)
            except AssertionError as e:
                bad += 1
                print("FAIL", fn.__name__, seed, "no_span" if no_span else "span", str(e)[:200], flush=True)
print("done", first, last, "failures", bad)
