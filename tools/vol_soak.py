"""Development aid: soak tests/test_gpu_volume.py::test_hierarchical_volume_kernel_random.  python tools/vol_soak.py FIRST LAST"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import xray_projection_render_b200 as X  # noqa: E402
from oracle import oracle as O  # noqa: E402
import test_gpu_volume as T  # noqa: E402

first, last = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(first, last):
    try:
        T.test_hierarchical_volume_kernel_random(X, O, seed)
    except AssertionError as e:
        bad += 1
        import traceback
        tb = traceback.extract_tb(e.__traceback__)[-1]
        print("FAIL", seed, tb.lineno, tb.line, str(e)[:200].replace("\n", " "), flush=True)
print("done", first, last, "failures", bad)
