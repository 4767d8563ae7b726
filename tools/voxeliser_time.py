"""Development aid: AssembleVoxelGridCUDA (brute-force symbol) of this library against the reference plugin's, Kelvin foam.
python tools/voxeliser_time.py [cells] [res]"""
import ctypes
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import xray_projection_render_b200 as X  # noqa: E402
from test_gpu_reference_pin import kelvin_cylinders  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4
res = int(sys.argv[2]) if len(sys.argv) > 2 else 256
L = X._lib.load()
ref = ctypes.CDLL(str(ROOT / "oracle" / "_ref" / "libcuda_render_ref.so"), mode=os.RTLD_LAZY | os.RTLD_LOCAL)
arr = kelvin_cylinders(X, cells)
fp = ctypes.POINTER(ctypes.c_float)
out = {k: np.zeros((res, res, res), dtype=np.float32) for k in ("ours", "reference")}
for name, lib in (("ours", L), ("reference", ref), ("ours", L), ("reference", ref)):
    fn = lib.AssembleVoxelGridCUDA
    fn.restype = ctypes.c_int
    t0 = time.perf_counter()
    rc = fn(arr, len(arr), res, ctypes.c_float(1.0), out[name].ctypes.data_as(fp))
    print(f"VOXELISE {name:9s} rc={rc} {len(arr)} struts res {res}: {(time.perf_counter() - t0) * 1e3:8.1f} ms (whole call, host buffers)", flush=True)
print("identical:", bool(np.array_equal(out["ours"].view(np.uint32), out["reference"].view(np.uint32))))
