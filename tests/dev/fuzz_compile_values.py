"""Edge-value fuzz of the C-ABI scene compiler: the bundled scenes and deformations with numeric leaves replaced by zeros, denormals,
huge and negative values (degenerate unit cells, zero radii, 1e308 extents).  The compiler must return (0 or an error code) promptly:
no crash, no runaway grid construction.  Run as a subprocess by tests/test_host_frontend.py.  usage: fuzz_compile_values.py <seed> <n>"""
import ctypes, json, random, sys, time
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
from pathlib import Path
import xray_projection_render_b200 as X
L = X._lib.load()
L.XRaySceneCompileJSON.restype = ctypes.c_int
L.XRaySceneCompileJSON.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
L.XRaySceneFree.argtypes = [ctypes.c_void_p]
seed = int(sys.argv[1]); n = int(sys.argv[2])
rng = random.Random(seed)
files = sorted((Path(__file__).resolve().parents[1] / 'scenes').glob('*.json'))
objs = [json.loads(p.read_text()) for p in files if 'deformation' not in p.name]
defs = [json.loads(p.read_text()) for p in files if 'deformation' in p.name]
edge = [0.0, -0.0, 1e-300, 1e300, -1e300, 1.7e308, 5e-324, 1e-12, 1e12, 3.0, -1.0, 1e-6, 0.5, 1e38, 1e-38, 2.0**53]
def mut(x, p):
    if isinstance(x, dict):
        return {k: mut(v, p) for k, v in x.items()}
    if isinstance(x, list):
        return [mut(v, p) for v in x]
    if isinstance(x, (int, float)) and not isinstance(x, bool) and rng.random() < p:
        return rng.choice(edge) if rng.random() < 0.7 else x * rng.choice([1e-9, 1e9, -1, 1 + 1e-15, 0.999999])
    return x
ok = err = 0; slow = 0
for it in range(n):
    o = mut(rng.choice(objs), rng.choice([0.02, 0.1, 0.5]))
    d = mut(rng.choice(defs), 0.3) if rng.random() < 0.4 else None
    t0 = time.time()
    h = ctypes.c_void_p()
    rc = L.XRaySceneCompileJSON(json.dumps(o).encode(), json.dumps(d).encode() if d else None, ctypes.byref(h))
    dt = time.time() - t0
    if dt > 1.0:
        slow += 1; print('SLOW', dt, json.dumps(o)[:300])
    if rc == 0:
        ok += 1; L.XRaySceneFree(h)
    else:
        err += 1
print(f"seed {seed}: {n} inputs, {ok} compiled, {err} rejected, {slow} slow")
