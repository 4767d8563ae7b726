"""Malformed-input fuzz of the C-ABI scene compiler (XRaySceneCompileJSON): truncations, byte flips, token insertions, deletions, deep
nesting and type-confused but well-formed JSON derived from the bundled scenes.  Run as a subprocess by
tests/test_host_frontend.py so that a crash shows up as a failed test, not a dead test run.  usage: fuzz_compile.py <seed> <n>"""
import ctypes, json, random, sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
from pathlib import Path
import xray_projection_render_b200 as X
L = X._lib.load()
L.XRaySceneCompileJSON.restype = ctypes.c_int
L.XRaySceneCompileJSON.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
L.XRaySceneFree.argtypes = [ctypes.c_void_p]
seed = int(sys.argv[1]); n = int(sys.argv[2])
rng = random.Random(seed)
scenes = [p.read_text() for p in sorted((Path(__file__).resolve().parents[1] / 'scenes').glob('*.json'))]
objs = [s for s in scenes if '"type"' in s]
tokens = ['{', '}', '[', ']', ',', ':', '"', 'null', 'true', '1e999', '-1e-999', 'NaN', '-', '0', '1', '"type"', '"objects"', '"uc"', '\\', '\\u00', ' ', '\n', '9'*400, '"'+'a'*3000+'"', '[[[[[[[[', ']]]]]]]]']
ok = err = 0
for it in range(n):
    s = rng.choice(objs)
    mode = rng.randrange(6)
    b = bytearray(s.encode())
    if mode == 0 and len(b) > 2:
        b = b[:rng.randrange(1, len(b))]
    elif mode == 1:
        for _ in range(rng.randrange(1, 6)):
            p = rng.randrange(len(b)); b[p] = rng.randrange(1, 256)
    elif mode == 2:
        for _ in range(rng.randrange(1, 4)):
            p = rng.randrange(len(b)); t = rng.choice(tokens).encode(); b[p:p] = t
    elif mode == 3:
        for _ in range(rng.randrange(1, 4)):
            p = rng.randrange(len(b)); q = min(len(b), p + rng.randrange(1, 40)); del b[p:q]
    elif mode == 4:  # deep nesting
        d = rng.randrange(10, 3000)
        b = bytearray(('{"type":"object_collection","objects":[' * d + ']}' * d).encode())
    else:  # structurally valid but semantically odd
        try:
            o = json.loads(s)
            def mut(x, depth=0):
                if isinstance(x, dict):
                    for k in list(x):
                        r = rng.random()
                        if r < 0.08: del x[k]
                        elif r < 0.16: x[k] = rng.choice([None, "x", [], {}, 1e308, -1e308, 0, -0.0, [1, 2], [1, 2, 3, 4], True])
                        else: mut(x[k], depth + 1)
                elif isinstance(x, list):
                    for i in range(len(x)):
                        if rng.random() < 0.1: x[i] = rng.choice([None, "x", [], {}, 1e308, 0])
                        else: mut(x[i], depth + 1)
            mut(o)
            b = bytearray(json.dumps(o).encode())
        except Exception:
            pass
    data = bytes(b).replace(b'\x00', b' ')
    d = None
    if rng.random() < 0.3:
        dd = rng.choice([s for s in scenes if 'deformation' in s or 'strains' in s or 'amplitude' in s] or scenes)
        db = bytearray(dd.encode())
        if rng.random() < 0.7 and len(db) > 2:
            p = rng.randrange(len(db)); db[p] = rng.randrange(1, 256)
        d = bytes(db).replace(b'\x00', b' ')
    h = ctypes.c_void_p()
    rc = L.XRaySceneCompileJSON(data, d, ctypes.byref(h))
    if rc == 0:
        ok += 1
        L.XRaySceneFree(h)
    else:
        err += 1
print(f"seed {seed}: {n} inputs, {ok} compiled, {err} rejected")
