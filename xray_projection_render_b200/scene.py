"""Scene / deformation file front end.

Mirrors the reference loaders: ``load_object`` / ``load_deformation`` (main.go:50-120) and the
``FromMap`` type rules of objects/objects.go and deformations/deformations.go, then hands the
validated map to the native compiler (``XRaySceneCompileJSON``) which flattens it into the
device instruction buffer.
"""
from __future__ import annotations

import ctypes
import json
import os
import re

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------------------
# File loading (main.go:63,97: the format is sniffed from the LAST FOUR characters)
# ----------------------------------------------------------------------------------------
def _yaml_loader():
    import yaml

    class GoLikeLoader(yaml.SafeLoader):
        """yaml.v3 resolves 1e-3 and 1. as floats (YAML 1.2 core schema); PyYAML (1.1) does not."""

    GoLikeLoader.add_implicit_resolver(
        "tag:yaml.org,2002:float",
        re.compile(r"^[-+]?(\.[0-9]+|[0-9]+(\.[0-9]*)?)([eE][-+]?[0-9]+)?$|^[-+]?\.(inf|Inf|INF)$|^\.(nan|NaN|NAN)$"),
        list("-+0123456789."))
    return yaml, GoLikeLoader


def load_map(path: str) -> tuple[dict, bool]:
    """Returns (map, strict) where strict means YAML typing rules apply (ints are not floats)."""
    with open(path, "r") as fh:
        text = fh.read()
    ext = path[-4:]
    if ext == "yaml":
        yaml, loader = _yaml_loader()
        # ints written like `1` stay int; the implicit float resolver above only adds forms
        # that PyYAML would otherwise read as strings
        data = yaml.load(text, Loader=loader)
        return data, True
    if ext == "json":
        return json.loads(text), False
    raise ValueError(f"Unknown file extension: {ext}")


# ----------------------------------------------------------------------------------------
# FromMap validation.  strict=True reproduces Go's `.(float64)` assertions on YAML input.
# ----------------------------------------------------------------------------------------
class SceneError(ValueError):
    pass


def _is_num(v) -> bool:
    return isinstance(v, (int, float)) and not isinstance(v, bool)


def _f64_strict(d: dict, key: str, strict: bool, msg: str) -> float:
    """Go: `d[key].(float64)` -- an int from YAML fails; JSON numbers are always float64."""
    v = d.get(key)
    if not _is_num(v) or (strict and isinstance(v, int)):
        raise SceneError(msg)
    return float(v)


def _to_f64(v, msg: str) -> float:
    """objects.go:261-270 ToFloat64: int or float64."""
    if not _is_num(v):
        raise SceneError(msg)
    return float(v)


def _to_vec(d: dict, key: str, msg: str) -> list[float]:
    """objects.go:272-282 ToVec: ints and floats are taken, anything else is silently skipped."""
    v = d.get(key)
    if not isinstance(v, list):
        raise SceneError(msg)
    if len(v) > 3:
        raise SceneError(f"{key} has more than 3 elements")  # Go would panic on vec[3]
    out = [0.0, 0.0, 0.0]
    for i, val in enumerate(v):
        if _is_num(val):
            out[i] = float(val)
    return out


def _vec_f64(d: dict, key: str, msg: str, elem_msg: str, strict_elems: bool, strict: bool) -> list[float]:
    v = d.get(key)
    if not isinstance(v, list):
        raise SceneError(msg)
    if len(v) > 3:
        raise SceneError(f"{key} has more than 3 elements")
    out = [0.0, 0.0, 0.0]
    for i, val in enumerate(v):
        if not _is_num(val) or (strict_elems and strict and isinstance(val, int)):
            raise SceneError(elem_msg % i)
        out[i] = float(val)
    return out


def _bounds(d: dict) -> dict:
    return {k: _to_f64(d.get(k), f"{k} is not a float64") for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")}


_COLLECTION_CHILDREN = {"sphere", "cube", "box", "cylinder", "parallelepiped", "gyroid", "tessellated_obj_coll",
                        "voxel_grid"}


def normalize_object(d: dict, strict: bool, base_dir: str = ".", _voxels: list | None = None, _top: bool = True) -> dict:
    """Validate like objects.go NewObject/FromMap and return a canonical, all-float map."""
    if not isinstance(d, dict):
        raise SceneError("object description is not a map")
    t = d.get("type")
    if t == "sphere":  # objects.go:43-61
        return {"type": t, "center": _vec_f64(d, "center", "center is not a Vec3", "center[%d] is not a float64", False, strict),
                "radius": _f64_strict(d, "radius", strict, "radius is not a float64"),
                "rho": _f64_strict(d, "rho", strict, "rho is not a float64")}
    if t == "cube":  # objects.go:100-117: center elements are asserted float64 too
        return {"type": t, "center": _vec_f64(d, "center", "center is not a Vec3", "center[%d] is not a float64", True, strict),
                "side": _f64_strict(d, "side", strict, "side is not a float64"),
                "rho": _f64_strict(d, "rho", strict, "rho is not a float64")}
    if t == "box":  # objects.go:146-169
        return {"type": t, "center": _to_vec(d, "center", "center is not a Vec3"),
                "sides": _to_vec(d, "sides", "sides is not a Vec3"), "rho": _to_f64(d.get("rho"), "rho is not a float64")}
    if t == "parallelepiped":  # objects.go:207-245
        return {"type": t, "origin": _to_vec(d, "origin", "origin is not a Vec3"), "v0": _to_vec(d, "v0", "v0 is not a Vec3"),
                "v1": _to_vec(d, "v1", "v1 is not a Vec3"), "v2": _to_vec(d, "v2", "v2 is not a Vec3"),
                "rho": _to_f64(d.get("rho"), "rho is not a float64")}
    if t == "cylinder":  # objects.go:304-332
        out = {"type": t, "p0": _to_vec(d, "p0", "p0 is not a Vec3"), "p1": _to_vec(d, "p1", "p0 is not a Vec3"),
               "radius": _f64_strict(d, "radius", strict, "radius is not a float64")}
        out["rho"] = 1.0 if "rho" not in d else _to_f64(d["rho"], "rho is not a float64")
        return out
    if t == "gyroid":  # objects.go:979-1012
        return {"type": t, "center": _vec_f64(d, "center", "center is not a Vec3", "center[%d] is not a float64", False, strict),
                "scale": _f64_strict(d, "scale", strict, "scale is not a float64"),
                "thickness": _f64_strict(d, "thickness", strict, "thickness is not a float64"),
                "rho": _f64_strict(d, "rho", strict, "rho is not a float64")}
    if t == "object_collection":
        if not _top:
            raise SceneError("unknown object type")  # objects.go:391-410
        return _normalize_collection(d, strict, base_dir, _voxels)
    if t == "tessellated_obj_coll":  # objects.go:531-566, 478-514
        uc = d.get("uc")
        if not isinstance(uc, dict):
            raise SceneError("uc is not a map")
        objs = uc.get("objects")
        if not isinstance(objs, dict):
            raise SceneError("objects is not a map")
        ucn = {"type": "unit_cell", "objects": _normalize_collection(objs, strict, base_dir, _voxels)}
        ucn.update(_bounds(uc))
        out = {"type": t, "uc": ucn}
        out.update(_bounds(d))
        return out
    if t == "voxel_grid":  # objects.go:715-761
        path = d.get("path")
        if isinstance(path, str) and "resolution" in d:
            if path[path.rfind(".") + 1:].lower() != "raw":
                raise SceneError("only raw files are supported")
            res = d.get("resolution")
            if not isinstance(res, list):
                raise SceneError("resolution must be provided for raw files as a list of 3 integers")
            if len(res) != 3:
                raise SceneError("resolution must be a list of 3 integers")
            for i, r in enumerate(res):
                if not _is_num(r):
                    raise SceneError(f"resolution[{i}] has unsupported type {type(r).__name__} (need int-like)")
            dims = [int(r) for r in res]
            dtype = d.get("dtype") if isinstance(d.get("dtype"), str) else "uint8"
            arr = voxel_grid_from_raw(os.path.join(base_dir, path), dims, dtype)
        elif "_array" in d:  # in-memory volume (numpy [z][x][y]); not part of the file schema
            arr = np.ascontiguousarray(d["_array"])
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
            dims = [arr.shape[1], arr.shape[2], arr.shape[0]]
        else:
            raise SceneError("voxel_grid needs a raw file path with a resolution")
        if _voxels is not None:
            _voxels.append(arr)
        return {"type": t, "resolution": [float(x) for x in dims], "path": path if isinstance(path, str) else ""}
    raise SceneError(f"unknown object type `{t}`" if _top else "unknown object type")


def reference_object_map(n: dict) -> dict:
    """The object as the reference's ``ToMap()`` methods marshal it into object.json (main.go:538-546): per type exactly the
    keys of objects.go:33-40, 91-98, 139-146, 198-207, 296-304, 370-379 (no greedy flag), 466-477, 523-534, 704-713
    (nx / ny / nz, dtype always "float64", the raw file's path), 973-981."""
    t = n["type"]
    if t == "object_collection":
        return {"type": t, "objects": [reference_object_map(o) for o in n["objects"]]}
    if t == "unit_cell":
        out = {"type": t, "objects": reference_object_map(n["objects"])}
        out.update({k: n[k] for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")})
        return out
    if t == "tessellated_obj_coll":
        out = {"type": t, "uc": reference_object_map(n["uc"])}
        out.update({k: n[k] for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")})
        return out
    if t == "voxel_grid":
        nx, ny, nz = (int(v) for v in n["resolution"])
        return {"type": t, "nx": nx, "ny": ny, "nz": nz, "dtype": "float64", "path": n.get("path", "")}
    keys = {"sphere": ("center", "radius", "rho"), "cube": ("center", "side", "rho"), "box": ("center", "sides", "rho"),
            "parallelepiped": ("origin", "v0", "v1", "v2", "rho"), "cylinder": ("p0", "p1", "radius", "rho"),
            "gyroid": ("center", "scale", "thickness", "rho")}[t]
    out = {"type": t}
    out.update({k: n[k] for k in keys})
    return out


def _normalize_collection(d: dict, strict: bool, base_dir: str, voxels) -> dict:
    objs = d.get("objects")
    if not isinstance(objs, list):
        raise SceneError("objects is not a list")
    out = {"type": "object_collection", "objects": []}
    g = d.get("greedy_dens_eval")
    if isinstance(g, bool):
        out["greedy_dens_eval"] = g
    for o in objs:
        if not isinstance(o, dict) or o.get("type") not in _COLLECTION_CHILDREN:
            raise SceneError("unknown object type")
        out["objects"].append(normalize_object(o, strict, base_dir, voxels, _top=False))
    return out


def voxel_grid_from_raw(path: str, resolution, dtype: str) -> np.ndarray:
    """objects.go:892-958 VoxelGridFromRaw -> array [z][x][y]; fp32 stays fp32, the rest becomes fp64."""
    nx, ny, nz = (int(r) for r in resolution)
    table = {"uint8": ("<u1", 255.0), "uint16": ("<u2", 65535.0), "uint32": ("<u4", 4294967295.0),
             "float32": ("<f4", None), "float64": ("<f8", None)}
    if dtype not in table:
        raise SceneError(f"unsupported data type: {dtype}")
    np_dt, scale = table[dtype]
    try:
        raw = np.fromfile(path, dtype=np.uint8)
    except OSError as exc:
        raise SceneError(f"error reading file: {exc}") from exc
    expected = nx * ny * nz * np.dtype(np_dt).itemsize
    if raw.size != expected:
        raise SceneError(f"file size ({raw.size}) does not match expected size ({expected}) for type {dtype}")
    vals = raw.view(np_dt)
    if dtype == "float32":
        rho = vals.astype(np.float32)
    else:
        rho = vals.astype(np.float64)
        if scale is not None:
            rho = rho / scale
    return rho.reshape(nz, nx, ny)


def normalize_deformation(d: dict, strict: bool) -> dict:
    """deformations.go NewDeformation/FromMap type rules."""
    if not isinstance(d, dict) or d.get("type") is None:
        raise SceneError("deformation type is nil")
    t = d["type"]

    def flist(key, n, msg):
        v = d.get(key)
        if not isinstance(v, list):
            raise SceneError(msg)
        if len(v) < n:
            raise SceneError(f"{key} needs {n} elements")
        out = []
        for a in v[:max(n, len(v))]:
            if not _is_num(a) or (strict and isinstance(a, int)):
                raise SceneError(f"{key} elements must be float64")  # Go panics on a.(float64)
            out.append(float(a))
        return out

    if t == "gaussian":
        return {"type": t, "amplitudes": flist("amplitudes", 3, "amplitudes must be a list"),
                "sigmas": flist("sigmas", 3, "sigmas must be a list"), "centers": flist("centers", 3, "centers must be a list")}
    if t == "linear":
        return {"type": t, "strains": flist("strains", 6, "strains must be a list")}
    if t == "rigid":
        return {"type": t, "displacements": flist("displacements", 3, "displacements must be a list")}
    if t == "sigmoid":
        if not isinstance(d.get("direction"), str):
            raise SceneError("direction must be a string")
        if d["direction"] not in ("x", "y", "z"):
            raise SceneError("Invalid direction")
        return {"type": t, "amplitude": _to_f64(d.get("amplitude"), "amplitude must be a float"),
                "center": _to_f64(d.get("center"), "center must be a float"),
                "lengthscale": _to_f64(d.get("lengthscale"), "lengthscale must be a float"), "direction": d["direction"]}
    if t == "affine":
        m = d.get("matrix")
        if not isinstance(m, list):
            raise SceneError("matrix must be a list")
        if len(m) != 3:
            raise SceneError("matrix must have 3 rows")
        rows = []
        for row in m:
            if not isinstance(row, list):
                raise SceneError("matrix row must be a list")
            if len(row) != 3:
                raise SceneError("matrix row must have 3 elements")
            for a in row:
                if not _is_num(a) or (strict and isinstance(a, int)):
                    raise SceneError("matrix elements must be float64")
            rows.append([float(a) for a in row])
        return {"type": t, "matrix": rows}
    if t == "composed":
        subs = d.get("deformations")
        if not isinstance(subs, list):
            raise SceneError("deformations must be a list")
        return {"type": t, "deformations": [normalize_deformation(s, strict) for s in subs]}
    raise SceneError(f"unknown deformation type {t}")


# ----------------------------------------------------------------------------------------
# Compiled scene handle
# ----------------------------------------------------------------------------------------
class Scene:
    """A compiled scene: the reference's globals ``lat[0]`` and ``df[0]`` flattened for the GPU."""

    def __init__(self, obj, deformation=None, base_dir: str = "."):
        L = _lib.load()
        strict = False
        if isinstance(obj, (str, os.PathLike)):
            path = os.fspath(obj)
            base_dir = os.path.dirname(os.path.abspath(path))
            obj, strict = load_map(path)
        dstrict = False
        if isinstance(deformation, (str, os.PathLike)):
            deformation, dstrict = load_map(os.fspath(deformation)) if os.fspath(deformation) else (None, False)
        self._voxels: list[np.ndarray] = []
        self.object_map = normalize_object(obj, strict, base_dir, self._voxels)
        self.deformation_map = normalize_deformation(deformation, dstrict) if deformation else None
        h = ctypes.c_void_p()
        oj = json.dumps(self.object_map).encode()
        dj = json.dumps(self.deformation_map).encode() if self.deformation_map else None
        _lib.check(L.XRaySceneCompileJSON(oj, dj, ctypes.byref(h)))
        self._h = h
        self._L = L
        for slot, arr in enumerate(self._voxels):
            nz, nx, ny = arr.shape
            dt = _lib.VOXEL_F32 if arr.dtype == np.float32 else _lib.VOXEL_F64
            _lib.check(L.XRaySceneSetVoxelData(h, slot, arr.ctypes.data_as(ctypes.c_void_p), nx, ny, nz, dt))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.XRaySceneFree(h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def min_feature_size(self) -> float:
        return self._L.XRaySceneMinFeatureSize(self._h)

    def auto_ds(self) -> float:
        """main.go:350-353"""
        return self.min_feature_size() / 5.0

    def bounds(self):
        lo = (ctypes.c_double * 3)()
        hi = (ctypes.c_double * 3)()
        self._L.XRaySceneBounds(self._h, lo, hi)
        return list(lo), list(hi)

    def program_bytes(self) -> bytes:
        n = ctypes.c_size_t()
        p = self._L.XRaySceneProgram(self._h, ctypes.byref(n))
        return ctypes.string_at(p, n.value)

    def density_host(self, x, y, z, density_multiplier: float = 1.0) -> float:
        """fp64 host evaluation of main.go density() (testing / inspection)."""
        return self._L.XRaySceneDensityHost(self._h, x, y, z, density_multiplier)
