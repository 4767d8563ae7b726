"""ctypes front end of the CPU oracle (oracle/xray_oracle.cpp).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
``--impl reference`` legs may import this module.  The product package never does.

Parity status: pinned against output of the reference's Go binary (the image stored in the reference's demo notebook,
reproduced bit for bit -- tests/test_reference_go_output.py); details in the header of xray_oracle.cpp.

The oracle receives scenes as a token stream (hex floats, bit exact) that this module
derives from the same ``map[string]interface{}`` shaped dict the reference's ``FromMap``
methods consume (objects/objects.go, deformations/deformations.go).
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libxray_oracle.so"
_lib = None

CUBE_HALF_DIAGONAL = 1.74  # main.go:46


def build(force: bool = False) -> Path:
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    src = _HERE / "xray_oracle.cpp"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(_HERE), str(_LIB_PATH)])
    return _LIB_PATH


def build_ref() -> Path | None:
    subprocess.check_call(["make", "-s", "-C", str(_HERE), "ref"])
    p = _HERE / "_ref" / "libcuda_render_ref.so"
    return p if p.exists() else None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        d, vp, i, l = ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_long
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_last_error.restype = ctypes.c_char_p
        L.oracle_scene_create.restype = vp
        L.oracle_scene_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(dp), i]
        L.oracle_scene_destroy.argtypes = [vp]
        L.oracle_scene_set_globals.argtypes = [vp, d, d]
        L.oracle_min_feature_size.restype = d
        L.oracle_min_feature_size.argtypes = [vp]
        L.oracle_object_density.restype = d
        L.oracle_object_density.argtypes = [vp, d, d, d]
        L.oracle_density.restype = d
        L.oracle_density.argtypes = [vp, d, d, d]
        L.oracle_deform.argtypes = [vp, dp]
        L.oracle_integrate.restype = d
        L.oracle_integrate.argtypes = [vp, i, dp, dp, d, d, d, ctypes.POINTER(l)]
        L.oracle_camera_from_angles.argtypes = [d, d, d, dp, dp]
        L.oracle_render_view.restype = l
        L.oracle_render_view.argtypes = [vp, dp, dp, i, d, d, d, i, i, i, i, dp, i]
        L.oracle_render_pixels.restype = l
        L.oracle_render_pixels.argtypes = [vp, dp, dp, i, d, d, d, i, ctypes.POINTER(i), i, dp, i]
        L.oracle_step_count.restype = l
        L.oracle_step_count.argtypes = [i, d, d, d]
        L.oracle_mat3_inv.argtypes = [dp, dp]
        L.oracle_max_threads.restype = i
        _lib = L
    return _lib


# ----------------------------------------------------------------------------------------
# dict -> token stream, following the reference's FromMap rules
# ----------------------------------------------------------------------------------------
def _h(x) -> str:
    return float(x).hex()


def _vec(v) -> str:
    assert len(v) == 3
    return " ".join(_h(c) for c in v)


def load_map(path: str) -> dict:
    """main.go:63,97 -- the type is sniffed from the last four characters."""
    ext = path[-4:]
    with open(path, "r") as fh:
        text = fh.read()
    if ext == "yaml":
        import yaml

        return yaml.safe_load(text)
    if ext == "json":
        return json.loads(text)
    raise ValueError(f"unknown file extension {ext!r}")


class _Ctx:
    def __init__(self):
        self.vox = []


def _obj_tokens(d: dict, ctx: _Ctx, base_dir: str) -> str:
    t = d.get("type")
    if t == "sphere":
        return f"sphere {_vec(d['center'])} {_h(d['radius'])} {_h(d['rho'])}"
    if t == "cube":
        return f"cube {_vec(d['center'])} {_h(d['side'])} {_h(d['rho'])}"
    if t == "box":
        return f"box {_vec(d['center'])} {_vec(d['sides'])} {_h(d['rho'])}"
    if t == "parallelepiped":
        return f"pped {_vec(d['origin'])} {_vec(d['v0'])} {_vec(d['v1'])} {_vec(d['v2'])} {_h(d['rho'])}"
    if t == "cylinder":
        rho = d.get("rho", 1.0)  # objects.go:326-330
        return f"cylinder {_vec(d['p0'])} {_vec(d['p1'])} {_h(d['radius'])} {_h(rho)}"
    if t == "gyroid":
        return f"gyroid {_vec(d['center'])} {_h(d['scale'])} {_h(d['thickness'])} {_h(d['rho'])}"
    if t == "object_collection":
        return _coll_tokens(d, ctx, base_dir, force_greedy=False)
    if t == "tessellated_obj_coll":
        uc = d["uc"]
        b = " ".join(_h(d[k]) for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax"))
        ub = " ".join(_h(uc[k]) for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax"))
        # uc.objects is parsed as an ObjectCollection whatever its "type" says (objects.go:481-487)
        return f"tess {b} {ub} {_coll_tokens(uc['objects'], ctx, base_dir, force_greedy=True)}"
    if t == "voxel_grid":
        if "_array_f32" in d:  # test hook for full-size volumes: fp32 array borrowed without the 2x fp64 copy
            arr = d["_array_f32"]
            assert arr.dtype == np.float32 and arr.flags["C_CONTIGUOUS"]
            nz, nx, ny = arr.shape
            ctx.vox.append(arr)
            return f"voxelf {nx} {ny} {nz} {len(ctx.vox) - 1}"
        if "_array" in d:  # test hook: in-memory volume, layout [z][x][y]
            arr = np.ascontiguousarray(d["_array"], dtype=np.float64)
            nz, nx, ny = arr.shape
        else:
            arr, (nx, ny, nz) = voxel_grid_from_raw(os.path.join(base_dir, d["path"]), d["resolution"], d.get("dtype", "uint8"))
        ctx.vox.append(arr)
        return f"voxel {nx} {ny} {nz} {len(ctx.vox) - 1}"
    raise ValueError(f"unknown object type {t!r}")


def _coll_tokens(d: dict, ctx: _Ctx, base_dir: str, force_greedy: bool) -> str:
    greedy = bool(d.get("greedy_dens_eval", False)) or force_greedy
    objs = d["objects"]
    for o in objs:
        if o.get("type") == "object_collection":
            raise ValueError("unknown object type")  # objects.go:391-410 rejects nested collections
    body = " ".join(_obj_tokens(o, ctx, base_dir) for o in objs)
    return f"collection {int(greedy)} {len(objs)} {body}"


def voxel_grid_from_raw(path: str, resolution, dtype: str):
    """objects.go:892-958: little-endian raw -> float64, layout [z][x][y]."""
    nx, ny, nz = (int(r) for r in resolution)
    np_dt = {"uint8": "<u1", "uint16": "<u2", "uint32": "<u4", "float32": "<f4", "float64": "<f8"}[dtype]
    raw = np.fromfile(path, dtype=np_dt)
    if raw.size != nx * ny * nz:
        raise ValueError("file size does not match expected size")
    scale = {"uint8": 255.0, "uint16": 65535.0, "uint32": 4294967295.0}.get(dtype)
    rho = raw.astype(np.float64)
    if scale is not None:
        rho = rho / scale
    return rho.reshape(nz, nx, ny), (nx, ny, nz)


def _deform_tokens(d: dict) -> str:
    t = d["type"]
    if t == "gaussian":
        return "gaussian " + " ".join(_h(x) for k in ("amplitudes", "sigmas", "centers") for x in d[k])
    if t == "affine":
        return "affine " + " ".join(_h(x) for row in d["matrix"] for x in row)
    if t == "linear":
        return "linear " + " ".join(_h(x) for x in d["strains"])
    if t == "rigid":
        return "rigid " + " ".join(_h(x) for x in d["displacements"])
    if t == "sigmoid":
        axis = {"x": 0, "y": 1, "z": 2}[d["direction"]]
        return f"sigmoid {_h(d['amplitude'])} {_h(d['center'])} {_h(d['lengthscale'])} {axis}"
    if t == "composed":
        subs = d["deformations"]
        return f"composed {len(subs)} " + " ".join(_deform_tokens(s) for s in subs)
    raise ValueError(f"unknown deformation type {t!r}")


class OracleScene:
    """The reference's globals (lat, df, flat_field, density_multiplier) for one render."""

    def __init__(self, obj: dict | str, deformation: dict | str | None = None, flat_field: float = 0.0,
                 density_multiplier: float = 1.0, base_dir: str = "."):
        L = lib()
        if isinstance(obj, str):
            base_dir = os.path.dirname(os.path.abspath(obj))
            obj = load_map(obj)
        if isinstance(deformation, str):
            deformation = load_map(deformation) if deformation else None
        ctx = _Ctx()
        desc = _obj_tokens(obj, ctx, base_dir).encode()
        ddesc = _deform_tokens(deformation).encode() if deformation else b""
        dp = ctypes.POINTER(ctypes.c_double)
        self._vox = ctx.vox
        arr = (dp * max(1, len(ctx.vox)))(*[v.ctypes.data_as(dp) for v in ctx.vox])
        self._h = L.oracle_scene_create(desc, ddesc, arr, len(ctx.vox))
        if not self._h:
            raise ValueError("oracle: " + L.oracle_last_error().decode())
        self._L = L
        self.set_globals(flat_field, density_multiplier)

    def set_globals(self, flat_field: float, density_multiplier: float):
        self._L.oracle_scene_set_globals(self._h, flat_field, density_multiplier)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.oracle_scene_destroy(self._h)
            self._h = None

    def min_feature_size(self) -> float:
        return self._L.oracle_min_feature_size(self._h)

    def auto_ds(self) -> float:
        """main.go:350-353"""
        return self.min_feature_size() / 5.0

    def object_density(self, x, y, z) -> float:
        return self._L.oracle_object_density(self._h, x, y, z)

    def density(self, x, y, z) -> float:
        return self._L.oracle_density(self._h, x, y, z)

    def deform(self, x, y, z):
        a = (ctypes.c_double * 3)(x, y, z)
        self._L.oracle_deform(self._h, a)
        return a[0], a[1], a[2]

    def integrate(self, integrator: str, origin, direction, ds, smin, smax):
        o = (ctypes.c_double * 3)(*origin)
        d = (ctypes.c_double * 3)(*direction)
        n = ctypes.c_long(0)
        v = self._L.oracle_integrate(self._h, _integ(integrator), o, d, ds, smin, smax, ctypes.byref(n))
        return v, n.value

    def render_view(self, eye, cam_rowmajor, res, fov_deg, R, ds, integrator="hierarchical", rows=None, jstride=1,
                    nthreads=0):
        """One view, out[i, j] (GPU ABI layout).  Returns (image float64 [res,res], n_samples)."""
        dp = ctypes.POINTER(ctypes.c_double)
        eye = np.ascontiguousarray(eye, dtype=np.float64)
        cam = np.ascontiguousarray(cam_rowmajor, dtype=np.float64).reshape(16)
        out = np.zeros((res, res), dtype=np.float64)
        i0, i1 = (0, res) if rows is None else rows
        n = self._L.oracle_render_view(self._h, eye.ctypes.data_as(dp), cam.ctypes.data_as(dp), res, fov_deg, R, ds,
                                       _integ(integrator), i0, i1, jstride, out.ctypes.data_as(dp), nthreads)
        return out, n

    def render_pixels(self, eye, cam_rowmajor, res, fov_deg, R, ds, ij, integrator="hierarchical", nthreads=0):
        dp = ctypes.POINTER(ctypes.c_double)
        eye = np.ascontiguousarray(eye, dtype=np.float64)
        cam = np.ascontiguousarray(cam_rowmajor, dtype=np.float64).reshape(16)
        ij = np.ascontiguousarray(ij, dtype=np.int32).reshape(-1, 2)
        out = np.zeros(len(ij), dtype=np.float64)
        n = self._L.oracle_render_pixels(self._h, eye.ctypes.data_as(dp), cam.ctypes.data_as(dp), res, fov_deg, R, ds,
                                         _integ(integrator), ij.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(ij),
                                         out.ctypes.data_as(dp), nthreads)
        return out, n


def _integ(name) -> int:
    if name in (0, "simple"):
        return 0
    if name in (1, "hierarchical"):
        return 1
    raise ValueError(name)


def camera_from_angles(az_deg: float, polar_deg: float, R: float):
    """main.go:226-239 -> (eye[3], camera row-major [4,4])."""
    dp = ctypes.POINTER(ctypes.c_double)
    eye = np.zeros(3)
    cam = np.zeros(16)
    lib().oracle_camera_from_angles(az_deg, polar_deg, R, eye.ctypes.data_as(dp), cam.ctypes.data_as(dp))
    return eye, cam.reshape(4, 4)


def step_count(integrator, ds, smin, smax) -> int:
    return lib().oracle_step_count(_integ(integrator), ds, smin, smax)


def generate_camera_angles(num_images: int, job_num: int = 0, jobs_modulo: int = 1, polar_angle: float = 90.0):
    """main.go:242-257 (out_of_plane=False): th = i*(360/N) + 90."""
    out = []
    for i_img in range(job_num, num_images, jobs_modulo):
        dth = 360.0 / float(num_images)
        out.append((float(i_img) * dth + 90.0, polar_angle))
    return out


def max_threads() -> int:
    return lib().oracle_max_threads()
