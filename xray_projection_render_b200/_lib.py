"""ctypes binding of libcuda_render.so (include/xray_cuda_render.h).

The library is searched like the reference's Go loader does (cuda_backend.go:103-114):
``$XRAY_CUDA_LIB`` first, then next to this package (``lib/libcuda_render.so``), then the
bare name through the dynamic loader.  There is no CPU fallback: if the library cannot be
loaded every call raises.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_NAME = "libcuda_render.so"

XRAY_MAX_DEVICES = 16
XRAY_NUM_STATS = 8
INTEGRATE_SIMPLE, INTEGRATE_HIERARCHICAL = 0, 1
PRECISION_FP32, PRECISION_FP64 = 0, 1
OUT_F32, OUT_F64 = 0, 1
VOXEL_F32, VOXEL_F64 = 0, 1

LEGACY_SYMBOLS = ("AssembleVoxelGridCUDA", "AssembleVoxelGridSpatialCUDA", "RenderVolumeProjectionsCUDA")
EXTENDED_SYMBOLS = (
    "XRayLastError", "XRayDeviceCount", "XRayReleaseCaches", "XRayRenderOptsInit", "XRaySceneCompileJSON", "XRaySceneFree",
    "XRaySceneMinFeatureSize", "XRaySceneProgram", "XRaySceneBounds", "XRaySceneNumVoxelSlots", "XRaySceneVoxelDims",
    "XRaySceneSetVoxelData", "XRaySceneDensityHost", "XRayCameraFromAngles", "XRayRenderSceneCUDA",
    "XRayRenderSceneDeviceCUDA", "XRayRenderVolumeExCUDA", "XRayRenderVolumeDeviceCUDA", "XRayRenderVolumeDeviceToHostCUDA",
    "XRayVoxelizeSceneCUDA",
    "XRayMeasureFp32Peak", "XRayBuildInfo",
)


class CylinderParams(ctypes.Structure):  # cuda_backend.h:22-27
    _fields_ = [("p0", ctypes.c_float * 3), ("p1", ctypes.c_float * 3), ("radius", ctypes.c_float),
                ("rho", ctypes.c_float)]


class XRayCameraParams(ctypes.Structure):  # cuda_backend.h:78-83
    _fields_ = [("eye", ctypes.c_float * 3), ("view", ctypes.c_float * 16), ("fov_y", ctypes.c_float),
                ("R", ctypes.c_float)]


class XRayCameraParams64(ctypes.Structure):
    _fields_ = [("eye", ctypes.c_double * 3), ("view", ctypes.c_double * 16), ("fov_y", ctypes.c_double),
                ("R", ctypes.c_double)]


class XRayRenderOpts(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("integration", ctypes.c_int32),
        ("precision", ctypes.c_int32),
        ("out_dtype", ctypes.c_int32),
        ("ds", ctypes.c_double),
        ("flat_field", ctypes.c_double),
        ("density_multiplier", ctypes.c_double),
        ("num_devices", ctypes.c_int32),
        ("devices", ctypes.c_int32 * XRAY_MAX_DEVICES),
        ("stream", ctypes.c_uint64),
        ("stats", ctypes.POINTER(ctypes.c_uint64)),
        ("view_begin", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 7),
    ]


class XRayError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libcuda_render error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def find_library() -> str:
    p = os.environ.get("XRAY_CUDA_LIB")
    if p:
        return p
    cand = _PKG / "lib" / LIB_NAME
    if cand.exists():
        return str(cand)
    return LIB_NAME


def library_info() -> dict:
    """Which binary this process loaded: path, size, modification time, SHA-256 and the build text compiled into it."""
    import hashlib
    import time

    L = load()
    path = find_library()
    st = os.stat(path)
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for chunk in iter(lambda: fh.read(1 << 20), b""):
            h.update(chunk)
    return {"path": str(path), "bytes": st.st_size, "mtime_utc": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime(st.st_mtime)),
            "sha256": h.hexdigest(), "build": L.XRayBuildInfo().decode()}


def load() -> ctypes.CDLL:
    """dlopen the plugin and declare every prototype; raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = find_library()
    try:
        L = ctypes.CDLL(path, mode=os.RTLD_LAZY | os.RTLD_LOCAL)
    except OSError as exc:  # pragma: no cover - exercised only on broken installs
        raise RuntimeError(
            f"CUDA library not available ({path}): {exc}. Build it with "
            f"`make -C {_PKG / 'csrc'}` or `python -c 'import __graft_entry__ as g; g.build()'`.") from exc
    c_int, c_float, c_double, c_void_p, c_char_p = (ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_void_p,
                                                    ctypes.c_char_p)
    fp = ctypes.POINTER(c_float)
    ip = ctypes.POINTER(c_int)
    dp = ctypes.POINTER(c_double)
    # legacy surface
    L.AssembleVoxelGridCUDA.restype = c_int
    L.AssembleVoxelGridCUDA.argtypes = [ctypes.POINTER(CylinderParams), c_int, c_int, c_float, fp]
    L.AssembleVoxelGridSpatialCUDA.restype = c_int
    L.AssembleVoxelGridSpatialCUDA.argtypes = [ctypes.POINTER(CylinderParams), c_int, c_int, c_float, c_int, ip, ip,
                                               c_int, fp]
    L.RenderVolumeProjectionsCUDA.restype = c_int
    L.RenderVolumeProjectionsCUDA.argtypes = [fp, c_int, c_int, c_int, ctypes.POINTER(XRayCameraParams), c_int, c_int,
                                              c_float, c_float, fp]
    # extended surface
    L.XRayLastError.restype = c_char_p
    L.XRayDeviceCount.restype = c_int
    L.XRayBuildInfo.restype = c_char_p
    L.XRayRenderOptsInit.argtypes = [ctypes.POINTER(XRayRenderOpts)]
    L.XRaySceneCompileJSON.restype = c_int
    L.XRaySceneCompileJSON.argtypes = [c_char_p, c_char_p, ctypes.POINTER(c_void_p)]
    L.XRaySceneFree.argtypes = [c_void_p]
    L.XRaySceneMinFeatureSize.restype = c_double
    L.XRaySceneMinFeatureSize.argtypes = [c_void_p]
    L.XRaySceneProgram.restype = c_void_p
    L.XRaySceneProgram.argtypes = [c_void_p, ctypes.POINTER(ctypes.c_size_t)]
    L.XRaySceneBounds.argtypes = [c_void_p, dp, dp]
    L.XRaySceneNumVoxelSlots.restype = c_int
    L.XRaySceneNumVoxelSlots.argtypes = [c_void_p]
    L.XRaySceneVoxelDims.restype = c_int
    L.XRaySceneVoxelDims.argtypes = [c_void_p, c_int, ip, ip, ip]
    L.XRaySceneSetVoxelData.restype = c_int
    L.XRaySceneSetVoxelData.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int]
    L.XRaySceneDensityHost.restype = c_double
    L.XRaySceneDensityHost.argtypes = [c_void_p, c_double, c_double, c_double, c_double]
    L.XRayCameraFromAngles.restype = c_int
    L.XRayCameraFromAngles.argtypes = [c_double, c_double, c_double, c_double, ctypes.POINTER(XRayCameraParams64)]
    cam64p = ctypes.POINTER(XRayCameraParams64)
    optp = ctypes.POINTER(XRayRenderOpts)
    L.XRayRenderSceneCUDA.restype = c_int
    L.XRayRenderSceneCUDA.argtypes = [c_void_p, cam64p, c_int, c_int, optp, c_void_p]
    L.XRayRenderSceneDeviceCUDA.restype = c_int
    L.XRayRenderSceneDeviceCUDA.argtypes = [c_void_p, cam64p, c_int, c_int, optp, c_void_p]
    L.XRayRenderVolumeExCUDA.restype = c_int
    L.XRayRenderVolumeExCUDA.argtypes = [c_void_p, c_int, c_int, c_int, c_int, cam64p, c_int, c_int, optp, c_void_p]
    L.XRayRenderVolumeDeviceCUDA.restype = c_int
    L.XRayRenderVolumeDeviceCUDA.argtypes = [c_void_p, c_int, c_int, c_int, cam64p, c_int, c_int, optp, c_void_p]
    L.XRayRenderVolumeDeviceToHostCUDA.restype = c_int
    L.XRayRenderVolumeDeviceToHostCUDA.argtypes = [c_void_p, c_int, c_int, c_int, cam64p, c_int, c_int, optp, c_void_p]
    L.XRayVoxelizeSceneCUDA.restype = c_int
    L.XRayVoxelizeSceneCUDA.argtypes = [c_void_p, c_int, c_double, fp]
    L.XRayMeasureFp32Peak.restype = c_int
    L.XRayMeasureFp32Peak.argtypes = [dp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = load().XRayLastError()
        raise XRayError(rc, msg.decode("utf-8", "replace") if msg else "")


def make_opts(integration="hierarchical", precision="fp32", out_dtype="f32", ds=-1.0, flat_field=0.0,
              density_multiplier=1.0, devices=None, stream=0, stats=None) -> XRayRenderOpts:
    o = XRayRenderOpts()
    load().XRayRenderOptsInit(ctypes.byref(o))
    o.integration = {"simple": INTEGRATE_SIMPLE, "hierarchical": INTEGRATE_HIERARCHICAL}[integration] \
        if isinstance(integration, str) else int(integration)
    o.precision = {"fp32": PRECISION_FP32, "fp64": PRECISION_FP64}[precision] if isinstance(precision, str) else int(precision)
    o.out_dtype = {"f32": OUT_F32, "f64": OUT_F64}[out_dtype] if isinstance(out_dtype, str) else int(out_dtype)
    o.ds = float(ds)
    o.flat_field = float(flat_field)
    o.density_multiplier = float(density_multiplier)
    if devices:
        o.num_devices = len(devices)
        for k, d in enumerate(devices):
            o.devices[k] = int(d)
    o.stream = int(stream)
    if stats is not None:
        o.stats = stats
    return o
