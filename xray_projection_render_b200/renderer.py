"""Host-side mirror of the reference's render orchestration on top of libcuda_render.so.

* ``render_scene`` / ``render_volume`` / ``render_volume_legacy``: arrays in, arrays out
  (the hot path through the C ABI, host buffers).
* ``render_scene_device`` / ``render_volume_device``: device-resident buffers (torch tensors or
  raw device pointers), asynchronous on a CUDA stream.
* ``XRayRenderer.render(params, camera_angles)``: same signature, parameter names, defaults and
  result dict as the reference's Python binding (xray_projection_render/xray_renderer.py:357-448,
  api.go:41-172), producing the same files (PNG frames, transforms.json, object.json,
  main.go:380-546) -- with the per-pixel work done on the GPU.
"""
from __future__ import annotations

import concurrent.futures
import ctypes
import json
import math
import os
import struct
import zlib
from typing import Dict, List, Optional

import numpy as np

from . import _lib
from .camera import (CUBE_HALF_DIAGONAL, camera_matrix, cameras_from_angles, generate_camera_angles, to_legacy)
from .scene import reference_object_map, Scene, SceneError


def _as_cam_array(cams):
    if isinstance(cams, ctypes.Array):
        return cams
    arr = (_lib.XRayCameraParams64 * len(cams))()
    for k, c in enumerate(cams):
        arr[k] = c
    return arr


def _stats_dict(buf) -> dict:
    return {"ref_samples": int(buf[0]), "evaluated_samples": int(buf[1]), "fp64_fallbacks": int(buf[2]),
            "primitive_tests": int(buf[3]), "rays": int(buf[4]), "launches": int(buf[5]), "marched_tiles": int(buf[6]),
            "march_reasons": int(buf[7]) & 0xFFFF, "span_renderer": bool(int(buf[7]) & 0x10000),
            "clipping_smin": bool(int(buf[7]) & (1 << 17)), "clipping_smax": bool(int(buf[7]) & (1 << 18))}


def render_scene(scene: Scene, cams, res: int, *, integration="hierarchical", precision="fp32", ds: float = -1.0,
                 flat_field: float = 0.0, density_multiplier: float = 1.0, out: np.ndarray | None = None,
                 devices=None, return_stats: bool = False):
    """Render ``len(cams)`` views -> array [n, res, res] with out[cam, i, j] (cuda_backend.h:96-97 layout).

    fp32 precision returns float32 images, fp64 precision float64 images (unless ``out`` says otherwise).
    """
    L = _lib.load()
    cams = _as_cam_array(cams)
    n = len(cams)
    if out is None:
        out = np.empty((n, res, res), dtype=np.float64 if precision == "fp64" else np.float32)
    if out.shape != (n, res, res) or not out.flags.c_contiguous or out.dtype not in (np.float32, np.float64):
        raise ValueError("out must be a C-contiguous float32/float64 array of shape [n, res, res]")
    stats = (ctypes.c_uint64 * _lib.XRAY_NUM_STATS)()
    o = _lib.make_opts(integration, precision, "f64" if out.dtype == np.float64 else "f32", ds, flat_field,
                       density_multiplier, devices, 0, stats if return_stats else None)
    _lib.check(L.XRayRenderSceneCUDA(scene.handle, cams, n, res, ctypes.byref(o), out.ctypes.data_as(ctypes.c_void_p)))
    return (out, _stats_dict(stats)) if return_stats else out


def render_scene_device(scene: Scene, cams, res: int, out, *, integration="hierarchical", precision="fp32",
                        ds: float = -1.0, flat_field: float = 0.0, density_multiplier: float = 1.0, stream: int = 0,
                        stats=None):
    """Device-resident output.  ``out``: torch CUDA tensor [n,res,res] (float32/float64) or an int device pointer
    (then ``out_dtype`` follows ``precision``).  Enqueued on ``stream`` (a cudaStream_t handle), not synchronised."""
    L = _lib.load()
    cams = _as_cam_array(cams)
    if hasattr(out, "data_ptr"):
        ptr = out.data_ptr()
        od = "f64" if out.element_size() == 8 else "f32"
    else:
        ptr = int(out)
        od = "f64" if precision == "fp64" else "f32"
    o = _lib.make_opts(integration, precision, od, ds, flat_field, density_multiplier, None, stream, stats)
    _lib.check(L.XRayRenderSceneDeviceCUDA(scene.handle, cams, len(cams), res, ctypes.byref(o), ctypes.c_void_p(ptr)))
    return out


def render_volume(volume: np.ndarray, cams, res: int, *, integration="simple", precision="fp32", ds: float = -1.0,
                  flat_field: float = 0.0, density_multiplier: float = 1.0, out: np.ndarray | None = None, devices=None,
                  return_stats: bool = False):
    """Voxel volume [z][x][y] (float32 or float64) through the extended entry point."""
    L = _lib.load()
    cams = _as_cam_array(cams)
    n = len(cams)
    volume = np.ascontiguousarray(volume)
    if volume.dtype not in (np.float32, np.float64):
        volume = volume.astype(np.float32)
    nz, nx, ny = volume.shape
    if ds <= 0:
        ds = 2.0 / float(min(nx, ny, nz)) / 5.0  # cuda_path.go:42-53
    if out is None:
        out = np.empty((n, res, res), dtype=np.float64 if precision == "fp64" else np.float32)
    stats = (ctypes.c_uint64 * _lib.XRAY_NUM_STATS)()
    o = _lib.make_opts(integration, precision, "f64" if out.dtype == np.float64 else "f32", ds, flat_field,
                       density_multiplier, devices, 0, stats if return_stats else None)
    vdt = _lib.VOXEL_F32 if volume.dtype == np.float32 else _lib.VOXEL_F64
    _lib.check(L.XRayRenderVolumeExCUDA(volume.ctypes.data_as(ctypes.c_void_p), vdt, nx, ny, nz, cams, n, res,
                                        ctypes.byref(o), out.ctypes.data_as(ctypes.c_void_p)))
    return (out, _stats_dict(stats)) if return_stats else out


def render_volume_legacy(volume: np.ndarray, cams32, res: int, ds: float, flat_field: float = 0.0,
                         out: np.ndarray | None = None) -> np.ndarray:
    """The exact call the unmodified Go host makes (cuda_backend.go:321-332): RenderVolumeProjectionsCUDA."""
    L = _lib.load()
    volume = np.ascontiguousarray(volume, dtype=np.float32)
    nz, nx, ny = volume.shape
    n = len(cams32)
    if out is None:
        out = np.empty((n, res, res), dtype=np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    _lib.check(L.RenderVolumeProjectionsCUDA(volume.ctypes.data_as(fp), nx, ny, nz, cams32, n, res, ctypes.c_float(ds),
                                             ctypes.c_float(flat_field), out.ctypes.data_as(fp)))
    return out


def render_volume_device(d_volume, dims, cams, res: int, out, *, ds: float, integration="simple", precision="fp32",
                         flat_field: float = 0.0, density_multiplier: float = 1.0, stream: int = 0, stats=None):
    """Device-resident fp32 volume (torch CUDA tensor [z][x][y] or int pointer) and device images."""
    L = _lib.load()
    cams = _as_cam_array(cams)
    nx, ny, nz = dims
    vptr = d_volume.data_ptr() if hasattr(d_volume, "data_ptr") else int(d_volume)
    optr = out.data_ptr() if hasattr(out, "data_ptr") else int(out)
    od = "f64" if (hasattr(out, "element_size") and out.element_size() == 8) else "f32"
    o = _lib.make_opts(integration, precision, od, ds, flat_field, density_multiplier, None, stream, stats)
    _lib.check(L.XRayRenderVolumeDeviceCUDA(ctypes.c_void_p(vptr), nx, ny, nz, cams, len(cams), res, ctypes.byref(o),
                                            ctypes.c_void_p(optr)))
    return out


def render_volume_device_to_host(d_volume, dims, cams, res: int, out: np.ndarray, *, ds: float, integration="simple",
                                 precision="fp32", flat_field: float = 0.0, density_multiplier: float = 1.0, return_stats: bool = False):
    """Device-resident fp32 volume (torch CUDA tensor [z][x][y] or int pointer), images into the HOST array ``out``
    (float32 / float64, pageable or pinned); synchronous.  The caller orders the volume's producer before the call."""
    L = _lib.load()
    cams = _as_cam_array(cams)
    nx, ny, nz = dims
    vptr = d_volume.data_ptr() if hasattr(d_volume, "data_ptr") else int(d_volume)
    stats = (ctypes.c_uint64 * _lib.XRAY_NUM_STATS)()
    o = _lib.make_opts(integration, precision, "f64" if out.dtype == np.float64 else "f32", ds, flat_field, density_multiplier, None, 0,
                       stats if return_stats else None)
    _lib.check(L.XRayRenderVolumeDeviceToHostCUDA(ctypes.c_void_p(vptr), nx, ny, nz, cams, len(cams), res, ctypes.byref(o),
                                                  out.ctypes.data_as(ctypes.c_void_p)))
    return (out, _stats_dict(stats)) if return_stats else out


def voxelize_scene(scene: Scene, res: int, density_multiplier: float = 1.0) -> np.ndarray:
    """density() on the export grid of main.go:208-214 -> float32 [k][i][j]."""
    out = np.empty((res, res, res), dtype=np.float32)
    _lib.check(_lib.load().XRayVoxelizeSceneCUDA(scene.handle, res, density_multiplier,
                                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
    return out


def measure_fp32_peak() -> float:
    v = ctypes.c_double()
    _lib.check(_lib.load().XRayMeasureFp32Peak(ctypes.byref(v)))
    return v.value


# ----------------------------------------------------------------------------------------
# Output formats (main.go:482-546)
# ----------------------------------------------------------------------------------------
def image_to_rgba8(img: np.ndarray, transparency: bool = False) -> np.ndarray:
    """main.go:482-498: uint16(val*0xffff) stored through SetRGBA64 into an 8-bit RGBA image
    (top 8 bits), pixel (i, j) at (x=i, y=res-1-j).  img is [i, j]."""
    res = img.shape[0]
    v16 = (img.astype(np.float64) * 0xFFFF).astype(np.int64).clip(0, 0xFFFF).astype(np.uint16)
    g8 = (v16 >> 8).astype(np.uint8)
    rgba = np.empty((res, res, 4), dtype=np.uint8)  # [y, x, c]
    gray_yx = g8.T[::-1, :]  # y = res-1-j, x = i
    rgba[..., 0] = rgba[..., 1] = rgba[..., 2] = gray_yx
    if transparency:
        alpha = np.where(img.astype(np.float64) < 1.0, 255, 0).astype(np.uint8).T[::-1, :]
        rgba[..., 3] = alpha
        # image.RGBA is alpha-premultiplied and png.Encode writes it un-premultiplied: a pixel with alpha 0 comes out as
        # (0, 0, 0, 0) whatever colour SetRGBA64 stored (image/png writer.go, cbTCA8 case)
        rgba[alpha == 0, 0:3] = 0
    else:
        rgba[..., 3] = 255
    return rgba


def write_png(path: str, rgba: np.ndarray) -> None:
    """Go's png.Encode of an *image.RGBA: 8-bit RGB (colour type 2) when every pixel is opaque, else 8-bit RGBA (type 6)."""
    h, w, _ = rgba.shape
    opaque = bool((rgba[..., 3] == 255).all())
    rows = np.ascontiguousarray(rgba[..., :3]) if opaque else rgba
    raw = b"".join(b"\x00" + rows[y].tobytes() for y in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n")
        fh.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if opaque else 6, 0, 0, 0)))
        fh.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        fh.write(chunk(b"IEND", b""))


def _go_sprintf_int(pattern: str, i: int) -> str:
    return pattern % i  # fname_pattern uses C/Go style verbs such as image_%03d.png


class XRayRenderer:
    """Drop-in for the reference's ``XRayRenderer`` (xray_renderer.py:63): same ``render`` contract,
    backed by the B200 plugin instead of the Go shared library."""

    RENDER_BATCH_BYTES = 1 << 30  # images rendered per library call by render()

    def __init__(self, library_path: Optional[str] = None, precision: str = "fp32", devices=None):
        if library_path:
            os.environ["XRAY_CUDA_LIB"] = library_path
        self.lib = _lib.load()
        self.precision = precision
        self.devices = devices

    def render(self, params: Dict, camera_angles: Optional[List[Dict[str, float]]] = None) -> Dict:
        p = {
            "input": params.get("input"),
            "output_dir": params.get("output_dir", "images"),
            "fname_pattern": params.get("fname_pattern", "image_%03d.png"),
            "resolution": params.get("resolution", 512),
            "num_images": params.get("num_images", 1),
            "out_of_plane": params.get("out_of_plane", False),
            "ds": params.get("ds", -1.0),
            "R": params.get("R", 4.0),
            "fov": params.get("fov", 40.0),
            "jobs_modulo": params.get("jobs_modulo", 1),
            "job_num": params.get("job_num", 0),
            "transforms_file": params.get("transforms_file", "transforms.json"),
            "deformation_file": params.get("deformation_file", ""),
            "time_label": params.get("time_label", 0.0),
            "transparency": params.get("transparency", False),
            "export_volume": params.get("export_volume", False),
            "polar_angle": params.get("polar_angle", 90.0),
            "density_multiplier": params.get("density_multiplier", 1.0),
            "flat_field": params.get("flat_field", 0.0),
            "integration": params.get("integration", "hierarchical"),
            "camera_angles": [],
        }
        if camera_angles is not None:
            p["camera_angles"] = camera_angles
        elif "camera_angles" in params:
            p["camera_angles"] = params["camera_angles"]
        if not p["input"]:
            raise ValueError("'input' parameter is required")
        # api.go:96-103
        if p["ds"] == 0:
            return {"success": False, "error": "ds is 0; use a negative value for automatic step-size selection",
                    "num_images": 0, "output_dir": ""}
        if p["density_multiplier"] == 0:
            return {"success": False, "error": "density_multiplier is 0; all densities will be zero and the render "
                                               "will produce a blank image", "num_images": 0, "output_dir": ""}
        try:
            n = self._render(p)
        except (SceneError, _lib.XRayError, OSError, ValueError) as exc:
            return {"success": False, "error": f"render failed: {exc}", "num_images": 0, "output_dir": ""}
        return {"success": True, "num_images": p["num_images"], "output_dir": p["output_dir"], "rendered": n}

    # main.go:310-546 render()
    def _render(self, p: dict) -> int:
        scene = Scene(p["input"], p["deformation_file"] or None)
        os.makedirs(p["output_dir"], exist_ok=True)
        res = int(p["resolution"])
        ds = float(p["ds"])
        if ds < 0:
            ds = scene.auto_ds()
        angles = list(p["camera_angles"])
        norm = []
        for a in angles:  # api.go CameraAngle JSON keys are case-insensitive in Go
            lower = {k.lower(): v for k, v in a.items()}
            norm.append({"azimuthal": float(lower["azimuthal"]), "polar": float(lower["polar"])})
        if not norm:
            norm = generate_camera_angles(int(p["num_images"]), int(p["job_num"]), int(p["jobs_modulo"]),
                                          bool(p["out_of_plane"]), float(p["polar_angle"]))
        R, fov = float(p["R"]), float(p["fov"])
        cams = cameras_from_angles(norm, R, fov)
        integration = "simple" if p["integration"] == "simple" else "hierarchical"  # api.go:128-132
        f = 1 / math.tan((fov / 2) * math.pi / 180.0)
        transform_params = {
            "flat_field": math.exp(-float(p["flat_field"])),
            "camera_angle_x": fov * math.pi / 180.0,
            "fl_x": f * float(res) / 2.0,
            "fl_y": f * float(res) / 2.0,
            "w": res,
            "h": res,
            "cx": float(res) / 2.0,
            "cy": float(res) / 2.0,
            "frames": [],
        }
        # The reference renders and saves one view at a time (main.go:430-522).  Here the views go through the library in
        # batches of at most RENDER_BATCH_BYTES of images (2880 views at 4096^2 would otherwise be 193 GB of host memory),
        # and a batch's frames are quantised and PNG-encoded by a few threads (zlib releases the GIL) while the next batch
        # renders; two batches are alive at most.  File names, frame order and bytes do not depend on the batching.
        n_views = len(norm)
        per_view = res * res * (8 if self.precision == "fp64" else 4)
        batch = max(1, min(n_views, self.RENDER_BATCH_BYTES // max(1, per_view)))
        transparency = bool(p["transparency"])
        filenames = [os.path.join(p["output_dir"], _go_sprintf_int(p["fname_pattern"], i_img)) for i_img in range(n_views)]

        def save(filename, img):
            write_png(filename, image_to_rgba8(img, transparency))

        workers = max(1, min(8, os.cpu_count() or 1, n_views))
        with concurrent.futures.ThreadPoolExecutor(max_workers=workers) as pool:
            in_flight = []  # futures of the previous batch
            for v0 in range(0, n_views, batch):
                sub = [cams[k] for k in range(v0, min(n_views, v0 + batch))]
                imgs = render_scene(scene, sub, res, integration=integration, precision=self.precision, ds=ds,
                                    flat_field=float(p["flat_field"]), density_multiplier=float(p["density_multiplier"]),
                                    devices=self.devices)
                for fut in in_flight:
                    fut.result()  # re-raises a failed write; bounds the images alive to two batches
                in_flight = [pool.submit(save, filenames[v0 + k], imgs[k]) for k in range(len(sub))]
            for fut in in_flight:
                fut.result()
        for i_img, filename in enumerate(filenames):
            dname, fname = os.path.split(filename)
            rel = os.path.join(os.path.basename(dname), fname)
            transform_params["frames"].append({"file_path": rel.replace(os.sep, "/"), "time": float(p["time_label"]),
                                               "transform_matrix": camera_matrix(cams[i_img]).tolist()})
        with open(p["transforms_file"], "w") as fh:
            json.dump(transform_params, fh, indent=2)
        obj_path = os.path.join(os.path.dirname(p["output_dir"]), "object.json")
        with open(obj_path, "w") as fh:
            json.dump(reference_object_map(scene.object_map), fh, indent=2)
        if p["export_volume"]:  # main.go:549-635 (volume.raw, uint8, [z][x][y] with x = i/res*2-1)
            vol = voxelize_scene(scene, res, float(p["density_multiplier"])).astype(np.float64)
            mx = vol.max()
            if mx == 0:
                mx = 1.0
            (vol / mx * 255).astype(np.uint8).tofile(os.path.join(os.path.dirname(p["output_dir"]), "volume.raw"))
        return len(norm)
