"""Camera helpers mirroring main.go:226-257 (computeCameraFromAngles, generateCameraAngles)."""
from __future__ import annotations

import ctypes
import math
import random

import numpy as np

from . import _lib

CUBE_HALF_DIAGONAL = 1.74  # main.go:46


def generate_camera_angles(num_images: int, job_num: int = 0, jobs_modulo: int = 1, out_of_plane: bool = False,
                           polar_angle: float = 90.0) -> list[dict]:
    """main.go:242-257: th = i*(360/N) + 90 for i = job, job+modulo, ...; polar fixed or random."""
    angles = []
    for i_img in range(job_num, num_images, jobs_modulo):
        dth = 360.0 / float(num_images)
        th = float(i_img) * dth + 90.0
        if out_of_plane:
            z = random.random() * 2 - 1
            phi = math.acos(z) * 180.0 / math.pi
        else:
            phi = polar_angle
        angles.append({"azimuthal": th, "polar": phi})
    return angles


def parse_float_list(s: str) -> list[float]:
    """main.go:260-275 parseFloatList."""
    if s == "":
        return []
    out = []
    for part in s.split(","):
        part = part.strip()
        if part == "":  # main.go:268-270: empty fields are skipped, not errors
            continue
        try:
            out.append(float(part))
        except ValueError:
            raise ValueError(f"invalid float value '{part}'") from None
    return out


def camera_from_angles(azimuthal_deg: float, polar_deg: float, R: float, fov_deg: float = 40.0) -> _lib.XRayCameraParams64:
    cam = _lib.XRayCameraParams64()
    _lib.check(_lib.load().XRayCameraFromAngles(azimuthal_deg, polar_deg, R, fov_deg, ctypes.byref(cam)))
    return cam


def cameras_from_angles(angles, R: float, fov_deg: float):
    """angles: iterable of {'azimuthal','polar'} dicts or (az, polar) pairs -> ctypes array of XRayCameraParams64."""
    angles = list(angles)
    arr = (_lib.XRayCameraParams64 * len(angles))()
    L = _lib.load()
    for k, a in enumerate(angles):
        az, pol = (a["azimuthal"], a["polar"]) if isinstance(a, dict) else a
        _lib.check(L.XRayCameraFromAngles(float(az), float(pol), float(R), float(fov_deg), ctypes.byref(arr[k])))
    return arr


def camera_matrix(cam: _lib.XRayCameraParams64) -> np.ndarray:
    """4x4 camera->world transform as written to transforms.json (main.go:449-453)."""
    return np.array(list(cam.view), dtype=np.float64).reshape(4, 4)


def to_legacy(cams64) -> "ctypes.Array":
    """Narrow to the legacy XRayCameraParams exactly as cuda_path.go:61-77 does (float32 casts)."""
    arr = (_lib.XRayCameraParams * len(cams64))()
    for k, c in enumerate(cams64):
        for a in range(3):
            arr[k].eye[a] = c.eye[a]
        for a in range(16):
            arr[k].view[a] = c.view[a]
        arr[k].fov_y = c.fov_y
        arr[k].R = c.R
    return arr


def from_legacy(cams32):
    """Widen legacy cameras back to fp64 (what the plugin integrates behind the legacy symbol)."""
    arr = (_lib.XRayCameraParams64 * len(cams32))()
    for k, c in enumerate(cams32):
        for a in range(3):
            arr[k].eye[a] = float(c.eye[a])
        for a in range(16):
            arr[k].view[a] = float(c.view[a])
        arr[k].fov_y = float(c.fov_y)
        arr[k].R = float(c.R)
    return arr
