"""B200-native per-pixel X-ray projection path (drop-in for igrega348/xray_projection_render's
libcuda_render.so plugin, extended to analytic scenes).  See DESIGN.md / INTEGRATION.md."""
from . import _lib
from .camera import (CUBE_HALF_DIAGONAL, camera_from_angles, camera_matrix, cameras_from_angles,
                     generate_camera_angles, parse_float_list, to_legacy, from_legacy)
from .renderer import (XRayRenderer, measure_fp32_peak, render_scene, render_scene_device, render_volume,
                       render_volume_device, render_volume_device_to_host, render_volume_legacy, voxelize_scene, image_to_rgba8, write_png)
from .scene import Scene, SceneError, load_map, normalize_deformation, normalize_object, voxel_grid_from_raw

__all__ = [
    "XRayRenderer", "Scene", "SceneError", "render_scene", "render_scene_device", "render_volume",
    "render_volume_device", "render_volume_device_to_host", "render_volume_legacy", "voxelize_scene", "measure_fp32_peak",
    "camera_from_angles", "cameras_from_angles", "camera_matrix", "generate_camera_angles", "parse_float_list",
    "to_legacy", "from_legacy", "load_map", "normalize_object", "normalize_deformation", "voxel_grid_from_raw",
    "image_to_rgba8", "write_png", "CUBE_HALF_DIAGONAL",
]
