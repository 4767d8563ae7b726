#!/usr/bin/env python3
"""The reference's Python workflow (its example_usage.py / demo notebook) on the B200 plugin.

`xray_projection_render_b200.XRayRenderer` takes the same parameter dictionary and returns the same result dictionary as
the reference's `XRayRenderer` (xray_projection_render/xray_renderer.py:357-448) and writes the same files (PNG frames,
transforms.json, object.json; volume.raw with export_volume).  Needs a CUDA device: there is no CPU fallback.

    python integration/example_usage.py [scene file] [output directory]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from xray_projection_render_b200 import XRayRenderer  # noqa: E402


def main() -> int:
    scene = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "tests" / "scenes" / "cube_w_hole.json")
    out = Path(sys.argv[2] if len(sys.argv) > 2 else "output")
    renderer = XRayRenderer()  # XRayRenderer(precision="fp64") reproduces the Go arithmetic to 1e-15; devices=[0, 1, ...] shards views

    # 1. chosen camera angles (degrees), as `camera_angles=` or params["camera_angles"]
    result = renderer.render({"input": scene, "output_dir": str(out / "custom" / "images"), "resolution": 256,
                              "transforms_file": str(out / "custom" / "transforms.json")},
                             camera_angles=[{"azimuthal": a, "polar": 90.0} for a in (0.0, 90.0, 180.0, 270.0)])
    print("custom angles:", result)

    # 2. equispaced views; the run of the reference's demo notebook whose stored image the tests reproduce bit for bit
    result = renderer.render({"input": scene, "output_dir": str(out / "ring" / "images"), "num_images": 3, "resolution": 300,
                              "R": 5.0, "fov": 45.0, "ds": 0.1, "transforms_file": str(out / "ring" / "transforms.json")})
    print("equispaced:", result)

    # 3. a shard of a larger job (the reference's --jobs_modulo / --job), simple integrator, flat field
    result = renderer.render({"input": scene, "output_dir": str(out / "shard" / "images"), "num_images": 16, "resolution": 512,
                              "jobs_modulo": 4, "job_num": 1, "integration": "simple", "ds": 0.01, "flat_field": 0.1,
                              "transforms_file": str(out / "shard" / "transforms.json")})
    print("shard 1 of 4:", result)
    return 0 if result["success"] else 1


if __name__ == "__main__":
    sys.exit(main())
