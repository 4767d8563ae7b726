"""Extract the one piece of REFERENCE OUTPUT the reference repository carries: the image stored in
`examples/demo.ipynb` (code cell 7, `display_data`, `image/png`) -- three equispaced 300 x 300 projections of
`examples/cube_w_hole.yaml` pasted side by side into a 900 x 300 greyscale image by the notebook (cell 6's
`display(dst)`; the stored output is from an earlier run of the notebook than its present cell sources).

It was rendered by the reference's Go library (`RenderProjections` through `XRayRenderer.render`), written as an
RGBA PNG by `main.go:482-546`, and read back by PIL (`Image.open(...)` pasted into an 'L' image: R = G = B, so L = R).
Nothing in this repository produced it.  The render parameters are not recorded in the notebook; the oracle reproduces
all 270 000 pixels (52 067 attenuated, 154 distinct grey levels) EXACTLY with R = 5, fov = 45, ds = 0.1,
hierarchical integrator, azimuths 90 / 210 / 330, polar 90 -- and with nothing nearby: ds = 0.1001 changes 40 504
pixels, R = 5.001 1 914, fov = 45.01 2 257, the simple integrator 45 699 (tests/test_reference_go_output.py).

Run in the build container (needs /root/reference and PIL):  python tests/golden/extract_notebook_image.py
Writes  tests/golden/reference_go_cube_w_hole_3x300.png   (the stored bytes, untouched)
        tests/golden/reference_go_cube_w_hole_3x300.npz   (the same image decoded: uint8 [300, 900], for numpy-only loading)
"""
import base64
import io
import json
from pathlib import Path

import numpy as np

NOTEBOOK = Path("/root/reference/examples/demo.ipynb")
OUT = Path(__file__).resolve().parent


def main():
    nb = json.loads(NOTEBOOK.read_text())
    found = []
    for idx, cell in enumerate(nb["cells"]):
        for out in cell.get("outputs", []):
            data = out.get("data", {})
            if "image/png" in data:
                found.append((idx, data["image/png"]))
    assert len(found) == 1, f"expected exactly one stored image, found {len(found)}"
    idx, b64 = found[0]
    png = base64.b64decode("".join(b64) if isinstance(b64, list) else b64)
    (OUT / "reference_go_cube_w_hole_3x300.png").write_bytes(png)
    from PIL import Image

    im = Image.open(io.BytesIO(png))
    assert im.mode == "L" and im.size == (900, 300), (im.mode, im.size)
    arr = np.array(im)
    np.savez_compressed(OUT / "reference_go_cube_w_hole_3x300.npz", image=arr,
                        source=f"/root/reference/examples/demo.ipynb cell {idx} outputs[0].data['image/png']")
    print("cell", idx, arr.shape, arr.dtype, "attenuated pixels", int((arr < 255).sum()), "levels", len(np.unique(arr)))


if __name__ == "__main__":
    main()
