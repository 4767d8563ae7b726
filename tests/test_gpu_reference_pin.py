"""Parity chain anchored on REFERENCE CODE RUN HERE.

The Go host cannot be built (no Go toolchain), but the reference's own CUDA plugin compiles from its two files
(oracle/Makefile -> oracle/_ref/libcuda_render_ref.so, built in the container from /root/reference and shipped to
the GPU box as a binary).  Its kernels are reference code for three things the oracle otherwise restates unpinned:

  * how the `view` matrix, `fov_y`, pixel indices (i <-> camera x, j <-> camera y) and the image layout are used
    (render_kernel, cuda_backend.cu:19-80; cuda_test.go:41-215 compares it with the Go CPU path);
  * the voxel memory layout idx = k*nx*ny + i*ny + j (cuda_backend.cu:103-128);
  * the voxeliser arithmetic of both AssembleVoxelGrid*CUDA symbols (cuda_backend.cu:208-357), bit for bit.

The reference kernel samples with the texture convention (voxel i centred at (i+0.5)/N) while the CPU path is corner
aligned ((N-1) scaling, objects.go:795-802); a smooth analytic field sampled at EACH convention's own voxel positions
removes that known difference, so what is left is interpolation error (O(h^2)) and the 9-bit texture weights.
"""
import ctypes
from pathlib import Path

import numpy as np
import pytest

from helpers import FOV, R, TOL_FP32

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
REF_SO = ROOT / "oracle" / "_ref" / "libcuda_render_ref.so"


@pytest.fixture(scope="module")
def ref(X):
    """ctypes handle of the reference's own plugin (same three symbols, RTLD_LOCAL so they do not clash with ours)."""
    if not REF_SO.exists():
        pytest.skip("oracle/_ref/libcuda_render_ref.so not built (needs /root/reference at build time)")
    import os

    L = ctypes.CDLL(str(REF_SO), mode=os.RTLD_LAZY | os.RTLD_LOCAL)
    fp, ip, c_int, c_float = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_float
    L.RenderVolumeProjectionsCUDA.restype = c_int
    L.RenderVolumeProjectionsCUDA.argtypes = [fp, c_int, c_int, c_int, ctypes.POINTER(X._lib.XRayCameraParams), c_int, c_int,
                                              c_float, c_float, fp]
    L.AssembleVoxelGridCUDA.restype = c_int
    L.AssembleVoxelGridCUDA.argtypes = [ctypes.POINTER(X._lib.CylinderParams), c_int, c_int, c_float, fp]
    L.AssembleVoxelGridSpatialCUDA.restype = c_int
    L.AssembleVoxelGridSpatialCUDA.argtypes = [ctypes.POINTER(X._lib.CylinderParams), c_int, c_int, c_float, c_int, ip, ip,
                                               c_int, fp]
    return L


def blob(x, y, z):
    """Smooth, asymmetric in every axis and in every pair of axes, ~0 at the faces of [-1,1]^3."""
    return 2.5 * np.exp(-((x - 0.22) ** 2 / (2 * 0.16 ** 2) + (y + 0.13) ** 2 / (2 * 0.20 ** 2) + (z - 0.09) ** 2 / (2 * 0.12 ** 2)))


def blob_volume(nx, ny, nz, convention):
    k, i, j = np.meshgrid(np.arange(nz), np.arange(nx), np.arange(ny), indexing="ij")
    if convention == "corner":   # objects.go:795-802: index = (x+1)/2*(N-1)
        x, y, z = 2.0 * i / (nx - 1) - 1, 2.0 * j / (ny - 1) - 1, 2.0 * k / (nz - 1) - 1
    else:                        # texture: texel i is centred at (i+0.5)/N (cuda_backend.cu:70-75)
        x, y, z = (i + 0.5) / nx * 2 - 1, (j + 0.5) / ny * 2 - 1, (k + 0.5) / nz * 2 - 1
    return np.ascontiguousarray(blob(x, y, z), dtype=np.float32)   # layout [z][x][y]


def ref_render(ref, vol, cams32, res, ds32, ff=0.0):
    nz, nx, ny = vol.shape
    out = np.zeros((len(cams32), res, res), dtype=np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = ref.RenderVolumeProjectionsCUDA(vol.ctypes.data_as(fp), nx, ny, nz, cams32, len(cams32), res, ctypes.c_float(ds32),
                                         ctypes.c_float(ff), out.ctypes.data_as(fp))
    assert rc == 0
    return out


@pytest.mark.parametrize("shape", [(128, 128, 128), (96, 128, 80)])
def test_reference_kernel_pins_camera_use_pixel_mapping_and_layout(X, O, ref, shape):
    """a2 / a19: the reference's render_kernel and the oracle (and this library) must see the same picture of an asymmetric
    smooth blob, to 5e-3 (the reference's own CPU-vs-CUDA bar is 0.10, cuda_test.go:117-122); swapping i and j, passing the
    transposed (column-major) view matrix, or permuting the voxel axes must break it by an order of magnitude."""
    nx, ny, nz = shape
    res = 64
    views = [(0.0, 90.0), (37.0, 72.0), (118.0, 105.0), (251.0, 90.0)]
    cams = X.cameras_from_angles(views, R, FOV)
    cams32 = X.to_legacy(cams)
    ds32 = float(np.float32(2.0 / min(shape) / 5.0))
    vol_c = blob_volume(nx, ny, nz, "corner")
    vol_t = blob_volume(nx, ny, nz, "texture")
    img_ref = ref_render(ref, vol_t, cams32, res, ds32).astype(np.float64)
    assert img_ref.min() < 0.5 and img_ref.max() > 0.999            # the blob really attenuates, the corners are empty

    # oracle fed the fp32-rounded values that cross the legacy boundary
    osc = O.OracleScene({"type": "voxel_grid", "_array": vol_c.astype(np.float64)})
    cw = X.from_legacy(cams32)
    img_orc = np.stack([osc.render_view(np.array(list(c.eye)), np.array(list(c.view)).reshape(4, 4), res, float(c.fov_y),
                                        float(c.R), ds32, "simple")[0] for c in cw])
    img_gpu = X.render_volume_legacy(vol_c, cams32, res, ds32).astype(np.float64)

    tol = 5e-3
    e_orc = np.abs(img_orc - img_ref).max()
    e_gpu = np.abs(img_gpu - img_ref).max()
    assert e_orc <= tol, f"oracle vs reference kernel: {e_orc:.2e}"
    assert e_gpu <= tol, f"library vs reference kernel: {e_gpu:.2e}"
    assert np.abs(img_gpu - img_orc).max() <= TOL_FP32

    # negative controls: each of these conventions is really pinned by the comparison above
    assert np.abs(img_orc.transpose(0, 2, 1) - img_ref).max() > 10 * tol          # i <-> j
    cams_T = X.from_legacy(cams32)
    for c in cams_T:                                                               # column-major instead of row-major view
        m = np.array(list(c.view)).reshape(4, 4).T.copy().reshape(16)
        for a in range(16):
            c.view[a] = m[a]
    img_T = np.stack([osc.render_view(np.array(list(c.eye)), np.array(list(c.view)).reshape(4, 4), res, float(c.fov_y),
                                      float(c.R), ds32, "simple")[0] for c in cams_T])
    assert np.abs(img_T - img_ref).max() > 10 * tol
    if nx == ny == nz:                                                             # [z][x][y] versus [z][y][x]
        osc_sw = O.OracleScene({"type": "voxel_grid", "_array": vol_c.transpose(0, 2, 1).astype(np.float64)})
        img_sw = np.stack([osc_sw.render_view(np.array(list(c.eye)), np.array(list(c.view)).reshape(4, 4), res, float(c.fov_y),
                                              float(c.R), ds32, "simple")[0] for c in cw])
        assert np.abs(img_sw - img_ref).max() > 10 * tol
    # fov is in degrees and vertical = horizontal (square detector): a 2 degree error is visible
    cams_f = X.from_legacy(cams32)
    for c in cams_f:
        c.fov_y = float(c.fov_y) + 2.0
    img_f = np.stack([osc.render_view(np.array(list(c.eye)), np.array(list(c.view)).reshape(4, 4), res, float(c.fov_y),
                                      float(c.R), ds32, "simple")[0] for c in cams_f])
    assert np.abs(img_f - img_ref).max() > 2 * tol


# ---- voxeliser: both legacy symbols against the same symbols of the reference, bit for bit ------------------------
KELVIN = [  # objects.go:588-625 MakeKelvin strut table (unit cell [0,1]^3)
    ((0.25, 0.00, 0.50), (0.50, 0.00, 0.75)), ((0.25, 1.00, 0.50), (0.50, 1.00, 0.75)), ((0.25, 0.00, 0.50), (0.50, 0.00, 0.25)),
    ((0.25, 1.00, 0.50), (0.50, 1.00, 0.25)), ((0.25, 0.00, 0.50), (0.00, 0.25, 0.50)), ((0.50, 0.00, 0.75), (0.75, 0.00, 0.50)),
    ((0.50, 1.00, 0.75), (0.75, 1.00, 0.50)), ((0.50, 0.00, 0.75), (0.50, 0.25, 1.00)), ((0.75, 0.00, 0.50), (0.50, 0.00, 0.25)),
    ((0.75, 1.00, 0.50), (0.50, 1.00, 0.25)), ((0.75, 0.00, 0.50), (1.00, 0.25, 0.50)), ((0.50, 0.00, 0.25), (0.50, 0.25, 0.00)),
    ((1.00, 0.50, 0.75), (0.75, 0.50, 1.00)), ((1.00, 0.75, 0.50), (0.75, 1.00, 0.50)), ((1.00, 0.50, 0.25), (0.75, 0.50, 0.00)),
    ((0.25, 1.00, 0.50), (0.00, 0.75, 0.50)), ((0.50, 1.00, 0.75), (0.50, 0.75, 1.00)), ((0.50, 1.00, 0.25), (0.50, 0.75, 0.00)),
    ((0.00, 0.25, 0.50), (0.00, 0.50, 0.75)), ((1.00, 0.25, 0.50), (1.00, 0.50, 0.75)), ((0.00, 0.25, 0.50), (0.00, 0.50, 0.25)),
    ((1.00, 0.25, 0.50), (1.00, 0.50, 0.25)), ((0.00, 0.50, 0.75), (0.25, 0.50, 1.00)), ((0.00, 0.50, 0.75), (0.00, 0.75, 0.50)),
    ((1.00, 0.50, 0.75), (1.00, 0.75, 0.50)), ((0.00, 0.75, 0.50), (0.00, 0.50, 0.25)), ((1.00, 0.75, 0.50), (1.00, 0.50, 0.25)),
    ((0.00, 0.50, 0.25), (0.25, 0.50, 0.00)), ((0.25, 0.50, 0.00), (0.50, 0.75, 0.00)), ((0.25, 0.50, 1.00), (0.50, 0.75, 1.00)),
    ((0.25, 0.50, 0.00), (0.50, 0.25, 0.00)), ((0.25, 0.50, 1.00), (0.50, 0.25, 1.00)), ((0.50, 0.75, 0.00), (0.75, 0.50, 0.00)),
    ((0.50, 0.75, 1.00), (0.75, 0.50, 1.00)), ((0.75, 0.50, 0.00), (0.50, 0.25, 0.00)), ((0.75, 0.50, 1.00), (0.50, 0.25, 1.00)),
]


def kelvin_cylinders(X, cells=4, rad=0.03, rho=0.6):
    """A cells^3 Kelvin lattice filling [-1,1]^3 as the flat cylinder list extractCylinders (cuda_voxel.go:13-35) hands over."""
    s = 2.0 / cells
    cyls = []
    for cx in range(cells):
        for cy in range(cells):
            for cz in range(cells):
                off = np.array([cx, cy, cz]) * s - 1.0
                for p0, p1 in KELVIN:
                    cyls.append((np.array(p0) * s + off, np.array(p1) * s + off, rad, rho))
    arr = (X._lib.CylinderParams * len(cyls))()
    for k, (p0, p1, r, rh) in enumerate(cyls):
        arr[k].p0[:] = [float(v) for v in p0]
        arr[k].p1[:] = [float(v) for v in p1]
        arr[k].radius = r
        arr[k].rho = rh
    return arr


def go_host_csr(arr, G):
    """cuda_backend.go:182-237: the CSR uniform grid the Go host builds before AssembleVoxelGridSpatialCUDA."""
    cs = 2.0 / G
    clampg = lambda v: max(0, min(G - 1, v))
    lists = [[] for _ in range(G ** 3)]
    for ci in range(len(arr)):
        c = arr[ci]
        lo = [min(float(c.p0[a]), float(c.p1[a])) - float(c.radius) for a in range(3)]
        hi = [max(float(c.p0[a]), float(c.p1[a])) + float(c.radius) for a in range(3)]
        rg = [range(clampg(int((lo[a] + 1.0) / cs)), clampg(int((hi[a] + 1.0) / cs)) + 1) for a in range(3)]
        for cz in rg[2]:
            for cy in rg[1]:
                for cx in rg[0]:
                    lists[(cz * G + cy) * G + cx].append(ci)
    offs = np.zeros(G ** 3 + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(l) for l in lists])
    idx = np.array([c for l in lists for c in l], dtype=np.int32)
    return offs, idx


@pytest.mark.parametrize("cells,res,dm", [(4, 128, 1.0), (2, 96, 2.5), (1, 33, 0.4)])
def test_voxeliser_symbols_bit_identical_to_reference(X, ref, cells, res, dm):
    """a20: AssembleVoxelGridCUDA and AssembleVoxelGridSpatialCUDA against the reference's own kernels
    (cuda_backend.cu:208-252, 297-357), Kelvin 4x4x4 (2304 struts, the case of cuda_voxel.go:37), G = 16 -- every voxel equal
    bit for bit, no margin mask."""
    L = X._lib.load()
    arr = kelvin_cylinders(X, cells)
    n = len(arr)
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
    want = np.zeros((res, res, res), dtype=np.float32)
    assert ref.AssembleVoxelGridCUDA(arr, n, res, ctypes.c_float(dm), want.ctypes.data_as(fp)) == 0
    got = np.full_like(want, -1.0)
    assert L.AssembleVoxelGridCUDA(arr, n, res, ctypes.c_float(dm), got.ctypes.data_as(fp)) == 0
    assert 0.001 < (want > 0).mean() < 0.9
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{(got != want).sum()} voxels differ (brute force)"
    G = 16
    offs, idx = go_host_csr(arr, G)
    want_s = np.zeros_like(want)
    assert ref.AssembleVoxelGridSpatialCUDA(arr, n, res, ctypes.c_float(dm), G, offs.ctypes.data_as(ip), idx.ctypes.data_as(ip),
                                            len(idx), want_s.ctypes.data_as(fp)) == 0
    got_s = np.full_like(want, -1.0)
    assert L.AssembleVoxelGridSpatialCUDA(arr, n, res, ctypes.c_float(dm), G, offs.ctypes.data_as(ip), idx.ctypes.data_as(ip),
                                          len(idx), got_s.ctypes.data_as(fp)) == 0
    assert np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32)), f"{(got_s != want_s).sum()} voxels differ (spatial)"


def test_voxeliser_overlapping_and_degenerate_cylinders_vs_reference(X, ref):
    """Sums of overlapping struts (clamped after the multiplier), a zero-length cylinder (skipped, cuda_backend.cu:236),
    negative rho, caps exactly on voxel planes."""
    L = X._lib.load()
    cyls = [((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5), 0.2, 0.7), ((-0.6, 0.5, 0.0), (0.6, 0.5, 0.1), 0.1, 0.6),
            ((0.0, 0.0, -0.8), (0.0, 0.0, 0.8), 0.15, 0.5), ((0.1, 0.1, 0.1), (0.1, 0.1, 0.1), 0.3, 0.9),
            ((-0.5, 0.0, 0.0), (0.5, 0.0, 0.0), 0.25, -0.4), ((-0.75, -0.75, -1.0), (-0.75, -0.75, 0.25), 0.125, 0.3)]
    arr = (X._lib.CylinderParams * len(cyls))()
    for k, (p0, p1, r, rho) in enumerate(cyls):
        arr[k].p0[:] = p0
        arr[k].p1[:] = p1
        arr[k].radius = r
        arr[k].rho = rho
    fp = ctypes.POINTER(ctypes.c_float)
    for res, dm in ((64, 1.5), (50, 1.0)):
        want = np.zeros((res, res, res), dtype=np.float32)
        got = np.full_like(want, -1.0)
        assert ref.AssembleVoxelGridCUDA(arr, len(cyls), res, ctypes.c_float(dm), want.ctypes.data_as(fp)) == 0
        assert L.AssembleVoxelGridCUDA(arr, len(cyls), res, ctypes.c_float(dm), got.ctypes.data_as(fp)) == 0
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{(got != want).sum()} voxels differ"


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_tile_culled_voxeliser_random_cylinders_vs_reference(X, ref, seed):
    """The brute-force symbol culls the cylinder list per 4 x 8 x 8 voxel tile before the per-voxel test
    (misc_kernels.cu); every voxel must still equal the reference's all-pairs kernel bit for bit: random segments of all
    lengths and radii (incl. tiny, huge, reversed, negative radius and rho, far outside the volume), res not a multiple
    of the tile, and enough cylinders through one place that the shared-memory list is flushed mid-way (> 1024)."""
    L = X._lib.load()
    rng = np.random.default_rng(100 + seed)
    n = [300, 2500, 1500, 40][seed]
    res = [77, 61, 40, 130][seed]
    arr = (X._lib.CylinderParams * n)()
    for k in range(n):
        if seed == 1 and k % 2 == 0:   # a sheaf through the same point: > 1024 survivors in the tiles around it
            c = np.array([0.13, -0.21, 0.34])
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            p0, p1 = c - d * rng.uniform(0.05, 1.5), c + d * rng.uniform(0.05, 1.5)
            r = rng.uniform(0.01, 0.05)
        else:
            p0 = rng.uniform(-1.4, 1.4, 3)
            p1 = p0 + rng.normal(size=3) * rng.choice([1e-6, 0.02, 0.3, 1.0, 3.0])
            r = rng.choice([1e-4, 0.01, 0.05, 0.2, 0.7]) * rng.uniform(0.5, 1.5)
        if k % 17 == 0:
            r = -r          # the test compares with r*r
        if k % 23 == 0:
            p1 = p0.copy()  # zero length: skipped
        if k % 29 == 0:
            p0, p1 = p0 * 40.0, p1 * 40.0   # far away
        arr[k].p0[:] = [float(v) for v in p0]
        arr[k].p1[:] = [float(v) for v in p1]
        arr[k].radius = float(r)
        arr[k].rho = float(rng.choice([0.3, 0.05, -0.2, 1.0]) * rng.uniform(0.5, 1.0))
    fp = ctypes.POINTER(ctypes.c_float)
    dm = [1.0, 0.02, 0.7, -1.0][seed]
    want = np.zeros((res, res, res), dtype=np.float32)
    got = np.full_like(want, -1.0)
    assert ref.AssembleVoxelGridCUDA(arr, n, res, ctypes.c_float(dm), want.ctypes.data_as(fp)) == 0
    assert L.AssembleVoxelGridCUDA(arr, n, res, ctypes.c_float(dm), got.ctypes.data_as(fp)) == 0
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{(got != want).sum()} voxels differ"
    if seed < 3:
        assert 0.0 < (want > 0).mean() < 1.0
