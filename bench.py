#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-pixel projection hot path.

python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-configs]

A "step" is one pass of the hot path over one batch: all views of the workload.
Headline workload = BASELINE.json configs[1]: tests/scenes/lattice.json (= examples/lattice.yaml),
360 views at 1024x1024, hierarchical integrator, auto ds, R=4, fov=40, polar=90, fp32 mode.
The same JSON line carries the other BASELINE configs as sub-records under "configs" (cfg1 in full; cfg3, cfg4, cfg5 on
a bounded number of their views, stated), each with its device-resident value, end-to-end value, roofline and an in-run
parity spot check against the oracle.
N>1 (torchrun, one rank per GPU): STRONG scaling of the fixed configs -- rank r renders views r, r+N, ... of the same
view list (the reference's --jobs_modulo/--job sharding, main.go:244); analytic scenes need no data-path collective; the
voxel config uploads the volume on rank 0 only and replicates it with one NCCL broadcast INSIDE the timed end-to-end
region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

SCENES = ROOT / "tests" / "scenes"

# workload name -> (object file, deformation file, views, res, integrator, ds (<=0 auto))
WORKLOADS = {
    "cube_w_hole": ("cube_w_hole.json", None, 1, 512, "hierarchical", -1.0),                  # configs[0]
    "lattice": ("lattice.json", None, 360, 1024, "hierarchical", -1.0),                        # configs[1]
    "gyroid_sigmoid": ("gyroid_example.json", "deformation_sigmoid.json", 720, 1024, "hierarchical", -1.0),  # configs[2]
    "voxel1024": (None, None, 1440, 2048, "simple", -1.0),                                     # configs[3] (8 GPUs)
    "pillar_array": ("pillar_array.json", None, 2880, 4096, "hierarchical", -1.0),             # configs[4]
    # the reference's two remaining example scenes (not BASELINE configs; regression coverage for sphere / box / pped runs)
    "lattice_linear": ("lattice.json", "deformation_linear.json", 360, 1024, "hierarchical", -1.0),   # the project's use case: a strained lattice
    "lattice_sigmoid": ("lattice.json", "deformation_sigmoid.json", 360, 1024, "hierarchical", -1.0),
    "balls": ("balls.json", None, 360, 1024, "hierarchical", -1.0),
    "box_w_pped": ("box_w_pped.json", None, 360, 1024, "hierarchical", -1.0),
}
R_CAM, FOV, POLAR = 4.0, 40.0, 90.0

# Algorithmic fp32 flops (FMA = 2), SURVEY.md 8(d): per evaluated sample and per primitive test.
FLOPS_POSITION = 6
FLOPS_TESS = 27          # 12 bound compares + fold 15
FLOPS_PRIM = {"cylinder": 25 + 2, "sphere": 9, "box": 9, "cube": 9, "parallelepiped": 24, "gyroid": 13 + 72}
FLOPS_COLLECTION = 2     # clamp; +1 per child test is folded into the primitive figure below


def scene_flops_model(obj: dict, deform: dict | None):
    """(flops per evaluated sample, flops per primitive test) for roofline accounting."""
    per_sample = FLOPS_POSITION
    if deform:
        per_sample += {"sigmoid": 6 + 12 + 2, "linear": 18, "affine": 15, "rigid": 3, "gaussian": 3 * 14 + 8}.get(deform["type"], 0)
    prim = []

    def walk(o):
        nonlocal per_sample
        t = o["type"]
        if t == "tessellated_obj_coll":
            per_sample += FLOPS_TESS + FLOPS_COLLECTION
            for c in o["uc"]["objects"]["objects"]:
                walk(c)
        elif t == "object_collection":
            per_sample += FLOPS_COLLECTION
            for c in o["objects"]:
                walk(c)
        elif t == "voxel_grid":
            prim.append(3 + 6 + 14)
        else:
            prim.append(FLOPS_PRIM[t] + 1)

    walk(obj)
    return per_sample, (sum(prim) / len(prim) if prim else 0.0), float(sum(prim))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def synthetic_volume(n: int) -> np.ndarray:
    """clamp(0.5+0.5 sin(6 pi x) sin(10 pi y) sin(14 pi z), 0, 1) * [r < 0.9], layout [z][x][y]."""
    ax = (2.0 * np.arange(n, dtype=np.float64) / (n - 1) - 1.0)
    sx, sy, sz = np.sin(6 * np.pi * ax), np.sin(10 * np.pi * ax), np.sin(14 * np.pi * ax)
    vol = np.empty((n, n, n), dtype=np.float32)
    r2x = ax * ax
    for k in range(n):
        v = 0.5 + 0.5 * sz[k] * np.outer(sx, sy)
        mask = (r2x[:, None] + r2x[None, :] + r2x[k]) < 0.81
        vol[k] = (np.clip(v, 0.0, 1.0) * mask).astype(np.float32)
    return vol


def rank_views(total_views: int, rank: int, world: int):
    from xray_projection_render_b200 import generate_camera_angles

    return generate_camera_angles(total_views, rank, world, False, POLAR)


# ----------------------------------------------------------------------------------------
# Workload description shared by both arms (so that the driver sees the same `config.workload`)
# ----------------------------------------------------------------------------------------
def workload_string(name: str, volume_n: int = 1024) -> str:
    obj, deform, views, res, integ, ds = WORKLOADS[name]
    scene = obj if obj else f"synthetic {volume_n}^3 fp32 volume"
    if deform:
        scene += " + " + deform
    return f"{name}: {scene}, {views} views at {res}x{res}, {integ}, auto ds, R={R_CAM}, fov={FOV}, polar={POLAR}"


# ----------------------------------------------------------------------------------------
# CPU baseline (oracle = C++ restatement of the Go path; kind "port")
# ----------------------------------------------------------------------------------------
def cpu_baseline(workload: str, budget_s: float = 15.0, nthreads: int = 0, volume=None):
    """The reference's CPU path (oracle/xray_oracle.cpp, all host threads) on a bounded sample of the workload's rays,
    sized for ~budget_s of wall time after a short calibration."""
    from oracle import oracle as O

    obj, deform, views, res, integ, ds = WORKLOADS[workload]
    if obj is None:
        if volume is None:
            volume = synthetic_volume(256)
        osc = O.OracleScene({"type": "voxel_grid", "_array_f32": np.ascontiguousarray(volume, dtype=np.float32)})
        ds_v = 2.0 / 1024 / 5.0  # the full-size workload's step: same per-sample cost model
        sample_res = 64
    else:
        osc = O.OracleScene(str(SCENES / obj), str(SCENES / deform) if deform else None)
        ds_v = osc.auto_ds() if ds <= 0 else ds
        sample_res = min(res, 256)
    # all host cores this process may use -- not OMP_NUM_THREADS, which torchrun pins to 1 for every rank
    threads = nthreads or max(O.max_threads(), len(os.sched_getaffinity(0)))
    angles = O.generate_camera_angles(views)
    # calibrate on a strip that takes a good fraction of a second, then size the sample for ~budget_s of CPU work
    eye, cm = O.camera_from_angles(*angles[0], R_CAM)
    mid = sample_res // 2
    rows0 = max(1, threads // 4)
    while True:
        t0 = time.perf_counter()
        osc.render_view(eye, cm, sample_res, FOV, R_CAM, ds_v, integ, rows=(max(0, mid - rows0 // 2), min(sample_res, mid - rows0 // 2 + rows0)),
                        nthreads=threads)
        dt0 = max(time.perf_counter() - t0, 1e-4)
        if dt0 > 0.25 or rows0 >= sample_res:
            break
        rows0 = min(sample_res, rows0 * 4)
    rows_per_s = rows0 / dt0
    want_rows = int(min(64 * sample_res, max(threads, rows_per_s * budget_s)))
    n_views = max(1, min(64, -(-want_rows // sample_res)))
    rows_each = max(1, min(sample_res, want_rows // n_views))
    r0 = (sample_res - rows_each) // 2
    picks = [angles[(len(angles) * k) // n_views] for k in range(n_views)]
    samples = rays = 0
    t0 = time.perf_counter()
    for az, pol in picks:
        eye, cm = O.camera_from_angles(az, pol, R_CAM)
        _, n = osc.render_view(eye, cm, sample_res, FOV, R_CAM, ds_v, integ, rows=(r0, r0 + rows_each), nthreads=threads)
        samples += n
        rays += rows_each * sample_res
    dt = time.perf_counter() - t0
    desc = (f"{n_views} views x {rows_each} central rows at {sample_res}^2 of {workload} ({rays} rays, {samples} samples, "
            f"{dt:.1f} s); per-ray cost is resolution independent")
    return {"gsamples_per_s": samples / dt / 1e9, "rays_per_s": rays / dt, "cores": threads, "kind": "port",
            "sample": desc, "seconds": dt}


def time_reference_cuda_plugin(X, vol_np, cams, res, ds, ref_samples_per_step):
    """Second comparator of BASELINE.md: the reference's own CUDA plugin (cuda_backend.cu built unmodified for
    sm_100 into oracle/_ref/ by oracle/Makefile), driven through the same legacy call with the same host
    buffers.  Speed only: its images follow the half-voxel texture convention (cuda_test.go:33-40)."""
    lib_path = ROOT / "oracle" / "_ref" / "libcuda_render_ref.so"
    if not lib_path.exists():
        return {"unavailable": "oracle/_ref/libcuda_render_ref.so not built (needs /root/reference at build time)"}
    try:
        ref = ctypes.CDLL(str(lib_path), mode=os.RTLD_LAZY | os.RTLD_LOCAL)
        fp = ctypes.POINTER(ctypes.c_float)
        ref.RenderVolumeProjectionsCUDA.restype = ctypes.c_int
        ref.RenderVolumeProjectionsCUDA.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                    ctypes.POINTER(X._lib.XRayCameraParams), ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_float, ctypes.c_float, fp]
        cams32 = X.to_legacy(cams)
        n = len(cams32)
        nz, nx, ny = vol_np.shape  # pageable, like the Go heap buffer the host hands over (cuda_backend.go:316)
        out = np.empty((n, res, res), dtype=np.float32)
        ours = np.empty((n, res, res), dtype=np.float32)
        args = (vol_np.ctypes.data_as(fp), nx, ny, nz, cams32, n, res, ctypes.c_float(ds), ctypes.c_float(0.0))
        rc = ref.RenderVolumeProjectionsCUDA(*args, out.ctypes.data_as(fp))  # warm-up (context, allocations)
        if rc != 0:
            return {"unavailable": f"reference plugin returned {rc}"}
        t0 = time.perf_counter()
        ref.RenderVolumeProjectionsCUDA(*args, out.ctypes.data_as(fp))
        t_ref = time.perf_counter() - t0
        X.render_volume_legacy(vol_np, cams32, res, float(np.float32(ds)), out=ours)
        t0 = time.perf_counter()
        X.render_volume_legacy(vol_np, cams32, res, float(np.float32(ds)), out=ours)
        t_ours = time.perf_counter() - t0
        return {"what": "reference cuda_backend.cu (sm_100 build) vs this library, same RenderVolumeProjectionsCUDA call, same pageable host buffers (as the Go host passes)",
                "views": n, "reference_s": t_ref, "ours_s": t_ours, "reference_gsamples_per_s": ref_samples_per_step / t_ref / 1e9,
                "ours_gsamples_per_s": ref_samples_per_step / t_ours / 1e9, "speedup": t_ref / t_ours,
                "max_abs_image_diff": float(np.abs(out - ours).max()),
                "note": "image difference is the reference kernel's half-voxel texture convention, not an error of either"}
    except OSError as exc:
        return {"unavailable": str(exc)}


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU implementation of the path (C++ restatement of the Go
    code, all host threads) on the same workload/metric.  Rank 0 only.  Every step is a bounded sample of at least
    3 s of wall time (short samples were dominated by thread start-up), so the step count is capped at 5."""
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    per_step = max(3.0, args.cpu_budget / steps)
    vol = synthetic_volume(256) if WORKLOADS[args.workload][0] is None else None
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_baseline(args.workload, budget_s=1.0, volume=vol)
    vals = []
    t_all = time.perf_counter()
    for _ in range(steps):
        vals.append(cpu_baseline(args.workload, budget_s=per_step, volume=vol))
    wall = time.perf_counter() - t_all
    samples = sum(x["gsamples_per_s"] * x["seconds"] for x in vals)
    secs = sum(x["seconds"] for x in vals)
    v = samples / secs
    rays = sum(x["rays_per_s"] * x["seconds"] for x in vals) / secs
    line = {
        "impl": "reference", "metric": "Gsamples/s", "value": v, "unit": "Gsamples/s", "n_gpus": args.gpus, "steps": steps,
        "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": wall / steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "projections_512_per_s": rays / 262144.0,
        "config": {"workload": workload_string(args.workload, args.volume_n),
                   "note": "CPU path timed on a bounded sample of the workload's rays (see cpu_baseline.sample); the rate is per sample, "
                           "so it does not depend on how many GPUs the other arm uses"},
        "cpu_baseline": {"value": v, "unit": "Gsamples/s", "cores": vals[-1]["cores"], "kind": "port", "sample": vals[-1]["sample"],
                         "per_step_values": [x["gsamples_per_s"] for x in vals]},
        "e2e": {"value": v, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# Our arm
# ----------------------------------------------------------------------------------------
# Algorithmic work of the interval renderer (render_span.cu), fp32-equivalent flops (an fp64 operation counted as 2):
# per candidate that reaches the exact stage: 3 period slabs + cap slab + quadratic + two ordinal look-ups ~ 110 fp64 ops;
# per ray: ray set-up, outer-box clip, sweep ~ 90 fp64 ops + 40 fp32.
FLOPS_SPAN_CANDIDATE = 220
FLOPS_SPAN_RAY = 220
STAT_KEYS = ("ref_samples", "evaluated_samples", "fp64_fallbacks", "primitive_tests", "rays", "launches", "marched_tiles")


def measure(X, torch, dist, name, rank, world, local_rank, steps, warmup, flush, views_total=0, volume_n=1024, volume=None,
            parity=True, ref_cuda=False, sampler=None, ncu_info=None):
    """Time one workload on this process group: device-resident value, end-to-end value (pageable and pinned host buffers),
    roofline, parity spot check.  Returns the record on rank 0, None elsewhere."""
    dev = torch.device("cuda", local_rank)
    obj, deform, views_full, res, integ, ds = WORKLOADS[name]
    views_total = views_total or views_full
    angles = rank_views(views_total, rank, world)  # strong scaling: this rank's share of the fixed view list
    cams = X.cameras_from_angles(angles, R_CAM, FOV)
    n = len(cams)
    stream = torch.cuda.current_stream()
    out_dev = torch.empty((max(n, 1), res, res), dtype=torch.float32, device=dev)
    out_page = np.empty((max(n, 1), res, res), dtype=np.float32)            # pageable: what a Go / numpy caller hands over
    out_pin = torch.empty((max(n, 1), res, res), dtype=torch.float32, pin_memory=True)
    stats = (ctypes.c_uint64 * X._lib.XRAY_NUM_STATS)()
    is_volume = obj is None
    bcast_ms = []

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if is_volume:
        nvox = volume_n
        if rank == 0 and volume is None:
            volume = synthetic_volume(nvox)
        vol_np = volume if rank == 0 else None   # pageable numpy array on rank 0 only
        vol_pin = torch.from_numpy(vol_np).pin_memory() if rank == 0 else None
        vol_dev = torch.empty((nvox, nvox, nvox), dtype=torch.float32, device=dev)
        if rank == 0:
            vol_dev.copy_(vol_pin, non_blocking=True)
        if dist is not None:
            dist.broadcast(vol_dev, src=0)
        ds_v = 2.0 / nvox / 5.0
        scene = None

        def step_device(st=None):
            if n:
                X.render_volume_device(vol_dev, (nvox, nvox, nvox), cams, res, out_dev, ds=ds_v, stream=stream.cuda_stream, stats=st)

        if world == 1:
            def step_e2e(host_out):  # the call the Go host makes, volume re-uploaded from host memory every call
                X.render_volume(vol_np, cams, res, ds=ds_v, out=host_out)
            h2d_step = vol_dev.numel() * 4
            e2e_api = "XRayRenderVolumeExCUDA"
        else:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

            def step_e2e(host_out):
                # rank 0 uploads the volume ONCE, one NCCL broadcast over NVLink replicates it, every rank renders its views
                # from the device copy and returns its images to host memory
                if rank == 0:
                    vol_dev.copy_(vol_pin, non_blocking=True)
                ev0.record(stream)
                dist.broadcast(vol_dev, src=0)
                ev1.record(stream)
                stream.synchronize()   # the library renders on its own stream: the broadcast has to be complete
                if n:  # kernels, image D2H and the copy into the caller's buffer overlapped inside the library
                    X.render_volume_device_to_host(vol_dev, (nvox, nvox, nvox), cams, res, host_out, ds=ds_v)
                bcast_ms.append(ev0.elapsed_time(ev1))
            h2d_step = vol_dev.numel() * 4 if rank == 0 else 0
            e2e_api = "pinned H2D on rank 0 + ncclBroadcast + XRayRenderVolumeDeviceToHostCUDA"
    else:
        scene = X.Scene(str(SCENES / obj), str(SCENES / deform) if deform else None)
        ds_v = scene.auto_ds() if ds <= 0 else ds

        def step_device(st=None):
            if n:
                X.render_scene_device(scene, cams, res, out_dev, integration=integ, precision="fp32", ds=ds_v,
                                      stream=stream.cuda_stream, stats=st)

        def step_e2e(host_out):
            if n:
                X.render_scene(scene, cams, res, integration=integ, precision="fp32", ds=ds_v, out=host_out)

        h2d_step = len(scene.program_bytes()) + n * ctypes.sizeof(X._lib.XRayCameraParams64) + 12 * int(3.48 / ds_v + 2)
        e2e_api = "XRayRenderSceneCUDA"

    # ---- kernel-resident arm: inputs (program / volume) already in HBM, output stays in HBM ----
    # work counters come from ONE untimed pass of the counting kernel variants; the timed passes run the production variants
    step_device(stats)
    st1 = {k: int(stats[i]) for i, k in enumerate(STAT_KEYS)}
    span_ran = bool(int(stats[7]) & 0x10000)
    for _ in range(warmup):
        step_device()
    barrier()
    if sampler is not None:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(steps):
        flush.fill_(k & 0xFF)  # evict the previous step's lines from L2 (not timed)
        evs[k][0].record(stream)
        step_device()
        evs[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler is not None else None
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)

    # ---- end-to-end arm: public host API, HOST buffers, H2D + D2H inside the timed region ----
    def time_e2e(host_out):
        for _ in range(min(warmup, 2)):
            step_e2e(host_out)
        bcast_ms.clear()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e(host_out)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    e2e_page_s = time_e2e(out_page[:max(n, 1)])
    bcast_page = list(bcast_ms)
    e2e_pin_s = time_e2e(out_pin.numpy()[:max(n, 1)])

    # ---- parity spot check (rank 0): 256 random pixels of this rank's first view against the oracle ----
    spot = None
    if parity and rank == 0 and n:
        from oracle import oracle as O

        rng = np.random.default_rng(2024)
        ij = np.stack([rng.integers(0, res, 256), rng.integers(0, res, 256)], axis=1).astype(np.int32)
        if is_volume:
            osc = O.OracleScene({"type": "voxel_grid", "_array_f32": vol_np})
        else:
            osc = O.OracleScene(str(SCENES / obj), str(SCENES / deform) if deform else None)
        c0 = cams[0]
        want, _ = osc.render_pixels(np.array(list(c0.eye)), np.array(list(c0.view)).reshape(4, 4), res, float(c0.fov_y), float(c0.R),
                                    ds_v, ij, integ)
        got = out_page[0][ij[:, 0], ij[:, 1]].astype(np.float64)
        spot = {"max_abs_dI": float(np.abs(got - want).max()), "pixels": 256, "view": [float(angles[0]["azimuthal"]), float(angles[0]["polar"])],
                "tolerance": 1e-4, "against": "oracle (C++ restatement of the Go CPU path), same camera and step",
                "attenuated_pixels": int((want < 0.999).sum())}

    # ---- reduce over ranks: max time, sum of work ----
    vec = torch.tensor([dev_ms, e2e_page_s * 1e3, e2e_pin_s * 1e3, float(np.mean(bcast_page)) if bcast_page else 0.0], dtype=torch.float64, device=dev)
    work = torch.tensor([st1[k] for k in STAT_KEYS] + [float(h2d_step), float(n * res * res * 4)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_page_ms, e2e_pin_ms, bcast_mean = (float(x) for x in vec.tolist())
    w = dict(zip(STAT_KEYS + ("h2d", "d2h"), (float(x) for x in work.tolist())))
    if rank != 0:
        return None

    secs = dev_ms_max / 1e3 / steps            # per step, max over ranks
    ref_samples, rays = w["ref_samples"], w["rays"]   # per step, all ranks
    rec = {
        "workload": workload_string(name, volume_n), "views_rendered": views_total, "views_in_config": views_full,
        "value": ref_samples / secs / 1e9, "unit": "Gsamples/s", "ms_per_step": dev_ms_max / steps,
        "projections_512_per_s": rays / secs / 262144.0, "rays_per_s": rays / secs,
        "e2e": {"value": ref_samples / (e2e_page_ms / 1e3 / steps) / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": int(w["h2d"]),
                "d2h_bytes_per_step": int(w["d2h"]), "ms_per_step": e2e_page_ms / steps, "host_buffers": "pageable (numpy), as a cgo / ctypes caller passes",
                "pinned": {"value": ref_samples / (e2e_pin_ms / 1e3 / steps) / 1e9, "ms_per_step": e2e_pin_ms / steps},
                "projections_512_per_s": rays / (e2e_page_ms / 1e3 / steps) / 262144.0, "api": e2e_api},
        "gpu_launches": int(w["launches"]) * steps, "launches_per_step": int(w["launches"]),
        "parity_spot_check": spot, "wall_s_timed_region": t_wall, "clocks": clocks,
        "work_per_step": {"ref_samples": ref_samples, "rays": rays, "intervals_or_evaluated_samples": w["evaluated_samples"],
                          "candidate_or_primitive_tests": w["primitive_tests"], "fp64_fallbacks": w["fp64_fallbacks"],
                          "warp_tiles_handed_to_marching_kernels": w["marched_tiles"]},
    }
    if world > 1 and is_volume:
        # on the root the interval starts when its own upload is done; the other ranks' intervals include waiting for that upload
        rec["e2e"]["broadcast_ms"] = float(np.mean(bcast_page)) if bcast_page else None
        rec["e2e"]["broadcast_ms_max_over_ranks_incl_wait_for_root_upload"] = bcast_mean
        rec["e2e"]["broadcast_gb_per_s"] = (4.0 * volume_n ** 3 / 1e9) / (rec["e2e"]["broadcast_ms"] / 1e3) if bcast_page else None
    # ---- roofline of the dominant kernel ----
    launches = max(1.0, w["launches"] / world)   # per rank and step
    peak_tflops = X.measure_fp32_peak()
    if is_volume:
        # voxel kernel: bound by the memory hierarchy.  Algorithmic (compulsory) bytes per view: the volume read once
        # (4 N^3) plus the image written once (4 res^2) -- SURVEY 8(d); peak = measured HBM copy bandwidth.
        peaks = {}
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        except (OSError, ValueError):
            pass
        hbm = float(peaks.get("hbm_gbs", 6491.2))
        bytes_view = 4.0 * volume_n ** 3 + 4.0 * res * res
        ach = bytes_view * views_total / world / secs / 1e9   # per GPU
        rec["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "B200_PROFILING.md fallback",
                           "algorithmic_bytes_per_launch": bytes_view * views_total / world / launches,
                           "avg_launch_ms": dev_ms_max / steps / launches,
                           "note": "compulsory bytes: volume once + image once per view; the trilinear tap stream itself is served on chip"}
    else:
        if span_ran:
            flops = w["primitive_tests"] * FLOPS_SPAN_CANDIDATE + rays * FLOPS_SPAN_RAY
            model = {"per_exact_interval_candidate": FLOPS_SPAN_CANDIDATE, "per_ray": FLOPS_SPAN_RAY, "kernel": "render_span_kernel"}
        else:
            per_sample, per_prim, _ = scene_flops_model(scene.object_map, scene.deformation_map)
            flops = w["evaluated_samples"] * per_sample + w["primitive_tests"] * per_prim
            model = {"per_evaluated_sample": per_sample, "per_primitive_test": per_prim, "kernel": "marching kernels"}
        ach = flops / world / secs / 1e12
        rec["roofline"] = {"bound": "fp32", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops if peak_tflops else None,
                           "traffic": None, "peak_source": "measured in this run: FFMA-chain microbenchmark (XRayMeasureFp32Peak); MEASURED_PEAKS.json has no FP32 figure",
                           "algorithmic_flops_per_launch": flops / world / launches, "avg_launch_ms": dev_ms_max / steps / launches,
                           "flops_model": model,
                           "note": "the executed algorithm removes the per-sample flops the reference spends (intervals instead of samples), so the "
                                   "fraction of the FP32 peak is small by construction; the binding resource is warp-instruction issue (see issue)"}
    if ncu_info and name in ncu_info:
        tr = ncu_info[name]
        vps = views_total / world   # views per rank and step
        rec["roofline"]["traffic"] = tr["dram_bytes_per_view"] * vps / launches
        rec["roofline"]["traffic_source"] = tr["capture"]
        if clocks and clocks.get("sm_mhz") and tr.get("warp_inst_per_view"):
            inst_s = tr["warp_inst_per_view"] * vps / secs
            peak_inst = 4.0 * 148 * clocks["sm_mhz"] * 1e6
            rec["roofline"]["issue"] = {"warp_inst_per_view": tr["warp_inst_per_view"], "warp_inst_per_ray": tr["warp_inst_per_view"] / (res * res),
                                        "achieved_ginst_per_s": inst_s / 1e9, "peak_ginst_per_s": peak_inst / 1e9, "frac": inst_s / peak_inst,
                                        "note": "instruction count from the committed ncu capture of this kernel, time from this run"}
    if is_volume and world == 1 and ref_cuda:
        rec["reference_cuda"] = time_reference_cuda_plugin(X, vol_np, cams, res, ds_v, ref_samples)
    return rec


def go_output_pin(X):
    """Not timed.  The one image the reference repository holds that its Go binary rendered (the cell output stored in
    examples/demo.ipynb: three 300 x 300 views of cube_w_hole; fixture under tests/golden/, see
    tests/test_reference_go_output.py) rendered again through the C ABI and compared as 8-bit grey levels."""
    try:
        stored = np.load(ROOT / "tests" / "golden" / "reference_go_cube_w_hole_3x300.npz")["image"]
        sc = X.Scene(str(SCENES / "cube_w_hole.json"))
        cams = X.cameras_from_angles(X.generate_camera_angles(3), 5.0, 45.0)
        out = {}
        for prec in ("fp64", "fp32"):
            imgs = X.render_scene(sc, cams, 300, precision=prec, ds=0.1, integration="hierarchical")
            grey = np.hstack([X.image_to_rgba8(np.asarray(v, dtype=np.float64))[..., 0] for v in imgs])
            d = np.abs(grey.astype(int) - stored.astype(int))
            out[prec] = {"pixels_differing": int((d != 0).sum()), "max_grey_level_difference": int(d.max())}
        return {"against": "output of the reference's Go binary: image stored in examples/demo.ipynb (3 views of cube_w_hole at 300x300; R=5, fov=45, "
                           "ds=0.1, hierarchical reproduce it exactly on the CPU oracle)",
                "pixels": int(stored.size), "attenuated_pixels": int((stored < 255).sum()), **out}
    except Exception as e:  # the pin is evidence beside the measurement, never a reason to lose the line
        return {"error": f"{type(e).__name__}: {e}"}


SUB_CONFIGS = (  # (key, workload, views rendered on N = 1)
    ("cfg1", "cube_w_hole", 1), ("cfg3", "gyroid_sigmoid", 90), ("cfg4", "voxel1024", 16), ("cfg5", "pillar_array", 8))


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch

    import xray_projection_render_b200 as X

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    try:
        ncu_info = json.load(open(ROOT / "profiles" / "ncu_dram_traffic_r2.json"))
    except (OSError, ValueError):
        ncu_info = None

    sampler = ClockSampler(local_rank) if rank == 0 else None
    is_volume = WORKLOADS[args.workload][0] is None
    volume = synthetic_volume(args.volume_n) if (rank == 0 and (is_volume or not args.no_configs)) else None
    main = measure(X, torch, dist, args.workload, rank, world, local_rank, args.steps, args.warmup, flush, views_total=args.views,
                   volume_n=args.volume_n, volume=volume if is_volume else None, sampler=sampler, ncu_info=ncu_info,
                   ref_cuda=not args.no_ref_cuda)
    subs = {}
    if not args.no_configs:
        for key, name, nv in SUB_CONFIGS:
            if name == args.workload or (world > 1 and key == "cfg1"):
                continue
            nv_total = nv * world if world > 1 else nv   # every rank keeps the same bounded share of the config's view list
            r = measure(X, torch, dist, name, rank, world, local_rank, 2, 3, flush, views_total=nv_total, volume_n=args.volume_n,
                        volume=volume if name == "voxel1024" else None, ncu_info=ncu_info, ref_cuda=(world == 1 and not args.no_ref_cuda))
            if rank == 0:
                subs[key] = r
    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return
    obj, deform, views, res, integ, ds = WORKLOADS[args.workload]
    line = {
        "metric": "Gsamples/s", "value": main["value"], "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if is_volume else ("f64" if main["roofline"].get("flops_model", {}).get("kernel") == "render_span_kernel" else "f32"),
        "data": "synthetic",
        "projections_512_per_s": main["projections_512_per_s"], "rays_per_s": main["rays_per_s"],
        "config": {"workload": main["workload"], "views_total": main["views_rendered"],
                   "sharding": "views modulo rank (main.go:244 --jobs_modulo/--job); fixed view list, so N ranks share the same work",
                   "mode": "fp32-mode request (|dI| <= 1e-4); scenes of convex primitives take the interval renderer, whose intervals are fp64",
                   "l2": "256 MiB buffer rewritten between timed steps; the step's output (>= 1.5 GB on one GPU) exceeds the 126 MB L2"},
        "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "clocks": main["clocks"], "roofline": main["roofline"],
        "parity_spot_check": main["parity_spot_check"], "work_per_step": main["work_per_step"], "wall_s_timed_region": main["wall_s_timed_region"],
    }
    if "reference_cuda" in main:
        line["reference_cuda"] = main["reference_cuda"]
    if subs:
        for r in subs.values():
            r.pop("clocks", None)
        line["configs"] = subs
    if world == 1:
        line["reference_output_pin"] = go_output_pin(X)
    try:
        line["library"] = X._lib.library_info()  # which libcuda_render.so ran, and when / with what it was compiled
    except Exception as e:  # evidence beside the measurement, never a reason to lose the line
        line["library"] = {"error": f"{type(e).__name__}: {e}"}
    if world == 1 and not args.no_cpu:
        cb = cpu_baseline(args.workload, budget_s=args.cpu_budget, volume=volume if is_volume else None)
        line["cpu_baseline"] = {"value": cb["gsamples_per_s"], "unit": "Gsamples/s", "cores": cb["cores"], "kind": cb["kind"],
                                "sample": cb["sample"], "projections_512_per_s": cb["rays_per_s"] / 262144.0}
        for key, name, _ in SUB_CONFIGS:
            if key in subs:
                c2 = cpu_baseline(name, budget_s=max(3.0, args.cpu_budget / 4), volume=volume if name == "voxel1024" else None)
                subs[key]["cpu_baseline"] = {"value": c2["gsamples_per_s"], "unit": "Gsamples/s", "cores": c2["cores"], "kind": c2["kind"],
                                             "sample": c2["sample"]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lattice", choices=sorted(WORKLOADS))
    ap.add_argument("--views", type=int, default=0, help="override the total number of views (diagnostics only)")
    ap.add_argument("--volume-n", type=int, default=1024)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline workload only (skip the cfg1/3/4/5 sub-records)")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference CUDA plugin (voxel workload)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", str(Path(__file__).resolve())] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
