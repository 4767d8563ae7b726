"""N>1 host logic on CPU: two gloo ranks shard the view list exactly as bench.py / the library do
(view v -> rank v mod N, the reference's --jobs_modulo/--job semantics, main.go:244), reduce their
timing with MAX and their work counters with SUM, and reassemble images in view order."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_views, out_q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import bench
    from oracle import oracle as O

    dist.init_process_group("gloo", rank=rank, world_size=world)
    angles = bench.rank_views(total_views, rank, world)          # what each bench rank renders
    mine = list(range(rank, total_views, world))
    assert [a["azimuthal"] for a in angles] == [90.0 + v * 360.0 / total_views for v in mine]
    # stand-in for the GPU work: a tiny oracle render per view (deterministic function of the angle)
    osc = O.OracleScene(str(ROOT / "tests" / "scenes" / "balls.json"))
    imgs = []
    n_samples = 0
    for a in angles:
        eye, cm = O.camera_from_angles(a["azimuthal"], a["polar"], 4.0)
        im, n = osc.render_view(eye, cm, 8, 40.0, 4.0, 0.03, "hierarchical", nthreads=1)
        imgs.append(im)
        n_samples += n
    t = torch.tensor([10.0 + rank], dtype=torch.float64)          # pretend timing: max over ranks
    w = torch.tensor([float(n_samples), float(len(angles) * 64)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    local = torch.from_numpy(np.stack(imgs))
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, local.numpy()))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        full = np.zeros((total_views, 8, 8))
        seen = []
        for views, arr in gathered:
            for k, v in enumerate(views):
                full[v] = arr[k]
                seen.append(v)
        out_q.put((float(t[0]), w.tolist(), sorted(seen), full))


@pytest.mark.timeout(300)
def test_two_rank_view_sharding_gloo():
    import torch.multiprocessing as mp

    from oracle import oracle as O

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    total = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, work, seen, full = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 11.0                       # max over ranks
    assert seen == list(range(total))         # every view rendered exactly once
    osc = O.OracleScene(str(ROOT / "tests" / "scenes" / "balls.json"))
    n_total = 0
    for v, (az, pol) in enumerate(O.generate_camera_angles(total)):
        eye, cm = O.camera_from_angles(az, pol, 4.0)
        ref, n = osc.render_view(eye, cm, 8, 40.0, 4.0, 0.03, "hierarchical", nthreads=1)
        n_total += n
        assert np.array_equal(full[v], ref)
    assert work == [float(n_total), float(total * 64)]   # summed work = single-rank work


def test_bench_workloads_cover_baseline_configs():
    sys.path.insert(0, str(ROOT))
    import bench

    w = bench.WORKLOADS
    assert w["cube_w_hole"][2:5] == (1, 512, "hierarchical")
    assert w["lattice"][2:5] == (360, 1024, "hierarchical")
    assert w["gyroid_sigmoid"][:5] == ("gyroid_example.json", "deformation_sigmoid.json", 720, 1024, "hierarchical")
    assert w["voxel1024"][2:5] == (1440, 2048, "simple")
    assert w["pillar_array"][2:5] == (2880, 4096, "hierarchical")
    per_sample, per_prim, brute = bench.scene_flops_model({"type": "object_collection", "objects": [
        {"type": "sphere"}, {"type": "cylinder"}]}, None)
    assert per_sample == 6 + 2 and per_prim == ((9 + 1) + (27 + 1)) / 2 and brute == (9 + 1) + (27 + 1)


def test_synthetic_volume_definition():
    """BASELINE.md cfg4 volume: clamp(0.5+0.5 sin(6 pi x) sin(10 pi y) sin(14 pi z),0,1) inside r<0.9, [z][x][y]."""
    sys.path.insert(0, str(ROOT))
    import bench

    n = 24
    vol = bench.synthetic_volume(n)
    assert vol.shape == (n, n, n) and vol.dtype == np.float32
    ax = 2.0 * np.arange(n) / (n - 1) - 1.0
    k, i, j = 5, 11, 17
    x, y, z = ax[i], ax[j], ax[k]
    want = min(max(0.5 + 0.5 * np.sin(6 * np.pi * x) * np.sin(10 * np.pi * y) * np.sin(14 * np.pi * z), 0.0), 1.0)
    want *= float(x * x + y * y + z * z < 0.81)
    assert abs(vol[k, i, j] - want) <= 1e-6
    assert vol[0, 0, 0] == 0.0 and vol.max() <= 1.0


def test_reference_arm_line_and_shared_workload_string(capsys):
    """bench.py --impl reference: one JSON line with the contract's keys, the SAME config.workload string our arm prints
    (bench.workload_string), at least 3 s of CPU work per step unless the budget says otherwise, step count capped at 5."""
    sys.path.insert(0, str(ROOT))
    import json
    import types

    import bench

    for name in bench.WORKLOADS:
        s = bench.workload_string(name)
        assert s.startswith(name + ": ") and f"{bench.WORKLOADS[name][2]} views at {bench.WORKLOADS[name][3]}x{bench.WORKLOADS[name][3]}" in s
    cb = bench.cpu_baseline("cube_w_hole", budget_s=0.3)
    assert cb["gsamples_per_s"] > 0 and cb["kind"] == "port" and cb["cores"] >= 1 and "central rows" in cb["sample"]
    orig = bench.cpu_baseline
    calls = []

    def fake(workload, budget_s=15.0, nthreads=0, volume=None):
        calls.append(budget_s)
        return {"gsamples_per_s": 0.5, "rays_per_s": 1000.0, "cores": 4, "kind": "port", "sample": "fake", "seconds": budget_s}

    bench.cpu_baseline = fake
    try:
        bench.run_reference(types.SimpleNamespace(steps=20, warmup=3, cpu_budget=15.0, workload="lattice", gpus=1, volume_n=1024), 0, 1)
    finally:
        bench.cpu_baseline = orig
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 5 and line["steps_requested"] == 20
    assert line["config"]["workload"] == bench.workload_string("lattice")
    assert line["e2e"] == {"value": line["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert all(b >= 3.0 for b in calls[1:])   # every timed step is a >= 3 s sample (calls[0] is the warm-up)
    assert line["cpu_baseline"]["kind"] == "port" and line["metric"] == "Gsamples/s" and line["higher_is_better"] is True
