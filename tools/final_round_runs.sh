#!/bin/bash
# Round-end evidence run (1 GPU): tests, smoke, ncu captures of the kernels that changed, every bench workload, the
# reference arm, the ncu launch lists.  Results land in gpurun_out/; copy what is to be judged into profiles/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -2 | tee gpurun_out/pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_final.txt
# one-primitive workloads run the lane-asynchronous kernel: capture it, refresh the per-view figures bench.py reads
RX='render_async_kernel<\(int\)[12], \(int\)[01], \(bool\)0'
tools/ncu_capture.sh gyroid_sigmoid r1i_gyroid_async "$RX"
tools/ncu_capture.sh pillar_array r1i_pillar_async "$RX"
python tools/update_ncu_json.py gyroid_sigmoid gpurun_out/raw_r1i_gyroid_async.csv ncu_r1i_gyroid_async_summary.txt
python tools/update_ncu_json.py pillar_array gpurun_out/raw_r1i_pillar_async.csv ncu_r1i_pillar_async_summary.txt
cp profiles/ncu_dram_traffic_r1.json gpurun_out/ncu_dram_traffic_r1.json
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1_final_lattice.json
python bench.py --workload gyroid_sigmoid --views 8 --steps 3 --warmup 3 --cpu-budget 10 2>&1 | tail -1 > gpurun_out/bench_r1_gyroid_sigmoid.json
python bench.py --workload pillar_array --views 8 --steps 3 --warmup 3 --cpu-budget 10 2>&1 | tail -1 > gpurun_out/bench_r1_pillar_array.json
python bench.py --workload cube_w_hole --views 1 --res 512 --steps 10 --warmup 3 --cpu-budget 5 2>&1 | tail -1 > gpurun_out/bench_r1_cube_w_hole.json
python bench.py --workload voxel1024 --views 8 --steps 3 --warmup 3 --cpu-budget 10 2>&1 | tail -1 > gpurun_out/bench_r1_voxel1024.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r1_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_gyroid.csv python bench.py --workload gyroid_sigmoid --views 8 --steps 2 --warmup 3 --no-cpu > gpurun_out/b3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_voxel.csv python bench.py --workload voxel1024 --views 4 --steps 2 --warmup 3 --no-cpu --no-ref-cuda > gpurun_out/b2.log 2>&1
for f in gpurun_out/bench_r1_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d.get("roofline",{}); c=d.get("cpu_baseline",{})
print(sys.argv[1].split("/")[-1], round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms", round(d["ms_per_step"],2), "frac", r.get("frac"), "issue", (r.get("issue") or {}).get("frac"),
      "brute", (r.get("survey_8d_brute_force") or {}).get("frac"), "cpu", c.get("value"), d.get("clocks",{}).get("reasons") if d.get("clocks") else None)
PY
done
