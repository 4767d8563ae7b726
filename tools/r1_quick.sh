#!/bin/bash
# Development aid (1 GPU): GPU suite + the three analytic BASELINE workloads, short.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider ) > gpurun_out/pytest_gpu.txt 2>&1
tail -6 gpurun_out/pytest_gpu.txt
for w in gyroid_sigmoid pillar_array lattice; do
python bench.py --workload $w --views 24 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'fb', int(r['fp64_fallbacks']), 'eval', int(r['evaluated_samples']))"
done
python tools/integ_split.py 2>&1 | tail -6
