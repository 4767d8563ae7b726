#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider ) > gpurun_out/pytest_gpu.txt 2>&1
tail -12 gpurun_out/pytest_gpu.txt
python tools/integ_split.py 2>&1 | tail -8
for w in lattice lattice_linear lattice_sigmoid balls box_w_pped cube_w_hole; do
python bench.py --workload $w --views 24 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2))"
done
