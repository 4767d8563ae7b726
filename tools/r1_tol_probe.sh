#!/bin/bash
# Development aid: how much would a tighter gyroid guard band buy?  (unsound settings, timing only)
mkdir -p gpurun_out
for sc in 1 0.5 0.25 0.1; do
  XRAY_DEBUG_GYROID_TOL_SCALE=$sc python bench.py --workload gyroid_sigmoid --views 8 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('tol x$sc', round(d['value'],1), 'fb', int(r['fp64_fallbacks']), 'eval', int(r['evaluated_samples']))"
done
