#!/bin/bash
# Development aid: voxel bench for each warp pixel footprint (XRAY_VOLUME_TILE).
for t in ${1:-0 1 2 3 4 5}; do
  XRAY_VOLUME_TILE=$t python bench.py --workload voxel1024 --views 4 --steps 3 --warmup 3 --no-cpu --no-ref-cuda ${@:2} 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tile $t', round(d['value'],1), round(d['ms_per_step'],2))"
done
