#!/bin/bash
# Development aid: A/B two builds of the library on the same box.  usage: tools/ab.sh "lattice gyroid_sigmoid" libA.so libB.so
for rep in 1; do
for w in $1; do
  for lib in "${@:2}"; do
    XRAY_CUDA_LIB=$PWD/$lib python bench.py --workload $w --views 8 --steps 5 --warmup 3 --no-cpu --no-ref-cuda 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['config']['workload'][:12], round(d['value'],1), round(d['ms_per_step'],2))"
  done
done
done
