#!/bin/bash
# Development aid: one full ncu capture of the timed (non-counting) hot kernel of a bench workload.
# usage: tools/ncu_capture.sh <workload> <tag> [kernel-regex] [extra bench.py args]
w=$1; tag=$2; rx=${3:-'render_fast_kernel<\(int\)[12], \(int\)[01], \(bool\)0'}
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$rx" -c 1 -f \
    -o gpurun_out/prof_$tag python bench.py --workload $w --views 2 --steps 1 --warmup 1 --no-cpu --no-ref-cuda $4 > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/src_$tag.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/raw_$tag.csv gpurun_out/src_$tag.csv > gpurun_out/summary_$tag.txt 2>&1
tail -5 gpurun_out/ncu_$tag.log
