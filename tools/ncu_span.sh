#!/bin/bash
# Development aid: full ncu capture of the span kernel on one scene.  usage: tools/ncu_span.sh <scene.json> <res> <views> <tag> [kernel-regex]
obj=$1; res=$2; nv=$3; tag=$4; rx=${5:-render_span_kernel}; deform=${6:--}
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$rx" -s 1 -c 1 -f \
    -o gpurun_out/prof_$tag python tools/render_once.py $obj $res $nv $deform > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/src_$tag.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/cuda_$tag.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/raw_$tag.csv gpurun_out/src_$tag.csv > gpurun_out/summary_$tag.txt 2>&1
python tools/ncu_lines.py gpurun_out/cuda_$tag.csv >> gpurun_out/summary_$tag.txt 2>&1
tail -3 gpurun_out/ncu_$tag.log
