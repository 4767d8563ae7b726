#!/bin/bash
mkdir -p gpurun_out
python tools/integ_split.py 2>&1 | tail -8
( compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4 ) | tee gpurun_out/sanitizer_memcheck.txt
( compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "gyroid_cell or one_primitive or async or unbounded or face_planes" 2>&1 | tail -4 ) | tee gpurun_out/sanitizer_racecheck.txt
