nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/tma_probe tools/experiments/tma_brick_probe.cu && /tmp/tma_probe | tee gpurun_out/tma_brick_probe_r2.txt
cuobjdump -sass /tmp/tma_probe | grep -c UTMALDG | sed 's/^/UTMALDG instructions in the probe SASS: /' | tee -a gpurun_out/tma_brick_probe_r2.txt
python bench.py --workload voxel1024 --views 8 --no-configs --no-cpu --no-ref-cuda --steps 3 --warmup 3 > gpurun_out/bench_vox_skip.json 2>/dev/null
XRAY_VOLUME_NO_SKIP=1 python bench.py --workload voxel1024 --views 8 --no-configs --no-cpu --no-ref-cuda --steps 3 --warmup 3 > gpurun_out/bench_vox_noskip.json 2>/dev/null
python -c "
import json
for f in ('skip','noskip'):
    d=json.load(open('gpurun_out/bench_vox_%s.json'%f)); print(f, 'value', d['value'], 'ms/step', d['ms_per_step'], 'evaluated', d['work_per_step']['intervals_or_evaluated_samples'], 'ref', d['work_per_step']['ref_samples'])
" | tee -a gpurun_out/tma_brick_probe_r2.txt
