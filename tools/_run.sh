timeout 600 python -m pytest tests/test_gpu_volume.py tests/test_gpu_properties.py -x -q -m gpu 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload voxel1024 --views 64 --steps 3 --warmup 3 --no-configs --no-cpu > gpurun_out/bench_r2f_n2_vox.json 2> gpurun_out/bench_r2f_n2_vox.err; tail -2 gpurun_out/bench_r2f_n2_vox.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2f_n2_vox.json').read().strip().splitlines()[-1])
print('vox N',d['n_gpus'],'views',d['config']['views_total'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2),'pinned',d['e2e']['pinned'], d['e2e'].get('broadcast_ms'), d['parity_spot_check']['max_abs_dI'])
"
