for n in 8 4 2 1; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_r2d_n$n.json 2> gpurun_out/bench_r2d_n$n.err; tail -2 gpurun_out/bench_r2d_n$n.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2d_n$n.json').read().strip().splitlines()[-1])
print('N',d['n_gpus'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2),'pinned',d['e2e']['pinned'], 'marched', d['work_per_step']['warp_tiles_handed_to_marching_kernels'])
for k,v in d.get('configs',{}).items():
    print('  ',k,'views',v['views_rendered'],'value',round(v['value'],1),'ms',round(v['ms_per_step'],3),'e2e',round(v['e2e']['value'],1),round(v['e2e']['ms_per_step'],2),'pin',round(v['e2e']['pinned']['value'],1),'spot',v['parity_spot_check']['max_abs_dI'],'bcast',v['e2e'].get('broadcast_ms'), v['e2e'].get('broadcast_gb_per_s'))
"
done
nvidia-smi topo -m | head -12
