python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -3 gpurun_out/bench_r2a.err; python -c "
import json
d=json.load(open('gpurun_out/bench_r2a.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'pinned',d['e2e']['pinned'])
print('roof',d['roofline']['frac'],d['parity_spot_check'])
print('cpu',d.get('cpu_baseline'))
for k,v in d.get('configs',{}).items():
    print(k,'value',round(v['value'],1),'ms',round(v['ms_per_step'],3),'e2e',round(v['e2e']['value'],1),round(v['e2e']['ms_per_step'],2),'pin',round(v['e2e']['pinned']['value'],1),'spot',v['parity_spot_check']['max_abs_dI'],'roof',v['roofline']['frac'], v.get('cpu_baseline',{}).get('value'), v.get('reference_cuda',{}).get('speedup'))
"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2a_ref.json 2>&1; cat gpurun_out/bench_r2a_ref.json | cut -c1-600
