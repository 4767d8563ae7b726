timeout 1500 python -m pytest tests/test_gpu_span.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_properties.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu_r2c.txt; cat gpurun_out/pytest_gpu_r2c.txt
python tools/span_reasons.py 2>&1 | grep -v simple | tee gpurun_out/span_reasons2.txt
python tools/span_time.py lattice pillar cube box_w_pped lattice_linear 2>&1 | tee gpurun_out/span_time5.txt
for v in A B; do XRAY_CUDA_LIB=$PWD/xray_projection_render_b200/lib_dev$v/libcuda_render.so python tools/span_time.py lattice pillar 2>&1 | tee -a gpurun_out/span_time5.txt; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_span5.csv python tools/render_once.py lattice.json 1024 16 - 2 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_span5.csv | cut -d, -f5,12- | tail -6
