timeout 1500 python -m pytest tests/test_gpu_span.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_properties.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu_r2d.txt; tail -8 gpurun_out/pytest_gpu_r2d.txt
python tools/span_reasons.py 2>&1 | grep -v simple | tee gpurun_out/span_reasons3.txt
python tools/span_time.py lattice pillar cube box_w_pped balls lattice_linear 2>&1 | tee gpurun_out/span_time8.txt
XRAY_SPAN_NO_BINS=1 python tools/span_time.py lattice pillar cube box_w_pped balls 2>&1 | tee -a gpurun_out/span_time8.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_span8.csv python tools/render_once.py lattice.json 1024 16 - 2 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_span8.csv | cut -d, -f5,12- | tail -8
