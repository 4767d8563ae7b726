python tools/span_reasons.py 2>&1 | grep -v simple | grep lattice | tee gpurun_out/span_reasons6.txt
timeout 1700 python -m pytest tests/test_gpu_span.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_properties.py -x -q -m gpu 2>&1 | tail -12
python tools/span_time.py lattice pillar cube 2>&1 | tee gpurun_out/span_time12.txt
