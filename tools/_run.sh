timeout 900 python -m pytest tests/test_gpu_volume.py -q -m gpu -x 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_properties.py tests/test_gpu_span.py -q -m gpu -x 2>&1 | tail -4
python tools/span_time.py lattice pillar 2>&1 | tail -2
