python tools/span_time.py lattice pillar cube lattice_linear 2>&1 | tail -4
python tools/span_reasons.py 2>&1 | grep -v "marched_tiles     0"
timeout 900 python -m pytest tests/test_gpu_span.py -q -m gpu -x 2>&1 | grep -v "^X =\|^O =\|^obj =" | tail -30
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_properties.py -x -q -m gpu 2>&1 | tail -6
