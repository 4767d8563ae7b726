python tools/span_time.py lattice pillar cube box_w_pped balls 2>&1 | tail -5
python tools/span_reasons.py 2>&1 | grep -v "marched_tiles     0"
timeout 900 python -m pytest tests/test_gpu_span.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -q -m gpu -x 2>&1 | tail -3
