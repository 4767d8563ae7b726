timeout 1700 python -m pytest tests/test_gpu_span.py -q -m gpu 2>&1 | tail -30
