timeout 900 python -m pytest tests/test_gpu_span.py -q -m gpu -k "overflow_walk" 2>&1 | grep -v "^X =\|^O =\|^obj =" | tail -30
