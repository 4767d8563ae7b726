timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mask_width or big" 2>&1 | tail -15
timeout 1500 python tools/fuzz_soak.py 150 1600 2>&1 | tail -30
