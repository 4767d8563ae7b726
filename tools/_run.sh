timeout 400 python tools/fuzz_soak.py 1600 5000 2>&1 | tail -8
timeout 300 python tools/span_soak.py 3300 9000 2>&1 | tail -8
timeout 300 python tools/wide_soak.py 1500 6000 2>&1 | tail -8
SOAK_NESTED=1 timeout 300 python tools/wide_soak.py 1200 5000 2>&1 | tail -8
SOAK_SIMPLE=1 timeout 200 python tools/vol_soak.py 4000 8000 2>&1 | tail -5
