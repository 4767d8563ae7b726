SOAK_SIMPLE=1 timeout 900 python tools/vol_soak.py 0 1200 2>&1 | tail -12
SOAK_NESTED=1 timeout 900 python tools/wide_soak.py 0 1200 2>&1 | tail -25
