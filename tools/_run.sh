timeout 1500 python tools/vol_soak.py 24 4000 2>&1 | tail -20
