timeout 1200 python tools/span_soak.py 300 3300 2>&1 | tail -30
