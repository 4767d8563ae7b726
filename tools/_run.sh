timeout 600 python tools/vol_time.py 512 1024 4 2>&1 | grep "fp32"
timeout 600 python tools/vol_time.py 1024 2048 4 2>&1 | grep "simple       fp32"
