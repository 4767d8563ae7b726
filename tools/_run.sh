python tools/span_time.py lattice pillar cube box_w_pped balls 2>&1 | tail -5
