timeout 120 python scripts/quick_time.py cfg1 cfg2 cfg3 2>&1 | tail -8
