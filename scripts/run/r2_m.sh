set -x
timeout 900 python -m pytest tests/test_deterministic_gpu.py -x -q 2>&1 | tail -3
timeout 300 python scripts/quick_time.py cfg2 cfg3 2>&1 | tail -8
timeout 300 python scripts/quick_time.py cfg3 2>&1 | tail -4
