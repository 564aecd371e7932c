set -x
timeout 900 python -m pytest tests/test_focal_gpu.py -x -q 2>&1 | tail -12
timeout 300 python scripts/quick_time.py cfg3 2>&1 | tail -4
