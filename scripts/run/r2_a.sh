set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_counting_path_gpu.py -x -q 2>&1 | tail -25
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/quick_time.py cfg1 cfg2 cfg3 2>&1 | tail -20
RN_SEG_COUNT=0 timeout 300 python scripts/quick_time.py cfg3 2>&1 | tail -6
