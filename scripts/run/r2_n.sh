set -x
timeout 900 python -m pytest tests/test_global_nccl_gpu.py tests/test_counting_path_gpu.py tests/test_deterministic_gpu.py -x -q 2>&1 | tail -4
timeout 300 python scripts/quick_time.py cfg1 cfg2 cfg3 2>&1 | tail -12
