set -x
timeout 900 python -m pytest tests/test_global_nccl_gpu.py -x -q 2>&1 | tail -12
timeout 600 python -m pytest tests/test_pairwise_gpu.py tests/test_counting_path_gpu.py -x -q 2>&1 | tail -4
