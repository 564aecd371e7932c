set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_pairwise_gpu.py -x -q 2>&1 | tail -40
