timeout 900 python -m pytest tests/test_pairwise_gpu.py -m gpu -x -q 2>&1 | tail -2
python scripts/host_overhead.py 2>&1 | tail -8
