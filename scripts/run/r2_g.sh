set -x
for i in 1 2 3; do timeout 600 python -m pytest tests/test_counting_path_gpu.py -x -q -k "levels or mixed" 2>&1 | tail -2; done
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
