timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/host_overhead.py 2>&1 | tail -7
