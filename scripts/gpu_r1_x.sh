timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/quick_time.py cfg1 cfg2 cfg3 2>&1 | grep -E "us/call|stamps"
RN_GRAPH=0 python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps"
python scripts/host_overhead.py 2>&1 | head -2
RN_GRAPH=0 python scripts/host_overhead.py 2>&1 | head -2
