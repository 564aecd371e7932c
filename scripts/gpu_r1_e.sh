timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for bps in 1 2; do for tu in 8192 16384 32768; do echo "== BPS=$bps TARGET_UNITS=$tu"; RN_PAIR_BPS=$bps RN_TARGET_UNITS=$tu timeout 120 python scripts/quick_time.py cfg3 2>&1 | tail -2; done; done
echo "== cfg1/cfg2 default"; timeout 120 python scripts/quick_time.py cfg1 cfg2 2>&1 | tail -4
