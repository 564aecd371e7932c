timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for tu in 8192 16384 32768; do echo "== TARGET_UNITS=$tu"; RN_TARGET_UNITS=$tu timeout 120 python scripts/quick_time.py cfg3 2>&1 | tail -2; done
echo "== cfg1/cfg2 default"; timeout 120 python scripts/quick_time.py cfg1 cfg2 2>&1 | tail -4
