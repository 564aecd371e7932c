timeout 900 python -m pytest tests/test_pairwise_gpu.py -m gpu -x -q 2>&1 | tail -3
for tu in 4096 8192 16384 32768; do echo "== TARGET_UNITS=$tu"; RN_PAIR_DEBUG=1 RN_TARGET_UNITS=$tu timeout 120 python scripts/pair_debug.py cfg3 2>&1 | tail -11 | head -7;  RN_TARGET_UNITS=$tu timeout 120 python scripts/quick_time.py cfg3 2>&1 | tail -4 | head -2; done
