timeout 900 python -m pytest tests/test_pairwise_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -3
for tu in 8192 16384; do echo "== TARGET_UNITS=$tu"; RN_PAIR_DEBUG=1 RN_TARGET_UNITS=$tu timeout 120 python scripts/pair_debug.py cfg3 2>&1 | tail -11 | head -7;  RN_PAIR_DEBUG=1 RN_TARGET_UNITS=$tu timeout 120 python scripts/quick_time.py cfg3 2>&1 | tail -4; done
timeout 120 python scripts/quick_time.py cfg1 cfg2 2>&1 | grep -v stamps
