for tu in 8192 32768; do echo "== TARGET_UNITS=$tu"; RN_PAIR_DEBUG=1 RN_TARGET_UNITS=$tu timeout 120 python scripts/quick_time.py cfg3 2>&1 | tail -4; done
