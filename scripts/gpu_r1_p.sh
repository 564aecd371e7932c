NW=32 RN_PAIR_DEBUG=1 RN_TARGET_UNITS=16384 timeout 120 python scripts/pair_debug.py cfg3 2>&1 | tail -20
