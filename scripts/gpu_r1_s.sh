export RN_PAIR_DEBUG=1 RN_TARGET_UNITS=1
python scripts/lone_warp.py 64 2048 2>&1 | tail -1
python scripts/lone_warp.py 64 4096 2>&1 | tail -1
python scripts/lone_warp.py 2048 2048 2>&1 | tail -1
python scripts/lone_warp.py 8192 2048 2>&1 | tail -1
