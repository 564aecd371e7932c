set -x
timeout 300 python scripts/quick_time.py cfg2 cfg3 2>&1 | tail -8
RN_SEG_DEBUG=1 timeout 300 python scripts/seg_debug.py cfg3 2>&1 | tail -12
RN_SEG_DEBUG=1 timeout 300 python scripts/seg_debug.py cfg2 2>&1 | tail -12
