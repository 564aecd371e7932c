for cg in 13 17 21 26; do echo "== COST_GEN=$cg"; RN_PAIR_COST_GEN=$cg python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps"; done
RN_PAIR_DEBUG=1 NW=32 python scripts/pair_debug.py cfg3 2>&1 | grep -E "loop exit|per-SM last|eighths/warp|gen==0|gen>=3"
