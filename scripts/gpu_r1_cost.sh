mkdir -p gpurun_out
run() { echo "== STRADDLE=$1 LEVELS=$2"; RN_PAIR_COST_STRADDLE=$1 RN_PAIR_COST_LEVELS=$2 python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps" | sed -E 's/.*(B=65536 n_pair=[0-9]+ [0-9.]+ us\/call).*/\1/; s/.*(16:[0-9.]+) .*(20:[0-9.]+ 21:[0-9.]+ 22:[0-9.]+ 23:[0-9.]+)/   \1 \2/'; }
(
run 12 4
run 12 3
run 16 3
run 20 2
run 20 3
) 2>&1 | tee gpurun_out/cost_levels.txt
