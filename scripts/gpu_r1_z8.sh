mkdir -p gpurun_out
run() { echo "== SWITCH=$1 GEN=$2"; RN_PAIR_COST_SWITCH=$1 RN_PAIR_COST_GEN=$2 python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps" | sed -E 's/.*(B=65536 n_pair=[0-9]+ [0-9.]+ us\/call).*/\1/; s/.*(16:[0-9.]+) .*(20:[0-9.]+ 21:[0-9.]+ 22:[0-9.]+ 23:[0-9.]+)/   \1 \2/'; }
(
for sw in 0 2 3 4 6; do for g in 17 21 25; do run $sw $g; done; done
) 2>&1 | tee gpurun_out/z8_cost.txt
