mkdir -p gpurun_out
run() { echo "== STRADDLE=$1 GEN=$2 SWITCH=$3"; RN_PAIR_COST_STRADDLE=$1 RN_PAIR_COST_GEN=$2 RN_PAIR_COST_SWITCH=$3 python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps" | sed -E 's/.*(B=65536 n_pair=[0-9]+ [0-9.]+ us\/call).*/\1/; s/.*(16:[0-9.]+) .*(20:[0-9.]+ 21:[0-9.]+ 22:[0-9.]+ 23:[0-9.]+)/   \1 \2/'; }
(
run 12 17 4
run 8 17 4
run 16 17 4
run 12 17 7
run 12 24 7
run 16 24 7
run 12 17 10
run 20 24 10
) 2>&1 | tee gpurun_out/cost_straddle2.txt
python -m pytest tests/test_pairwise_gpu.py -m gpu -x -q 2>&1 | tail -2
