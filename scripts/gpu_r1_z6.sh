mkdir -p gpurun_out
python -m pytest tests/test_pairwise_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/z6_tests.txt
run() { echo "== CHUNKS=$1 SNAP=$2 GEN=$3"; RN_PAIR_CHUNKS=$1 RN_PAIR_SNAP=$2 RN_PAIR_COST_GEN=$3 python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps" | sed -E 's/.*(B=65536 n_pair=[0-9]+ [0-9.]+ us\/call).*/\1/; s/.*(16:[0-9.]+) .*(20:[0-9.]+ 21:[0-9.]+ 22:[0-9.]+ 23:[0-9.]+)/   \1 \2/'; }
(
run 64 0 17
run 64 0 20
run 32,64 0 17
run 32,64 1 17
run 32,48,64 0 17
run 32,48,64 2 17
run 32,48,56,64 0 17
run 32,48,56,64 3 17
run 32,48,56,60,64 2 17
run 24,40,48,56,60,64 2 17
run 40,56,64 2 17
run 48,64 1 17
run 48,56,64 1 17
run 48,56,60,64 1 17
) 2>&1 | tee gpurun_out/z6_chunks.txt
