mkdir -p gpurun_out
run() { echo "== SWITCH=$1 GEN=$2"; RN_PAIR_COST_SWITCH=$1 RN_PAIR_COST_GEN=$2 python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps" | sed -E 's/.*(B=65536 n_pair=[0-9]+ [0-9.]+ us\/call).*/\1/; s/.*(16:[0-9.]+) .*(20:[0-9.]+ 21:[0-9.]+ 22:[0-9.]+ 23:[0-9.]+)/   \1 \2/'; }
(
for sw in 4 8; do for g in 17 24 30 36; do run $sw $g; done; done
) 2>&1 | tee gpurun_out/za_cost.txt
python bench.py --steps 200 --warmup 20 --no-cpu 2> gpurun_out/za_bench_err.log | tee gpurun_out/za_bench_n1.json
