mkdir -p gpurun_out
python -m pytest tests/test_pairwise_gpu.py tests/test_dropin_gpu.py tests/test_listwise_pairs_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/zf_tests.txt
for dbg in 0 128 0 128; do echo "== RN_PAIR_DEBUG=$dbg"; RN_PAIR_DEBUG=$dbg python scripts/quick_time.py cfg2 cfg3 2>&1 | grep -E "us/call|stamps" | sed -E 's/.*(B=[0-9]+ n_pair=[0-9]+ [0-9.]+ us\/call).*(loss=[0-9.]+).*/\1 \2/; s/.*(16:[0-9.]+) .*(20:[0-9.]+ 21:[0-9.]+ 22:[0-9.]+ 23:[0-9.]+)/   \1 \2/'; done | tee gpurun_out/zf_time.txt
