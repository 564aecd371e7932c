mkdir -p gpurun_out
python -m pytest tests/test_pairwise_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/z_tests.txt
for plan in 64 24,40,48,56,60,64 32,48,56,60,64 16,32,40,48,52,56,60,64 32,64 16,32,48,64 8,16,24,32,40,48,56,64 40,56,60,62,64 20,36,48,56,60,62,63,64; do
  echo "== RN_PAIR_CHUNKS=$plan"
  RN_PAIR_CHUNKS=$plan python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps" | sed -E 's/.*(B=65536 n_pair=[0-9]+ [0-9.]+ us\/call).*/\1/; s/.*(16:[0-9.]+) .*(20:[0-9.]+ 21:[0-9.]+ 22:[0-9.]+ 23:[0-9.]+)/   \1 \2/'
done 2>&1 | tee gpurun_out/z_chunks.txt
RN_PAIR_DEBUG=1 NW=32 python scripts/pair_debug.py cfg3 2>&1 | tail -40 > gpurun_out/z_pair_debug.txt
