mkdir -p gpurun_out
for v in base floop gen1 small; do
for plan in 64 32,64 24,40,48,56,60,64; do
  echo "== variant=$v RN_PAIR_CHUNKS=$plan"
  RN_LIB_PATH=$PWD/scripts/dev/_variants/lib_$v.so RN_PAIR_CHUNKS=$plan python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps" | sed -E 's/.*(B=65536 n_pair=[0-9]+ [0-9.]+ us\/call).*/\1/; s/.*(16:[0-9.]+) .*(20:[0-9.]+ 21:[0-9.]+ 22:[0-9.]+ 23:[0-9.]+)/   \1 \2/'
done; done 2>&1 | tee gpurun_out/z2_variants.txt
RN_LIB_PATH=$PWD/scripts/dev/_variants/lib_small.so python -m pytest tests/test_pairwise_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/z2_tests.txt
RN_LIB_PATH=$PWD/scripts/dev/_variants/lib_small.so RN_PAIR_CHUNKS=24,40,48,56,60,64 RN_PAIR_DEBUG=1 NW=32 python scripts/pair_debug.py cfg3 2>&1 | tail -40 > gpurun_out/z2_pair_debug_small_chunks.txt
cp gpurun_out/pair_debug.npz gpurun_out/z2_pair_debug_small_chunks.npz
RN_LIB_PATH=$PWD/scripts/dev/_variants/lib_small.so RN_PAIR_DEBUG=1 NW=32 python scripts/pair_debug.py cfg3 2>&1 | tail -40 > gpurun_out/z2_pair_debug_small.txt
