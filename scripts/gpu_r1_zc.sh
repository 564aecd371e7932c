mkdir -p gpurun_out
python -m pytest tests/test_host_pipeline_gpu.py tests/test_pairwise_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/zc_tests.txt
for dbg in 0 32 64 96; do echo "== RN_PAIR_DEBUG=$dbg"; RN_PAIR_DEBUG=$dbg python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps" | sed -E 's/.*(B=[0-9]+ n_pair=[0-9]+ [0-9.]+ us\/call).*(loss=[0-9.]+).*/\1 \2/; s/.*(16:[0-9.]+) .*(20:[0-9.]+ 21:[0-9.]+ 22:[0-9.]+ 23:[0-9.]+)/   \1 \2/'; done | tee gpurun_out/zc_time.txt
python scripts/host_pipeline_probe.py 2>&1 | tail -4 | tee gpurun_out/zc_host.txt
