mkdir -p gpurun_out
for plan in "64 0" "32,64 1" "32,48,56,60,64 2"; do set -- $plan
echo "== CHUNKS=$1 SNAP=$2"
RN_PAIR_CHUNKS=$1 RN_PAIR_SNAP=$2 RN_LIB_PATH=$PWD/scripts/dev/_variants/lib_trace.so RN_PAIR_DEBUG=1 python scripts/pair_trace_dump.py cfg3 gpurun_out/z7_trace_$2.npz 2>&1 | tail -16
done | tee gpurun_out/z7_trace.txt
RN_PAIR_CHUNKS=24,40,48,56,60,64 RN_PAIR_SNAP=2 python -m pytest tests/test_pairwise_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/z7_tests_chunks.txt
