mkdir -p gpurun_out
RN_LIB_PATH=$PWD/scripts/dev/_variants/lib_trace.so RN_PAIR_DEBUG=5 python scripts/pair_trace_dump.py cfg3 gpurun_out/z4_trace_nored.npz 2>&1 | tail -14 | tee gpurun_out/z4_trace_nored.txt
for dbg in 0 4; do echo "== RN_PAIR_DEBUG=$dbg"; RN_PAIR_DEBUG=$dbg python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps" | sed -E 's/.*(B=65536 n_pair=[0-9]+ [0-9.]+ us\/call).*/\1/; s/.*(16:[0-9.]+) .*(20:[0-9.]+ 21:[0-9.]+ 22:[0-9.]+ 23:[0-9.]+)/   \1 \2/'; done | tee gpurun_out/z4_nored_time.txt
