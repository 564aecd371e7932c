mkdir -p gpurun_out
RN_LIB_PATH=$PWD/scripts/dev/_variants/lib_trace.so RN_PAIR_DEBUG=1 python scripts/pair_trace_dump.py cfg3 gpurun_out/z3_trace.npz 2>&1 | tail -14 | tee gpurun_out/z3_trace.txt
RN_SEG_MERGED=0 RN_LIB_PATH=$PWD/scripts/dev/_variants/lib_trace.so RN_PAIR_DEBUG=1 python scripts/pair_trace_dump.py cfg3 gpurun_out/z3_trace_nomerge.npz 2>&1 | tail -14 | tee gpurun_out/z3_trace_nomerge.txt
