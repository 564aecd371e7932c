mkdir -p gpurun_out
RN_LIB_PATH=$PWD/scripts/dev/_variants/lib_trace.so RN_PAIR_DEBUG=1 python scripts/pair_trace_dump.py cfg3 gpurun_out/trace_r1c.npz 2>&1 | tail -16 | tee gpurun_out/trace_r1c.txt
