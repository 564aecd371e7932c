set -x
timeout 900 python -m pytest tests/test_counting_path_gpu.py tests/test_pairwise_gpu.py -x -q 2>&1 | tail -5
RN_SEG_DEBUG=1 timeout 300 python scripts/seg_debug.py cfg3 2>&1 | tail -8
timeout 300 python scripts/quick_time.py cfg1 cfg2 cfg3 2>&1 | tail -12
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err; tail -5 gpurun_out/r2_bench_n1_d.err
