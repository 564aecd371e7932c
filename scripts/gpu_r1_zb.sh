mkdir -p gpurun_out
python -m pytest tests/test_host_pipeline_gpu.py tests/test_pairwise_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/zb_tests.txt
python bench.py --steps 200 --warmup 20 --no-cpu 2> gpurun_out/zb_bench_err.log | tee gpurun_out/zb_bench_n1.json
tail -5 gpurun_out/zb_bench_err.log
