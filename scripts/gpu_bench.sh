set -x
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 200 --warmup 20 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_err.log
