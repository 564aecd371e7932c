set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 200 --warmup 20 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/b_ncu.log 2>&1
tail -40 gpurun_out/launches.csv
