set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 200 --warmup 20 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json
tail -3 gpurun_out/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench_ref.json
python scripts/quick_time.py cfg1 cfg2 cfg3 2>&1 | tee gpurun_out/quick_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 5 -c 1 -f -o gpurun_out/prof_kpair_r1 python scripts/quick_time.py cfg3 > gpurun_out/ncu_kpair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_seg -s 5 -c 1 -f -o gpurun_out/prof_kseg_r1 python scripts/quick_time.py cfg3 > gpurun_out/ncu_kseg.log 2>&1
ls -la gpurun_out | head -30
