set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python scripts/quick_time.py cfg2 cfg3 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 6 -c 2 -o gpurun_out/prof_kpair_r1a python scripts/quick_time.py cfg3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sort_pass -s 24 -c 4 -o gpurun_out/prof_sort_r1a python scripts/quick_time.py cfg3 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/ncu_full2.log
