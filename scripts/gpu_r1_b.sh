set -x
timeout 120 python scripts/quick_time.py cfg1 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 120 python scripts/quick_time.py cfg2 cfg3 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 12 --csv --log-file gpurun_out/launches_cfg3_b.csv python scripts/quick_time.py cfg3 > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
