set -x
python scripts/quick_time.py cfg2 cfg3
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 40 --csv --log-file gpurun_out/launches_cfg3.csv python scripts/quick_time.py cfg3 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
