timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 6 -c 1 -o gpurun_out/prof_kpair_r1m python scripts/quick_time.py cfg3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
