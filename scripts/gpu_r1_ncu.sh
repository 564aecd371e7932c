mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/final_b_ncu.log 2>&1
grep -c k_pair gpurun_out/final_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 5 -c 1 -f -o gpurun_out/prof_kpair_r1c python scripts/quick_time.py cfg3 > gpurun_out/final_ncu_kpair.log 2>&1
ls -la gpurun_out/prof_kpair_r1c.ncu-rep
