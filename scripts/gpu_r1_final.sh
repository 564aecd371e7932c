# Round-1c record run: tests, bench (both arms), ncu launch list of the bench command, ncu full capture of k_pair / k_seg
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/final_tests.txt
python bench.py --steps 200 --warmup 20 2> gpurun_out/final_bench_err.log | tee gpurun_out/final_bench_n1.json
tail -3 gpurun_out/final_bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/final_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/final_b_ncu.log 2>&1
grep -c k_pair gpurun_out/final_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 5 -c 1 -f -o gpurun_out/prof_kpair_r1c python scripts/quick_time.py cfg3 > gpurun_out/final_ncu_kpair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_seg -s 5 -c 1 -f -o gpurun_out/prof_kseg_r1c python scripts/quick_time.py cfg3 > gpurun_out/final_ncu_kseg.log 2>&1
python scripts/quick_time.py cfg1 cfg2 cfg3 2>&1 | grep -v "^$" | tee gpurun_out/final_quick_time.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
