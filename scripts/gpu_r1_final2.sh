# Record run without the ncu captures (tests, both bench arms, phase stamps, smoke)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/final_tests.txt
python bench.py --steps 200 --warmup 20 2> gpurun_out/final_bench_err.log | tee gpurun_out/final_bench_n1.json
tail -3 gpurun_out/final_bench_err.log
python scripts/quick_time.py cfg1 cfg2 cfg3 2>&1 | grep -v "^$" | tee gpurun_out/final_quick_time.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
