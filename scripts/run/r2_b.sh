set -x
timeout 900 python -m pytest tests/test_counting_path_gpu.py -x -q 2>&1 | tail -25
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/quick_time.py cfg1 cfg2 cfg3 2>&1 | tail -20
RN_PAIR_PREPART=0 timeout 300 python scripts/quick_time.py cfg3 2>&1 | tail -6
# sanitizer evidence: memcheck / racecheck / initcheck over the small parity cases
for tool in memcheck racecheck initcheck; do
  timeout 420 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.txt --print-limit 20 \
    python -m pytest tests/test_counting_path_gpu.py tests/test_pairwise_gpu.py -x -q -k "baseline_configs or fallback_to_radix or known_answers or degenerate or test_cfg1" 2>&1 | tail -4
  tail -5 gpurun_out/sanitizer_$tool.txt
done
