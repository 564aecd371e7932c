mkdir -p gpurun_out
python -m pytest tests/test_pairwise_gpu.py -m gpu -x -q -k "known or options_small or cfg3" 2>&1 | tail -2
python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps" | cut -c1-230
