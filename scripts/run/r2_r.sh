set -x
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scripts/run/r2_q.sh 2>&1 | tail -4
