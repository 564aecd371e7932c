#!/bin/bash
# Dev helper: retry a gpurun call while the pod answers "transient / busy" (nothing is charged for those).
# Usage: scripts/dev/gpurun_retry.sh <timeout_s> <script> [gpus]
for i in $(seq 1 12); do
  if [ -n "$3" ]; then out=$(/usr/local/graft/bin/gpurun --gpus $3 --timeout $1 -- "bash $2" 2>&1); else out=$(/usr/local/graft/bin/gpurun --timeout $1 -- "bash $2" 2>&1); fi
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then sleep 150; continue; fi
  echo "$out" | tail -60; exit 0
done
echo "gave up"; echo "$out" | tail -5
