mkdir -p gpurun_out
for f in "" 1 "" 1; do echo "== RN_QT_FLUSH=$f"; RN_QT_FLUSH=$f python scripts/quick_time.py cfg3 2>&1 | grep -E "us/call|stamps"; done | tee gpurun_out/ze_cold_stamps.txt
