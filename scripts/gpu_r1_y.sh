mkdir -p gpurun_out
for a in "64 2048" "64 32768" "4096 4096"; do RN_PAIR_DEBUG=1 python scripts/lone_warp.py $a 2>&1 | tail -1; done | tee gpurun_out/lone_warp.txt
for tu in 4096 8192 16384 32768 65536; do echo "RN_TARGET_UNITS=$tu"; RN_TARGET_UNITS=$tu python scripts/quick_time.py cfg3 2>&1 | grep -v "^$" ; done | tee gpurun_out/target_units.txt
