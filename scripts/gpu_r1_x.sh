set -x
mkdir -p gpurun_out
python scripts/quick_time.py cfg2 cfg3 2>&1 | tee gpurun_out/quick_time.txt
RN_PAIR_DEBUG=1 NW=32 python scripts/pair_debug.py cfg3 2>&1 | tee gpurun_out/pair_debug.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 5 -c 1 -f -o gpurun_out/prof_kpair_r1b python scripts/quick_time.py cfg3 > gpurun_out/ncu_kpair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_seg -s 5 -c 1 -f -o gpurun_out/prof_kseg_r1b python scripts/quick_time.py cfg3 > gpurun_out/ncu_kseg.log 2>&1
ls -la gpurun_out
