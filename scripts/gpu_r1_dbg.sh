mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo scripts/dev/tile_bench.cu -o scripts/dev/tile_bench 2>&1 | tail -3
./scripts/dev/tile_bench 2>&1 | tee gpurun_out/tile_bench.txt
RN_PAIR_DEBUG=1 NW=32 python scripts/pair_debug.py cfg3 2>&1 | tee gpurun_out/pair_debug.txt
RN_PAIR_DEBUG=1 python scripts/quick_time.py cfg3 2>&1 | tail -5
