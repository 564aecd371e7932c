mkdir -p gpurun_out
python -m pytest tests/test_global_nccl_gpu.py tests/test_launch_and_peer_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/zg_tests_2gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 2> gpurun_out/zg_bench_n2_err.log | tee gpurun_out/zg_bench_n2.json
tail -2 gpurun_out/zg_bench_n2_err.log
