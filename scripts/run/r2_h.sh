set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_global_nccl_gpu.py tests/test_launch_and_peer_gpu.py -x -q 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2_bench_n2_h.json 2> gpurun_out/r2_bench_n2_h.err; tail -3 gpurun_out/r2_bench_n2_h.err
