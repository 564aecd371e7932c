set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_global_nccl_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 2> gpurun_out/bench2_err.log | tee gpurun_out/bench_n2.json
tail -5 gpurun_out/bench2_err.log
