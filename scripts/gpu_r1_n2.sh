mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m pytest tests/test_global_nccl_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 2> gpurun_out/bench_n${N}_err.log | tee gpurun_out/bench_n${N}.json | cut -c1-330
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_err.log | tail -5
RN_GLOBAL_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 10 2> /dev/null | tee gpurun_out/bench_n${N}_nccl.json | cut -c1-330
