mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 2> gpurun_out/chk2_err.log | tee gpurun_out/chk_bench_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('N=2 step %.1f us  b2b %.1f | k_pair %.2f us frac %.3f (plain %.2f) | e2e %.1f us | %s'%(d['ms_per_step']*1e3,d['step_us']['back_to_back_no_flush'],r['kernel_ms']*1e3,r['frac'],r['kernel_ms_plain_launches']*1e3,d['e2e']['ms_per_step']*1e3,d['config']['multi_gpu'][:60]))"
tail -3 gpurun_out/chk2_err.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | cut -c1-200
