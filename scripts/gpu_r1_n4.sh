mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 50 --warmup 5 2> gpurun_out/n4_err.log | tee gpurun_out/bench_n4_r1c.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('N=4 step %.1f us  b2b %.1f | value %.3f T pairs/s | k_pair %.2f us frac %.3f | e2e %.1f us'%(d['ms_per_step']*1e3,d['step_us']['back_to_back_no_flush'],d['value']/1e12,r['kernel_ms']*1e3,r['frac'],d['e2e']['ms_per_step']*1e3))"
tail -2 gpurun_out/n4_err.log
