mkdir -p gpurun_out
python bench.py --steps 100 --warmup 10 --no-cpu 2> gpurun_out/chk_err.log | tee gpurun_out/chk_bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('step %.1f us  %s | k_pair %.2f us frac %.3f | e2e %.1f us'%(d['ms_per_step']*1e3,d['step_us'],r['kernel_ms']*1e3,r['frac'],d['e2e']['ms_per_step']*1e3))"
tail -2 gpurun_out/chk_err.log
