mkdir -p gpurun_out
for dbg in 0 8 0 8; do echo "== RN_PAIR_DEBUG=$dbg"; RN_PAIR_DEBUG=$dbg python bench.py --steps 100 --warmup 10 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('step %.1f us  k_pair %.2f us  frac %.3f  e2e %.1f us'%(d['ms_per_step']*1e3,d['roofline']['kernel_ms']*1e3,d['roofline']['frac'],d['e2e']['ms_per_step']*1e3))"; done | tee gpurun_out/zd_cold.txt
