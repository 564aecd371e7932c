"""Dev tool: per-kernel timeline of the global in-batch step (torchrun, >= 2 GPUs): CUPTI kernel records of rank 0 for
a few steps -- duration of every kernel of the step and the gaps between them."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import global_mode
from scripts.quick_time import _ramp

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = G.cfg5(world, 0)
lo, hi = rank * 65536, (rank + 1) * 65536
t = lambda k: torch.tensor(np.ascontiguousarray(d[k][lo:hi]), device="cuda")
s, y, g, w = t("s"), t("y"), t("g").reshape(1, -1), t("w")
step = lambda: global_mode.global_pairwise_fwd_bwd(s, y, g, rw_pos=w, label_func="diff", power=-0.5)
_ramp()
for _ in range(20):
    step()
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        step()
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    n = len(ev) // 10
    print("kernels per step:", n)
    rows = {}
    for k, e in enumerate(ev):
        j = k % n
        prev_end = ev[k - 1].time_range.end if k else e.time_range.start
        rows.setdefault(j, []).append((e.name[:60], e.time_range.end - e.time_range.start, e.time_range.start - prev_end))
    tot = 0.0
    for j in range(n):
        name = rows[j][0][0]
        dur = np.mean([r[1] for r in rows[j][2:]]); gap = np.mean([r[2] for r in rows[j][2:]])
        tot += dur + gap
        print(f"  {name:60s} dur {dur:7.1f} us   gap before {gap:6.1f} us")
    print(f"  step total {tot:.1f} us")
dist.destroy_process_group()
