"""Dev tool (torchrun): where the global-mode step's device time goes -- pack, all-gather, kernels, reduce-scatter."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops, global_mode

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
R = 65536
d = G.cfg5(world, 0)
lo, hi = rank * R, (rank + 1) * R
s, y, w = (torch.tensor(np.ascontiguousarray(d[k][lo:hi]), device=dev) for k in ("s", "y", "w"))
keys = torch.tensor(np.ascontiguousarray(d["g"][lo:hi]), device=dev).reshape(1, -1)
lay = ops.packed_block_layout(R, 1, True, False)

def stages(ev):
    ev[0].record()
    cols = [keys.reshape(-1).view(torch.uint8), s.view(torch.uint8), y.view(torch.uint8), w.view(torch.uint8)]
    used = sum(c.numel() for c in cols)
    if used != lay["stride"]:
        cols.append(torch.zeros(lay["stride"] - used, dtype=torch.uint8, device=dev))
    block = torch.cat(cols)
    ev[1].record()
    gbuf = torch.empty(world * lay["stride"], dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gbuf, block)
    ev[2].record()
    res = ops.pairwise_fwd_bwd_blocked(gbuf, world, R, 1, True, False, label_func="diff", power=-0.5, part=(rank, world))
    ev[3].record()
    mine = torch.empty(res["chunk"], dtype=torch.float32, device=dev)
    dist.reduce_scatter_tensor(mine, res["out"], op=dist.ReduceOp.SUM)
    ev[4].record()
    return res

for _ in range(20):
    stages([torch.cuda.Event(enable_timing=True) for _ in range(5)])
dist.barrier(); torch.cuda.synchronize()
acc = np.zeros(4); N = 100
for _ in range(N):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    res = stages(ev)
    torch.cuda.synchronize()
    acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    dist.barrier()
import ctypes as C
from rec_now_b200 import _lib
ts = (C.c_uint64 * 34)()
_lib.lib().rn_debug_timestamps(res["_scratch"].data_ptr(), ts, 34, None)
t = list(ts); t0 = t[0]
print(f"rank {rank}: pack {acc[0]/N*1e3:.1f} us  all-gather {acc[1]/N*1e3:.1f} us  kernels {acc[2]/N*1e3:.1f} us  reduce-scatter {acc[3]/N*1e3:.1f} us | "
      "stamps " + " ".join(f"{i}:{(x - t0) / 1e3:.1f}" for i, x in enumerate(t[:24]) if x), flush=True)
dist.destroy_process_group()
