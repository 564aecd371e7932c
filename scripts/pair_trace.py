"""Dev tool: dump the per-unit trace of the pair kernel (RN_PAIR_DEBUG=2) to gpurun_out/pair_trace.npz."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops, _lib
from scripts.quick_time import _ramp

d = getattr(G, sys.argv[1] if len(sys.argv) > 1 else "cfg3")()
s, y = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda()
keys = torch.tensor(d["g"]).cuda().reshape(1, -1)
w = torch.tensor(d["w"]).cuda() if "w" in d else None
kw = dict(label_func=d["label_func"], power=d["power"], rw_pos=w)
_ramp()
for _ in range(100):
    out = ops.pairwise_fwd_bwd(s, y, keys, **kw)
torch.cuda.synchronize()
scr = out["_scratch"]
ts = (C.c_uint64 * 34)()
_lib.lib().rn_debug_timestamps(scr.data_ptr(), ts, 34, None)
B = s.numel(); NW = 32
total = scr.numel()
gstat_off = total - ((8 * 4 * B + 255) // 256) * 256
nwarp = 148 * NW
buf = scr[gstat_off:gstat_off + nwarp * (32 + 48 * 8)].view(torch.int64).cpu().numpy()
np.savez(sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/pair_trace.npz", ts=np.array(list(ts), dtype=np.uint64), rec=buf[:4 * nwarp].reshape(-1, 4),
         trace=buf[4 * nwarp:].reshape(nwarp, 48)[:, :36].reshape(nwarp, 6, 6))
print("saved", buf.shape)
