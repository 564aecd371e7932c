"""Dev tool: per-CTA arrival times at the grid barriers of k_seg (RN_SEG_DEBUG=1)."""
import ctypes as C, sys
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
B = s.numel()
gstat_off = _lib.lib().rn_debug_arena_offset(B, 1, 0)
grid = torch.cuda.get_device_properties(0).multi_processor_count
NPH = 7
a = scr[gstat_off:gstat_off + 8 * NPH * grid].view(torch.int64).cpu().numpy().reshape(NPH, grid)
t0 = ts[0]
print("phases of the counting path: 0 count done, 1 past barrier 1, 2 offsets done, 3 past barrier 2, 4 scatter / partition done; helper CTAs only: 5 costs in shared memory, 6 scanned")
for ph in range(NPH):
    if ph >= 5:
        a[ph] = np.where(a[ph] < t0, t0, a[ph])
    t = (a[ph] - t0) / 1e3
    print(f"phase {ph}: arrival min {t.min():.1f} med {np.median(t):.1f} max {t.max():.1f} us; slowest CTAs {np.argsort(-t)[:6]} ; by CTA/8: " + " ".join(f"{t[k:k+8].max():.1f}" for k in range(0, grid, 8)))
print("stamps", " ".join(f"{i}:{(x - t0) / 1e3:.1f}" for i, x in enumerate(list(ts)[:24]) if x))
