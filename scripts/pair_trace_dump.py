"""Dev tool: per-warp time breakdown of the pair kernel (library built with -DRN_TRACE, RN_PAIR_DEBUG=1) plus the
sorted-row arrays of the call, saved for offline analysis.  Usage: RN_LIB_PATH=.../lib_trace.so RN_PAIR_DEBUG=1 python scripts/pair_trace_dump.py cfg3 out.npz"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops, _lib
from scripts.quick_time import _ramp

d = getattr(G, sys.argv[1])()
s, y = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda()
keys = torch.tensor(d["g"]).cuda().reshape(1, -1)
w = torch.tensor(d["w"]).cuda() if "w" in d else None
kw = dict(label_func=d["label_func"], power=d["power"], rw_pos=w)
_ramp()
for _ in range(30):
    out = ops.pairwise_fwd_bwd(s, y, keys, **kw)
torch.cuda.synchronize()
scr = out["_scratch"]
ts = (C.c_uint64 * 34)()
_lib.lib().rn_debug_timestamps(scr.data_ptr(), ts, 34, None)
B = s.numel(); nib = (B + 63) // 64
al = lambda x: (x + 255) // 256 * 256
off = scr.numel()
def back(nbytes):
    global off
    off -= al(nbytes); return off
o_gstat = __import__('rec_now_b200._lib', fromlist=['lib']).lib().rn_debug_arena_offset(B, 1, 0); _ = back(8 * 4 * B); o_misc = back(8 * (B + 1)); o_units = back(8 * (2 * nib + 16384 + 1)); o_blk = back(16 * nib)
o_perm = back(4 * B); o_cnt = back(4 * B); o_loss = back(4 * B); o_gacc = back(4 * B); o_swn = back(4 * B); o_swp = back(4 * B)
o_sy = back(4 * B); o_ss = back(4 * B); o_aj = back(8 * B)
g = lambda o, n, dt: scr[o:o + n].view(dt).cpu().numpy()
rec = g(o_gstat, 148 * 32 * 64, torch.int64).reshape(-1, 8)
np.savez(sys.argv[2], rec=rec, ts=np.array(list(ts), dtype=np.uint64), blk=g(o_blk, 16 * nib, torch.int32).reshape(-1, 2),
         aj=g(o_aj, 8 * B, torch.int32).reshape(-1, 2), sy=g(o_sy, 4 * B, torch.float32), perm=g(o_perm, 4 * B, torch.int32))
t20 = ts[20]
segs = (rec[:, 2] >> 40) & 0xFF; tiles = rec[:, 2] >> 48; rec = rec.copy(); rec[:, 2] &= (1 << 40) - 1
end = (rec[:, 1] - t20) / 1e3
print("medians: busy %d take %d pre %d tile %d post %d flush %d" % tuple(np.median(rec[:, 2:8], axis=0)), "segs/warp mean %.2f tiles/warp mean %.2f" % (segs.mean(), tiles.mean()))
print("sums (Mcyc): busy %.1f take %.1f pre %.1f tile %.1f post %.1f flush %.1f" % tuple(rec[:, 2:8].sum(0) / 1e6))
print("exit pct", np.percentile(end, [0, 10, 50, 90, 99, 100]).round(1), "barrier", (ts[22] - t20) / 1e3)
for i in np.argsort(-end)[:10]:
    print(i // 32, i % 32, "end %.1f busy %d take %d pre %d tile %d post %d flush %d" % (end[i], *rec[i, 2:8]))
i = np.argsort(end)[2368]
print("median:", i // 32, i % 32, "end %.1f busy %d take %d pre %d tile %d post %d flush %d" % (end[i], *rec[i, 2:8]))
