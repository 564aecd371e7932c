"""Dev tool: spread of the warps' finishing times when every piece is the same work (one group, two label levels:
all tiles fast).  Usage: RN_PAIR_DEBUG=1 python scripts/uniform_probe.py [npos nneg]"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from rec_now_b200 import ops, _lib
from scripts.quick_time import _ramp
npos, nneg = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10240, 4096)
rng = np.random.default_rng(0)
B = npos + nneg
y = np.r_[np.ones(npos), np.zeros(nneg)].astype(np.float32)
perm = rng.permutation(B)
s = torch.tensor(rng.standard_normal(B).astype(np.float32)[perm]).cuda(); y = torch.tensor(y[perm]).cuda()
keys = torch.zeros(1, B, dtype=torch.int64).cuda()
w = torch.tensor(rng.uniform(0.5, 1.5, B).astype(np.float32)).cuda()
_ramp()
for _ in range(50):
    out = ops.pairwise_fwd_bwd(s, y, keys, rw_pos=w, label_func="diff", power=-0.5)
torch.cuda.synchronize()
scr = out["_scratch"]
ts = (C.c_uint64 * 34)(); _lib.lib().rn_debug_timestamps(scr.data_ptr(), ts, 34, None); t = list(ts)
total = scr.numel(); gstat_off = _lib.lib().rn_debug_arena_offset(B, 1, 0)
rec = scr[gstat_off:gstat_off + 148 * 32 * 64].view(torch.int64).cpu().numpy().reshape(-1, 8)
t20 = t[20]; start = (rec[:, 0] - t20) / 1e3; end = (rec[:, 1] - t20) / 1e3; eig = rec[:, 7] & 0xFFFFFFFF; gen = rec[:, 3] >> 32
ok = rec[:, 1] > 0
dur = end - start
print(f"n_pair {int(out['n_pair'])} tiles {t[33]}  eighths/warp min {eig[ok].min()} med {np.median(eig[ok])} max {eig[ok].max()}  general tiles total {gen[ok].sum()}")
print(f"k_pair: list {(t[16]-t20)/1e3:.1f}  loop end (cta0) {(t[21]-t20)/1e3:.1f}  barrier {(t[22]-t20)/1e3:.1f} us")
print(f"dur: min {dur[ok].min():.1f} p10 {np.percentile(dur[ok],10):.1f} med {np.median(dur[ok]):.1f} p90 {np.percentile(dur[ok],90):.1f} p99 {np.percentile(dur[ok],99):.1f} max {dur[ok].max():.1f} us")
E = end.reshape(148, 32)
print("per-SM last exit: min %.1f med %.1f max %.1f; per-SM median exit: min %.1f max %.1f" % (E.max(1).min(), np.median(E.max(1)), E.max(1).max(), np.median(E, 1).min(), np.median(E, 1).max()))
wq = np.arange(len(dur)) % 32
print("median dur by warp index:", " ".join(f"{np.median(dur[wq == k]):.0f}" for k in range(32)))
print("max dur by warp index:   ", " ".join(f"{dur[wq == k].max():.0f}" for k in range(32)))
