"""Dev tool: cycles per fast tile for a lone warp (one I-block x many J-blocks)."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from rec_now_b200 import ops, _lib
npos, nneg = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
B = npos + nneg
y = np.r_[np.ones(npos), np.zeros(nneg)].astype(np.float32)
s = rng.standard_normal(B).astype(np.float32)
g = np.zeros(B, np.int64)
perm = rng.permutation(B)
s, y = torch.tensor(s[perm]).cuda(), torch.tensor(y[perm]).cuda()
keys = torch.tensor(g).cuda().reshape(1, -1)
w = torch.tensor(rng.uniform(0.5, 1.5, B).astype(np.float32)).cuda()
for _ in range(3):
    out = ops.pairwise_fwd_bwd(s, y, keys, rw_pos=w, label_func="diff", power=-0.5)
torch.cuda.synchronize()
ts = (C.c_uint64 * 34)()
_lib.lib().rn_debug_timestamps(out["_scratch"].data_ptr(), ts, 34, None)
t = list(ts); d = t[24:32]
print(f"npos {npos} nneg {nneg} n_pair {int(out['n_pair'])} units={t[32] & 0xFFFFFFFF} C={t[32] >> 32} tiles={t[33]}  k_pair loop {(t[21]-t[20])/1e3:.1f} us  "
      f"longest unit {d[0] >> 32} cyc; total busy {d[1]} cyc; fast tiles {d[5]} general {d[6]} -> {d[1]/max(d[5]+d[6],1):.0f} cyc/tile")
