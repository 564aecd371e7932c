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

def _ramp():
    """Clock ramp: ~0.5 s of SFU work so that the timed calls run at the boost clock."""
    import ctypes as C, time, torch
    from rec_now_b200 import _lib
    sink = torch.zeros(4, device="cuda"); n = C.c_int64(0)
    t = time.perf_counter() + 0.5
    while time.perf_counter() < t:
        _lib.lib().rn_bench_mufu(2000, sink.data_ptr(), C.byref(n), None); torch.cuda.synchronize()
_ramp()
for _ in range(50):
    out = ops.pairwise_fwd_bwd(s, y, keys, rw_pos=w, label_func="diff", power=-0.5)
torch.cuda.synchronize()
ts = (C.c_uint64 * 34)()
_lib.lib().rn_debug_timestamps(out["_scratch"].data_ptr(), ts, 34, None)
t = list(ts); d = t[24:32]
print(f"npos {npos} nneg {nneg} n_pair {int(out['n_pair'])} units={t[32] & 0xFFFFFFFF} C={t[32] >> 32} tiles={t[33]}  k_pair loop {(t[21]-t[20])/1e3:.1f} us  "
      f"longest unit {d[0] >> 32} cyc; total busy {d[1]} cyc; fast tiles {d[5]} general {d[6]} -> {d[1]/max(d[5]+d[6],1):.0f} cyc/tile; inside fast tiles {d[3]/max(d[5],1):.0f} cyc/tile, general {d[7]/max(d[6],1):.0f}; k_pair start->barrier {(t[22]-t[20])/1e3:.1f} us")
