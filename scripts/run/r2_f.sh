set -x
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
RN_SEG_DEBUG=1 timeout 300 python scripts/seg_debug.py cfg3 2>&1 | tail -9
timeout 300 python scripts/quick_time.py cfg1 cfg2 cfg3 2>&1 | tail -12
timeout 600 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops
from scripts.quick_time import _ramp
d = G.cfg4(0)
s, y, k = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda(), torch.tensor(d["g"]).cuda()
_ramp()
for name, kw in (("counting", {}), ("sorted", dict(sorted_form=True))):
    for _ in range(10): out = ops.listwise_fwd_bwd(k, y, s, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): out = ops.listwise_fwd_bwd(k, y, s, **kw)
    e1.record(); torch.cuda.synchronize()
    print(name, "cfg4 listwise", e0.elapsed_time(e1) / 200 * 1e3, "us/call", float(out["loss"]), int(out["n_valid"]))
    import ctypes as C
    from rec_now_b200 import _lib
    ts = (C.c_uint64 * 34)()
    _lib.lib().rn_debug_timestamps(out["_scratch"].data_ptr(), ts, 34, None)
    t = list(ts); t0 = t[0]
    print("   stamps:", " ".join(f"{i}:{(x - t0) / 1e3:.1f}" for i, x in enumerate(t[:24]) if x and x >= t0))
PY
