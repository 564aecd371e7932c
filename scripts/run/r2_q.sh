set -x
timeout 900 python -m pytest tests/test_gauc_gpu.py tests/test_listwise_counting_gpu.py -x -q 2>&1 | tail -4
timeout 300 python - <<'PY'
import sys, torch, numpy as np, ctypes as C
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import metrics, _lib, ops
from scripts.quick_time import _ramp
_ramp()
def stamps(scr):
    ts = (C.c_uint64 * 34)()
    _lib.lib().rn_debug_timestamps(scr.data_ptr(), ts, 34, None)
    t = list(ts); t0 = t[0]
    return " ".join(f"{i}:{(x - t0) / 1e3:.1f}" for i, x in enumerate(t[:24]) if x and x >= t0)
for name in ("cfg2", "cfg3"):
    d = getattr(G, name)(0)
    s, y, g = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda(), torch.tensor(d["g"]).cuda()
    for _ in range(10): out = metrics.gauc(s, y, g, return_details=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): out = metrics.gauc(s, y, g, return_details=True)
    e1.record(); torch.cuda.synchronize()
    print(name, "gauc", float(out["gauc"]), e0.elapsed_time(e1) / 200 * 1e3, "us/call", stamps(out["_scratch"]))
d = G.cfg4(0)
s, y, k = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda(), torch.tensor(d["g"]).cuda()
for _ in range(10): out = ops.listwise_fwd_bwd(k, y, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): out = ops.listwise_fwd_bwd(k, y, s)
e1.record(); torch.cuda.synchronize()
print("cfg4 listwise counting", e0.elapsed_time(e1) / 200 * 1e3, "us/call", stamps(out["_scratch"]))
PY
