timeout 300 python - <<'PY'
import sys, torch, numpy as np, ctypes as C
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import metrics, _lib
from scripts.quick_time import _ramp
_ramp()
d = G.cfg3(0)
s, y, g = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda(), torch.tensor(d["g"]).cuda()
for _ in range(10): out = metrics.gauc(s, y, g, return_details=True)
torch.cuda.synchronize()
ts = (C.c_uint64 * 34)()
_lib.lib().rn_debug_timestamps(out["_scratch"].data_ptr(), ts, 34, None)
t = list(ts); t0 = t[0]
print("stamps:", " ".join(f"{i}:{(x - t0) / 1e3:.1f}" for i, x in enumerate(t[:24]) if x and x >= t0))
PY
