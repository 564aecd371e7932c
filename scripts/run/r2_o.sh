set -x
timeout 900 python -m pytest tests/test_gauc_gpu.py -x -q 2>&1 | tail -12
timeout 300 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import metrics
from scripts.quick_time import _ramp
_ramp()
for name in ("cfg2", "cfg3"):
    d = getattr(G, name)(0)
    s, y, g = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda(), torch.tensor(d["g"]).cuda()
    for _ in range(10): out = metrics.gauc(s, y, g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): out = metrics.gauc(s, y, g)
    e1.record(); torch.cuda.synchronize()
    print(name, "gauc", float(out), e0.elapsed_time(e1) / 200 * 1e3, "us/call")
PY
