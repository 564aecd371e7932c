"""Quick device timing of rn_pairwise_fwd_bwd on the BASELINE configs (dev tool, not the bench)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops

def run(d, iters=50):
    s, y = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda()
    keys = torch.tensor(d["g"]).cuda().reshape(1, -1)
    w = torch.tensor(d["w"]).cuda() if "w" in d else None
    kw = dict(label_func=d["label_func"], power=d["power"], rw_pos=w)
    for _ in range(5):
        out = ops.pairwise_fwd_bwd(s, y, keys, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = ops.pairwise_fwd_bwd(s, y, keys, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    n = int(out["n_pair"].item())
    print(f"{d['name']}: B={s.numel()} n_pair={n} {ms*1e3:.1f} us/call  {n/ms/1e6:.2f} Gpairs/s  "
          f"SFU-frac(3 MUFU, 4.65e12/s)={3*n/(ms*1e-3)/4.65e12:.3f} loss={out['loss'].item():.6f} err={ops.device_error(out['_scratch'])}")

if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg3"]
    for name in which:
        run(getattr(G, name)())
