"""Dev tool: does replaying the 3 launches of one call from a CUDA graph shrink the inter-kernel gaps?"""
import sys
import torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops
from scripts.quick_time import _ramp, stamps

d = getattr(G, sys.argv[1] if len(sys.argv) > 1 else "cfg3")()
s, y = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda()
keys = torch.tensor(d["g"]).cuda().reshape(1, -1)
w = torch.tensor(d["w"]).cuda() if "w" in d else None
kw = dict(label_func=d["label_func"], power=d["power"], rw_pos=w)
_ramp()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(5):
        out = ops.pairwise_fwd_bwd(s, y, keys, **kw)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = ops.pairwise_fwd_bwd(s, y, keys, **kw)
torch.cuda.synchronize()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 200
e0.record()
for _ in range(iters):
    g.replay()
e1.record(); torch.cuda.synchronize()
n = int(out["n_pair"].item())
print(f"graph replay: {e0.elapsed_time(e1) / iters * 1e3:.1f} us/call n_pair={n} loss={out['loss'].item():.6f}")
stamps(out)
e0.record()
for _ in range(iters):
    out2 = ops.pairwise_fwd_bwd(s, y, keys, **kw)
e1.record(); torch.cuda.synchronize()
print(f"stream launches: {e0.elapsed_time(e1) / iters * 1e3:.1f} us/call")
stamps(out2)
