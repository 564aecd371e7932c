"""Per-warp phase stamps of the one-CTA kernel of small batches (RN_SMALL_DEBUG=1): where its time goes (dev tool)."""
import os, sys
os.environ["RN_SMALL_DEBUG"] = "1"
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops, _lib

d = getattr(G, sys.argv[1] if len(sys.argv) > 1 else "cfg1")()
s, y = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda()
keys = torch.tensor(d["g"]).cuda().reshape(1, -1)
for _ in range(20):
    out = ops.pairwise_fwd_bwd(s, y, keys)
torch.cuda.synchronize()
off = _lib.lib().rn_debug_arena_offset(s.numel(), 1, 0)
raw = out["_scratch"][off:off + 8 * 32 * 8].cpu().numpy().view(np.uint64).reshape(8, 32).astype(np.int64)
t0 = raw[0].min()
names = {0: "start", 1: "grouped", 2: "placed", 5: "compacted", 6: "positive pass", 7: "negative pass", 4: "results read"}
for ph in (0, 1, 2, 5, 6, 7, 4):
    r = (raw[ph] - t0) / 1e3
    print(f"phase {ph} {names[ph]:14s}: min {r.min():5.2f} med {np.median(r):5.2f} max {r.max():5.2f} us   by warp/4: " +
          " ".join(f"{r[k:k + 4].mean():.1f}" for k in range(0, 32, 4)))
