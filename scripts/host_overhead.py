"""Dev tool: host enqueue time vs device time of the public API calls."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops
from rec_now_b200.rec_block import pairwise_loss_from_batch as PW, listwise_loss_from_batch as LW

d = G.cfg3(0)
s, y, w = (torch.tensor(d[k]).cuda() for k in ("s", "y", "w"))
g = torch.tensor(d["g"]).cuda()
keys = g.reshape(1, -1)

def timeit(name, fn, iters=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(iters): fn()
    t_host = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} host enqueue {t_host/iters*1e6:7.1f} us/call   device {e0.elapsed_time(e1)/iters*1e3:7.1f} us/call")

timeit("ops.pairwise_fwd_bwd", lambda: ops.pairwise_fwd_bwd(s, y, keys, rw_pos=w, label_func="diff", power=-0.5))
def api():
    lg = s.detach().requires_grad_(True)
    loss = PW.pairwise_loss(lg, y, g, click_occurance_power=-0.5, label_pair_to_weight_func=PW.label_gain_times_sample_weight, sample_weight=w)
    loss.backward()
timeit("pairwise_loss + backward", api)
timeit("canon_keys(int64)", lambda: ops.canon_keys([g]))
d4 = G.cfg4(0)
s4, y4, g4 = (torch.tensor(d4[k]).cuda() for k in ("s", "y", "g"))
timeit("ops.listwise_fwd_bwd cfg4", lambda: ops.listwise_fwd_bwd(g4, y4, s4))
def lw():
    lg = s4.detach().requires_grad_(True)
    m, lab, lgt = LW.to_listwise_sample(g4, y4, lg)
    LW.listwise_loss_via_softmax_cross_entropy_with_logits(lab, lgt).backward()
timeit("to_listwise_sample+loss+backward cfg4", lw)
d2 = G.cfg2(0)
s2, y2, g2 = (torch.tensor(d2[k]).cuda() for k in ("s", "y", "g"))
timeit("ops.pairwise_fwd_bwd cfg2", lambda: ops.pairwise_fwd_bwd(s2, y2, g2.reshape(1, -1)))
