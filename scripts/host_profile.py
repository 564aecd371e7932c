"""Dev tool: host-side cost of the pieces of the public pairwise API (no device sync inside the loops)."""
import sys, time, cProfile, pstats
import torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops
from rec_now_b200.rec_block import pairwise_loss_from_batch as PW

d = G.cfg3(0)
s, y, w = (torch.tensor(d[k]).cuda() for k in ("s", "y", "w"))
g = torch.tensor(d["g"]).cuda()
kw = dict(click_occurance_power=-0.5, label_pair_to_weight_func=PW.label_gain_times_sample_weight, sample_weight=w)

def t(name, fn, iters=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters): fn()
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{name:50s} host {th / iters * 1e6:7.1f} us")

t("ops.pairwise_fwd_bwd", lambda: ops.pairwise_fwd_bwd(s, y, g.reshape(1, -1), rw_pos=w, label_func="diff", power=-0.5))
t("pairwise_loss forward, no grad", lambda: PW.pairwise_loss(s, y, g, **kw))
lg = s.detach().requires_grad_(True)
t("pairwise_loss forward, requires_grad", lambda: PW.pairwise_loss(lg, y, g, **kw))
def fb():
    loss = PW.pairwise_loss(lg, y, g, **kw); loss.backward(); lg.grad = None
t("forward + backward()", fb)
def fg():
    loss = PW.pairwise_loss(lg, y, g, **kw); torch.autograd.grad(loss, lg)
t("forward + autograd.grad", fg)
t("torch.empty x2", lambda: (torch.empty(4, device="cuda"), torch.empty(65536, device="cuda")))
pr = cProfile.Profile(); pr.enable()
for _ in range(300): fb()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
with torch.autograd.set_multithreading_enabled(False):
    t("forward + backward(), engine single-threaded", fb)
