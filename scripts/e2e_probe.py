"""Dev tool: where the e2e step's time goes (host enqueue vs device) for variants of the input copy."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200.rec_block import pairwise_loss_from_batch as PW
d = G.cfg3(0); B = 65536; dev = torch.device("cuda")
host = {k: np.ascontiguousarray(d[k]) for k in ("g", "s", "y", "w")}
def timeit(name, fn, iters=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(iters): fn()
    th = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    print(f"{name:46s} host {th/iters*1e6:7.1f} us   device-timeline {e0.elapsed_time(e1)/iters*1e3:7.1f} us")
pin = {k: torch.tensor(host[k]).pin_memory() for k in host}
dbuf = {k: torch.empty_like(v, device=dev) for k, v in pin.items()}
sizes = {"g": 8 * B, "s": 4 * B, "y": 4 * B, "w": 4 * B}
pin_all = torch.empty(sum(sizes.values()), dtype=torch.uint8).pin_memory()
dev_all = torch.empty_like(pin_all, device=dev)
print("dev_all device", dev_all.device, "pinned", pin_all.is_pinned())
v = {}; o = 0
for k, dt in (("g", torch.int64), ("s", torch.float32), ("y", torch.float32), ("w", torch.float32)):
    pin_all[o:o + sizes[k]].view(dt).copy_(torch.tensor(host[k])); v[k] = dev_all[o:o + sizes[k]].view(dt); o += sizes[k]
h_loss = torch.empty(1).pin_memory(); h_grad = torch.empty(B).pin_memory()
def copy4():
    for k in dbuf: dbuf[k].copy_(pin[k], non_blocking=True)
def copy1(): dev_all.copy_(pin_all, non_blocking=True)
timeit("4 H2D copies", copy4); timeit("1 H2D copy", copy1)
def step(cols):
    lg = cols["s"].requires_grad_(True)
    loss = PW.pairwise_loss(lg, cols["y"], cols["g"], click_occurance_power=-0.5,
                            label_pair_to_weight_func=PW.label_gain_times_sample_weight, sample_weight=cols["w"])
    loss.backward()
    h_loss.copy_(loss.detach().reshape(1), non_blocking=True); h_grad.copy_(lg.grad, non_blocking=True)
    lg.grad = None; cols["s"].requires_grad_(False)
timeit("step on separate tensors (no H2D)", lambda: step(dbuf))
timeit("step on views of one buffer (no H2D)", lambda: step(v))
timeit("4 copies + step", lambda: (copy4(), step(dbuf)))
timeit("1 copy + step on views", lambda: (copy1(), step(v)))
