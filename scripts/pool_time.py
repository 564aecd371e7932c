"""Quick device timing of the segment-pooling kernels on the bench's shape (dev tool, also the ncu target)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from rec_now_b200.rec_block import embedding_util as EU

rng = np.random.default_rng(0)
b, c, d, v = 65536, 32, 64, 1 << 20
slots = torch.tensor(rng.integers(0, 16, (b, c)).astype(np.int32), device="cuda")
ids = torch.tensor(rng.integers(0, v, (b, c)).astype(np.int64), device="cuda")
w = torch.tensor(rng.uniform(0.5, 1.5, (b, c)).astype(np.float32), device="cuda")
table = torch.randn((v, d), device="cuda", requires_grad=True)
targets = [1, 3, 4, 7, 8, 10, 13, 15]
for _ in range(3):
    out = EU.segment_pool(table, slots, targets, ids, w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = EU.segment_pool(table, slots, targets, ids, w)
e1.record(); torch.cuda.synchronize()
print(f"forward {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
g = torch.randn_like(out)
out.backward(g); torch.cuda.synchronize()
e0.record()
for _ in range(10):
    table.grad = None
    EU.segment_pool(table, slots, targets, ids, w).backward(g)
e1.record(); torch.cuda.synchronize()
print(f"forward + backward {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
