"""Dev tool: one rank's share of a global in-batch step on ONE GPU (the gathered blocked rows are built locally):
device time, phase stamps and tile counts of rank r of `world`."""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops, _lib
from scripts.quick_time import _ramp, stamps

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b_loc = 65536
d = G.cfg5(world, 0)
lay = ops.packed_block_layout(b_loc, 1, True, False)
gbuf = torch.empty(world * lay["stride"], dtype=torch.uint8, device="cuda")
for r in range(world):
    lo, hi = r * b_loc, (r + 1) * b_loc
    t = lambda k: torch.tensor(np.ascontiguousarray(d[k][lo:hi]), device="cuda")
    blk = ops.pack_row_block(t("g").reshape(1, -1), t("s"), t("y"), t("w"), None, lay["stride"])
    gbuf[r * lay["stride"]:(r + 1) * lay["stride"]] = blk
_ramp()
for persistent in (True, False):
    kw = dict(label_func="diff", power=-0.5, part=(rank, world), persistent=persistent)
    for _ in range(5):
        out = ops.pairwise_fwd_bwd_blocked(gbuf, world, b_loc, 1, True, False, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = ops.pairwise_fwd_bwd_blocked(gbuf, world, b_loc, 1, True, False, **kw)
    e1.record(); torch.cuda.synchronize()
    print(f"world {world} rank {rank} persistent={persistent}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/call, n_pair {int(out['n_pair'])}, "
          f"path {ops.last_segmentation_path(out['_scratch'])}")
    stamps(out)
