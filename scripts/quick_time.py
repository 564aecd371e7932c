"""Quick device timing of rn_pairwise_fwd_bwd on the BASELINE configs (dev tool, not the bench)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops


def _ramp():
    """Clock ramp: ~0.5 s of SFU work so that the timed calls run at the boost clock."""
    import ctypes as C, time, torch
    from rec_now_b200 import _lib
    sink = torch.zeros(4, device="cuda"); n = C.c_int64(0)
    t = time.perf_counter() + 0.5
    while time.perf_counter() < t:
        _lib.lib().rn_bench_mufu(2000, sink.data_ptr(), C.byref(n), None); torch.cuda.synchronize()

def run(d, iters=200):
    _ramp()
    s, y = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda()
    keys = torch.tensor(d["g"]).cuda().reshape(1, -1)
    w = torch.tensor(d["w"]).cuda() if "w" in d else None
    import os
    kw = dict(label_func=os.environ.get("RN_QT_LABEL_FUNC", d["label_func"]), power=d["power"], rw_pos=w)      # (e.g. lambda)
    for _ in range(5):
        out = ops.pairwise_fwd_bwd(s, y, keys, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import os
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if os.environ.get("RN_QT_FLUSH") else None
    if flush is not None:                       # cold L2 as in bench.py: per-call events, the flush outside them
        ms = 0.0
        for k in range(iters):
            flush.fill_(k & 0xFF)
            e0.record(); out = ops.pairwise_fwd_bwd(s, y, keys, **kw); e1.record(); torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ms /= iters
    else:
        e0.record()
        for _ in range(iters):
            out = ops.pairwise_fwd_bwd(s, y, keys, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    if os.environ.get("RN_QT_GRAPH"):           # device time without the host: 50 calls captured in one CUDA graph, replayed
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(50):
                out = ops.pairwise_fwd_bwd(s, y, keys, **kw)
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"   graph replay (50 calls per graph): {e0.elapsed_time(e1) / 500 * 1e3:.2f} us/call")
    n = int(out["n_pair"].item())
    print(f"{d['name']}: B={s.numel()} n_pair={n} {ms*1e3:.1f} us/call  {n/ms/1e6:.2f} Gpairs/s  "
          f"SFU-frac(3 MUFU, 4.65e12/s)={3*n/(ms*1e-3)/4.65e12:.3f} loss={out['loss'].item():.6f} err={ops.device_error(out['_scratch'])}")
    from rec_now_b200 import _lib
    print('   graph launches on this thread:', _lib.lib().rn_debug_graph_launches())
    stamps(out)

def stamps(out):
    import ctypes as C
    from rec_now_b200 import _lib
    ts = (C.c_uint64 * 34)()
    _lib.lib().rn_debug_timestamps(out["_scratch"].data_ptr(), ts, 34, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    t = list(ts)
    t0 = t[0]
    print("   phase stamps (us from k_seg start):", " ".join(f"{i}:{(x - t0) / 1e3:.1f}" for i, x in enumerate(t[:24]) if x))
    d = t[24:32]
    print(f"   units={t[32] & 0xFFFFFFFF} C={t[32] >> 32} tiles={t[33]}")
    if d[4]:
        M = (1 << 64) - 1
        print(f"   debug: longest unit {d[0] >> 32} cyc (u={d[0] & 0xFFFFFFFF}); busy/warp-avg {d[1] / 4736:.0f} cyc; "
              f"loop exit first {((~d[3] & M) - t0) / 1e3:.1f} us last {(d[2] - t0) / 1e3:.1f} us; units {d[4]} fast tiles {d[5]} "
              f"general {d[6]} general cyc/tile {d[7] / max(d[6], 1):.0f}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg3"]
    for name in which:
        run(getattr(G, name)())
