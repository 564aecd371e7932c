"""Dev tool: per-warp timeline of the pair kernel (RN_PAIR_DEBUG=1).  Usage: RN_PAIR_DEBUG=1 python scripts/pair_debug.py cfg3"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200 import ops, _lib

d = getattr(G, sys.argv[1] if len(sys.argv) > 1 else "cfg3")()
s, y = torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda()
keys = torch.tensor(d["g"]).cuda().reshape(1, -1)
w = torch.tensor(d["w"]).cuda() if "w" in d else None
kw = dict(label_func=d["label_func"], power=d["power"], rw_pos=w)

def _ramp():
    """Clock ramp: ~0.5 s of SFU work so that the timed calls run at the boost clock."""
    import ctypes as C, time, torch
    from rec_now_b200 import _lib
    sink = torch.zeros(4, device="cuda"); n = C.c_int64(0)
    t = time.perf_counter() + 0.5
    while time.perf_counter() < t:
        _lib.lib().rn_bench_mufu(2000, sink.data_ptr(), C.byref(n), None); torch.cuda.synchronize()
_ramp()
for _ in range(50):
    out = ops.pairwise_fwd_bwd(s, y, keys, **kw)
torch.cuda.synchronize()
scr = out["_scratch"]
ts = (C.c_uint64 * 34)()
_lib.lib().rn_debug_timestamps(scr.data_ptr(), ts, 34, None)
t = list(ts)
# locate gstat: the arena layout is deterministic; find it by scanning for the record pattern instead of
# duplicating the layout here: records start at the gstat offset = total - align(8*4*B)
B = s.numel()
NW = int(__import__('os').environ.get('NW', '16'))
total = scr.numel()
gstat_off = _lib.lib().rn_debug_arena_offset(B, 1, 0)
rec = scr[gstat_off:gstat_off + 148 * NW * 64].view(torch.int64).cpu().numpy().reshape(-1, 8)
t20 = t[20]
start = (rec[:, 0] - t20) / 1e3; end = (rec[:, 1] - t20) / 1e3
busy = rec[:, 2]; units = rec[:, 3] & 0xFFFFFFFF; gen = rec[:, 3] >> 32
ok = rec[:, 1] > 0
print(f"warps {ok.sum()}  k_pair start->barrier {(t[22]-t20)/1e3:.1f} us; fin {(t[23]-t[22])/1e3:.1f} us")
print(f"first-unit start: min {start[ok].min():.1f} med {np.median(start[ok]):.1f} max {start[ok].max():.1f} us")
print(f"loop exit:        min {end[ok].min():.1f} p10 {np.percentile(end[ok],10):.1f} med {np.median(end[ok]):.1f} p90 {np.percentile(end[ok],90):.1f} max {end[ok].max():.1f} us")
sm_end = end.reshape(148, NW).max(1); sm_busy = busy.reshape(148, NW).sum(1)
print(f"per-SM last exit: min {sm_end.min():.1f} med {np.median(sm_end):.1f} max {sm_end.max():.1f} us")
print(f"per-SM busy warp-cycles: min {sm_busy.min()/1e3:.0f}K med {np.median(sm_busy)/1e3:.0f}K max {sm_busy.max()/1e3:.0f}K")
print(f"units/warp: min {units[ok].min()} med {np.median(units[ok])} max {units[ok].max()}; general tiles/warp max {gen[ok].max()}")
o = np.argsort(-end)[:5]
for i in o: print(f"  slow warp cta {i//NW} w {i%NW}: start {start[i]:.1f} end {end[i]:.1f} busy {busy[i]} units {units[i]} gen {gen[i]}")
sm_units = units.reshape(148, NW).sum(1); sm_gen = gen.reshape(148, NW).sum(1)
o = np.argsort(sm_end)
print("per-SM (sorted by last exit): sm exit busyK units gen | warp exits sorted")
for i in list(o[:3]) + list(o[-4:]):
    we = np.sort(end.reshape(148, NW)[i])
    print(f"  sm {i:3d} exit {sm_end[i]:.1f} busy {sm_busy[i]/1e3:.0f}K units {sm_units[i]} gen {sm_gen[i]} | " + " ".join(f"{x:.0f}" for x in we[::max(1,NW//8)]) + f" {we[-1]:.0f}")
print("corr(exit, busy) =", np.corrcoef(sm_end, sm_busy)[0,1], " corr(exit, gen) =", np.corrcoef(sm_end, sm_gen)[0,1])

fast = rec[:, 4]; fcyc = rec[:, 5]; gcyc = rec[:, 6]; eig = rec[:, 7]
print(f"eighths/warp: min {eig[ok].min()} med {np.median(eig[ok])} max {eig[ok].max()}; fast tiles/warp med {np.median(fast[ok])} max {fast[ok].max()}")
dur = (end - start)
for name, m in (("gen==0", ok & (gen == 0)), ("gen>=1", ok & (gen >= 1)), ("gen>=3", ok & (gen >= 3))):
    if m.sum(): print(f"  {name}: n {m.sum()} dur med {np.median(dur[m]):.1f} p90 {np.percentile(dur[m],90):.1f} max {dur[m].max():.1f} us; eighths med {np.median(eig[m])}; busy med {np.median(busy[m])}; fastcyc/fast-eighth med {np.median(fcyc[m]/np.maximum(eig[m]-0,1)):.0f}")
print("corr(dur, eighths) =", np.corrcoef(dur[ok], eig[ok])[0,1], " corr(dur, gen) =", np.corrcoef(dur[ok], gen[ok])[0,1], " corr(dur, units) =", np.corrcoef(dur[ok], units[ok])[0,1])
wq = np.arange(len(dur)) % NW
print("dur by warp%4 (SMSP):", [round(float(np.median(dur[ok & (wq % 4 == k)])), 1) for k in range(4)])
print("dur by warp idx/8:", [round(float(np.median(dur[ok & (wq // 8 == k)])), 1) for k in range(NW // 8)])
cta = np.arange(len(dur)) // NW
o = np.argsort(-dur)[:12]
for i in o: print(f"  slow warp cta {cta[i]} w {wq[i]}: dur {dur[i]:.1f} busy {busy[i]} segs {units[i]} gen {gen[i]} fast {fast[i]} eighths {eig[i]} fastcyc {fcyc[i]} gencyc {gcyc[i]}")
o = np.argsort(dur)[:6]
for i in o: print(f"  quick warp cta {cta[i]} w {wq[i]}: dur {dur[i]:.1f} busy {busy[i]} segs {units[i]} gen {gen[i]} fast {fast[i]} eighths {eig[i]} fastcyc {fcyc[i]} gencyc {gcyc[i]}")
np.savez("gpurun_out/pair_debug.npz", rec=rec, t20=t20, ts=np.array(t, dtype=np.uint64))
