"""Dev probe: host time per submit and steady-state step time of the host-buffer front end."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200.host import HostPairwise
d = G.cfg3(); B = d["s"].size
hin = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in ("g", "s", "y", "w")}
for depth in (1, 2, 3):
    hp = HostPairwise(B, depth=depth)
    outs = [dict(loss=torch.empty(1).pin_memory(), n_pair_f32=torch.empty(1).pin_memory(),
                 n_pair=torch.empty(1, dtype=torch.int64).pin_memory(), dlogits=torch.empty(B).pin_memory()) for _ in range(depth)]
    bound = [hp.bind(hin["g"], hin["s"], hin["y"], rw_pos=hin["w"], label_func="diff", power=-0.5, **o) for o in outs]
    def run(n):
        tk = []; tsub = 0.0
        for k in range(n):
            t0 = time.perf_counter(); t = bound[k % depth].submit(); tsub += time.perf_counter() - t0
            tk.append(t)
            if len(tk) >= depth: hp.wait(tk.pop(0))
        for t in tk: hp.wait(t)
        return tsub / n
    run(50); torch.cuda.synchronize()
    t0 = time.perf_counter(); sub = run(300); dt = (time.perf_counter() - t0) / 300
    print(f"depth {depth}: {dt*1e6:.1f} us/step, submit host time {sub*1e6:.1f} us (includes waiting for a free slot)")
    hp.close()
