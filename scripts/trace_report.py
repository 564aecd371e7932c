"""Dev tool: summarise a pair-kernel unit trace (scripts/pair_trace.py)."""
import sys
import numpy as np
z = np.load(sys.argv[1])
ts = z['ts'].astype(np.int64); tr = z['trace']
t20 = ts[20]
st = (tr[:, :, 0] - t20) / 1e3; en = (tr[:, :, 1] - t20) / 1e3
valid = (tr[:, :, 0] > t20) & (st < 200)
nt = (tr[:, :, 2] >> 16) & 0xFFFF; ngen = tr[:, :, 2] & 0xFFFF; uid = tr[:, :, 2] >> 32
cyc = tr[:, :, 3] >> 32; tcyc = tr[:, :, 3] & 0xFFFFFFFF
pA = tr[:, :, 4] >> 32; pB = tr[:, :, 4] & 0xFFFFFFFF; pD = tr[:, :, 5] >> 32; pE = tr[:, :, 5] & 0xFFFFFFFF
print("k_pair stamps: warp0 exit %.1f barrier %.1f end %.1f; valid units %d" % ((ts[21] - t20) / 1e3, (ts[22] - t20) / 1e3, (ts[23] - t20) / 1e3, valid.sum()))
edges = [0, 2, 4, 8, 12, 16, 20, 24, 26, 28, 30, 34, 100]
for lo, hi in zip(edges[:-1], edges[1:]):
    m = valid & (st >= lo) & (st < hi)
    if not m.sum(): continue
    print(f"start in [{lo:3d},{hi:3d}) n {m.sum():5d} uid med {np.median(uid[m]):6.0f} cyc {np.median(cyc[m]):6.0f} tile {np.median(tcyc[m]):6.0f} | A {np.median(pA[m]):5.0f} B {np.median(pB[m]):5.0f} D {np.median(pD[m]):4.0f} E {np.median(pE[m]):5.0f} | gen/unit {ngen[m].mean():.2f}")
for t in range(0, 56, 4):
    print(f"t={t:2d}: running {((st <= t) & (en > t) & valid).sum()}", end="; ")
print()
