"""Dev tool: per-instruction view of an ncu SASS source page (csv): executed count, samples, top stalls.
    ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv ; python scripts/ncu_sass_hot.py x.csv [min_exec]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
min_exec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_exec = 0; tot_samp = 0
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    ex = int(r[ix["Instructions Executed"]] or 0); sm = int(r[ix["# Samples"]] or 0)
    tot_exec += ex; tot_samp += sm
    if ex < min_exec: continue
    st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{n:5d} {ex:9d} {sm:6d}  {r[ix['Source']].strip()[:90]:90s} " + " ".join(f"{k}={v}" for v, k in st if v))
print("total exec", tot_exec, "samples", tot_samp)
