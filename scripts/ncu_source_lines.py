#!/usr/bin/env python
"""Aggregate the ncu source page per CUDA source line: instructions executed and stall samples.

    python scripts/ncu_source_lines.py report.ncu-rep [top_n] [sort: samples|inst]
"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "samples"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; rows = []; hdr = None
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit():
        try:
            rows.append((cur, int(r[0]), r[1].strip()[:84], int(r[7] or 0), int(r[6] or 0)))
        except ValueError:
            pass
ti = sum(x[3] for x in rows); ts = sum(x[4] for x in rows)
print(f"total inst {ti}  total samples {ts}")
k = 4 if key == "samples" else 3
for f, ln, src, ie, sm in sorted(rows, key=lambda x: -x[k])[:top]:
    print(f"{f}:{ln:4d} inst {ie:9d} ({100*ie/max(ti,1):4.1f}%) samples {sm:6d} ({100*sm/max(ts,1):4.1f}%)  {src}")
