#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text file for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_kpair.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_xu.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of {rep} (ncu --set full --clock-control none); one block per captured launch"]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        lines.append(f"\n== {r[name_col][:110]}")
        d = dict(zip(hdr, zip(units, r)))
        for k in KEYS:
            if k in d:
                lines.append(f"{k} [{d[k][0]}] = {d[k][1]}")
        st = sorted(((float(v[1].replace(',', '') or 0), k[len(STALL):-len('_per_issue_active.ratio')])
                     for k, v in d.items() if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")), reverse=True)
        lines.append("stalls per issue: " + ", ".join(f"{n}={x:.2f}" for x, n in st[:8]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
