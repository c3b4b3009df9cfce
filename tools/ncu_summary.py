#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the columns DESIGN.md quotes (one row per captured launch).
usage: ncu -i prof.ncu-rep --page raw --csv | python tools/ncu_summary.py "header comment" > profiles/xxx.csv"""
import csv, sys
COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
hdr, units = rows[0], rows[1]
body = [r for r in rows[2:] if r != hdr and r != units]  # several reports may be concatenated
if "--longest" in sys.argv:  # keep the longest launch of each kernel name
    sys.argv.remove("--longest")
    ti, ni, best = hdr.index("gpu__time_duration.sum"), hdr.index("Kernel Name"), {}
    for r in body:
        if r[ni] not in best or float(r[ti].replace(",", "")) > float(best[r[ni]][ti].replace(",", "")):
            best[r[ni]] = r
    body = list(best.values())
idx = [hdr.index(c) if c in hdr else -1 for c in COLS]
if len(sys.argv) > 1:
    print("# " + sys.argv[1])
w = csv.writer(sys.stdout)
w.writerow(COLS + ["dram_GBps"])
w.writerow([units[i] if i >= 0 else "" for i in idx] + ["GB/s"])
for r in body:
    vals = [r[i] if i >= 0 else "" for i in idx]
    try:
        f = lambda s: float(s.replace(",", ""))
        t = f(vals[4]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3}.get(units[idx[4]], 1e-9)
        sc = lambda i: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[i]], 1.0)
        gbps = (f(vals[5]) * sc(5) + f(vals[6]) * sc(6)) / t / 1e9
        vals.append("%.1f" % gbps)
    except Exception:
        vals.append("")
    w.writerow(vals)
