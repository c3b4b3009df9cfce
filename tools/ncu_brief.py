#!/usr/bin/env python
"""Brief view of an `ncu --page raw --csv` export: one block per kernel launch with the metrics the round kernels are
judged by (duration, DRAM bytes, multiplier-pipe utilisation, issue utilisation, stall reasons > 0.2 per issue).
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_brief.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("-----", d.get("Kernel Name", "?")[:90])
    for k in KEYS:
        if k in d:
            print(f"  {k}: {d[k]}")
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                v = float(d[k].replace(",", ""))
            except ValueError:
                continue
            if v >= 0.2:
                print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]}: {v:.2f}")
