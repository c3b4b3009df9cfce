#!/bin/bash
# Nsight Compute evidence for every kernel (run on the GPU box; only the CSV summaries are kept).
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 700 ncu --set full --clock-control none --kernel-name-base demangled --kernel-id ::regex:k_:1 -o /tmp/allk_a -f python tools/all_kernels.py 22 > gpurun_out/ncu_allk_a.log 2>&1
ncu -i /tmp/allk_a.ncu-rep --page raw --csv > /tmp/allk_a.csv 2>/dev/null
ALLK_GKR_ONLY=1 timeout 700 ncu --clock-control none --kernel-name-base demangled --metrics launch__grid_size,launch__block_size,launch__registers_per_thread,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio -k "regex:k_gkr_|k_layer_eval|k_eq_split" -o /tmp/allk_b -f python tools/all_kernels.py 22 > gpurun_out/ncu_allk_b.log 2>&1
ncu -i /tmp/allk_b.ncu-rep --page raw --csv > /tmp/allk_b.csv 2>/dev/null
python tools/ncu_summary.py "first (largest) launch of every kernel instantiation, ncu --set full" < /tmp/allk_a.csv > gpurun_out/allk_a_summary.csv
python tools/ncu_summary.py --longest "GKR kernels: longest launch per kernel, metric subset" < /tmp/allk_b.csv > gpurun_out/allk_b_summary.csv
ls -la /tmp/allk_* gpurun_out/
