set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null
timeout 200 python bench.py --quick --steps 3 --warmup 1 --products 1 --factors 3 --n-vars 28 --gkr-log-inputs 0 --gkr-uniform-log-gates 0 > gpurun_out/target_d3_final.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --quick --steps 2 --warmup 1 --gkr-log-inputs 0 --gkr-uniform-log-gates 0 > gpurun_out/ncu_final.log 2>&1
METRICS=$(grep -o -- "--metrics [^ ]*" tools/ncu_all_kernels.sh | head -1 | cut -d" " -f2)
ALLK_GKR_ONLY=1 timeout 400 ncu --clock-control none --kernel-name-base demangled --metrics $METRICS -k "regex:k_gkr_|k_layer_eval|k_eq_split" -o /tmp/allk_b -f python tools/all_kernels.py 22 > gpurun_out/ncu_allk_b.log 2>&1
ncu -i /tmp/allk_b.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py --longest "GKR kernels (reference wiring and general wiring): longest launch per kernel, metric subset, layers 2^20 wide" > gpurun_out/allk_b_summary.csv
ls -la gpurun_out | tail -12
