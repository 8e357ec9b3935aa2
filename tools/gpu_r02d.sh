#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02d_tests.log
tail -5 gpurun_out/r02d_tests.log
timeout 900 python tools/bench_variants.py > gpurun_out/r02d_variants.log 2>&1
cat gpurun_out/r02d_variants.log
ncu --metrics smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:"ols_lean" -s 4 -c 1 --csv --log-file gpurun_out/r02d_lean_metrics.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --sims-per-step 1048576 > gpurun_out/r02d_ncu.log 2>&1
cat gpurun_out/r02d_lean_metrics.csv | tail -8
