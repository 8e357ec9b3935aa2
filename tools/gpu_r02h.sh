#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/bench_variants.py > gpurun_out/r02h_variants.log 2>&1
cat gpurun_out/r02h_variants.log
ncu --query-metrics 2>/dev/null | grep -i "pipe_fma\|pipe_alu\|pipe_xu\|pipe_fp64\|dispatch" | head -60 > gpurun_out/r02h_metric_names.log
M=sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum
for v in main v_half; do
  so=mcpower_b200/libmcpower_b200.so; [ $v != main ] && so=build_variants/$v/libmcpower_b200.so
  MCPOWER_B200_LIB=$PWD/$so ncu --metrics $M --clock-control none -k regex:"ols_lean" -s 4 -c 1 --csv --log-file gpurun_out/r02h_pipes_$v.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --sims-per-step 1048576 > gpurun_out/r02h_ncu_$v.log 2>&1
done
# config 2 launch lists, COUNTS and BYTES
cat > /tmp/c2.py <<'PY'
import sys, mcpower_b200 as M
from mcpower_b200 import spec as S
mode = M.MODE_COUNTS | (M.MODE_BYTES if sys.argv[1] == "bytes" else 0)
seed = M.splitmix64(2137)
grid = list(range(20, 1001, 20))
M.run_batch(S.sweep_spec(), grid, 4096, seed, mode=mode)
r = M.run_batch(S.sweep_spec(), grid, 1000000, seed, mode=mode)
print(sys.argv[1], r.kernel_ms, r.n_waves, r.gpu_launches)
PY
for m in counts bytes; do
  python /tmp/c2.py $m > gpurun_out/r02h_c2_$m.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02h_launches_c2_$m.csv python /tmp/c2.py $m > /dev/null 2>&1
done
cat gpurun_out/r02h_c2_*.log
