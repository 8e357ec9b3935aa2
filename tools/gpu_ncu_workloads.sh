#!/bin/bash
# ncu key metrics + launch lists of the secondary workloads (BASELINE configs 3-5) at HEAD
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,launch__registers_per_thread,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for w in c3 c4 c5; do
  python tools/run_workload.py $w > gpurun_out/wl_${w}_plain.log 2>&1
  ncu --metrics $M --clock-control none -s 3 -c 6 --csv --log-file gpurun_out/wl_${w}_metrics.csv python tools/run_workload.py $w > gpurun_out/wl_${w}_ncu.log 2>&1
  cat gpurun_out/wl_${w}_plain.log
done
ls -la gpurun_out | grep wl_
