#!/bin/bash
# ncu source-level capture of the config-3 streaming kernel (doomer sweep) + the design kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/run_workload.py c3 65536 > gpurun_out/c3_c3_plain.log 2>&1; cat gpurun_out/c3_c3_plain.log
ncu --set full --import-source on --clock-control none -k regex:"ols_power_kernel_q4|scen_design" -s 2 -c 2 -o gpurun_out/c3_c3 -f python tools/run_workload.py c3 65536 > gpurun_out/c3_ncu.log 2>&1
tail -3 gpurun_out/c3_ncu.log
ls -la gpurun_out/c3_c3.ncu-rep
ncu -i gpurun_out/c3_c3.ncu-rep --page source --csv --print-source cuda > gpurun_out/c3_c3_cuda.csv 2>/dev/null
ncu -i gpurun_out/c3_c3.ncu-rep --page source --csv > gpurun_out/c3_c3_sass.csv 2>/dev/null
ncu -i gpurun_out/c3_c3.ncu-rep --page raw --csv > gpurun_out/c3_c3_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep r02n
