#!/bin/bash
# Final single-GPU session at HEAD: whole GPU suite, smoke, launch list + full ncu capture of the headline kernels (plain run first),
# profiles/inst_per_fit.json regenerated from that capture, then both bench arms (so that frac_executed comes from the same code)
cd "$(dirname "$0")/.."
T=${1:-r02z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.log 2>&1
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --sims-per-step 1048576"
$CMD > gpurun_out/${T}_plain_1mi.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ols_lean|ols_solve" -s 6 -c 2 -f -o gpurun_out/${T}_lean $CMD > gpurun_out/${T}_ncu_full.log 2>&1
ncu -i gpurun_out/${T}_lean.ncu-rep --page raw --csv > gpurun_out/${T}_lean_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${T}_lean.ncu-rep > gpurun_out/${T}_lean_summary.txt 2>&1
ncu -i gpurun_out/${T}_lean.ncu-rep --page source --csv > gpurun_out/${T}_lean_sass.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/${T}_lean_sass.csv 12 > gpurun_out/${T}_lean_hot.txt 2>&1
python tools/make_inst_per_fit.py gpurun_out/${T}_lean_raw.csv 1048576 "profiles/${T}_ncu_summary.md" > gpurun_out/${T}_inst_per_fit.log 2>&1
cp profiles/inst_per_fit.json gpurun_out/${T}_inst_per_fit.json
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/${T}_bench_ref.json
head -40 gpurun_out/${T}_lean_summary.txt
ls -la gpurun_out/ | grep ${T}
