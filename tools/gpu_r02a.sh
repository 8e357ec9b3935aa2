#!/bin/bash
# Round-2a GPU session: HEAD re-capture of the headline kernels (launch list + full ncu + source page), new pipe benches.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02a_bench.log 2> gpurun_out/r02a_bench.err || exit 1
python - > gpurun_out/r02a_pipe.log 2>&1 <<'PY'
import mcpower_b200 as M
for k in list(range(0, 20)):
    try:
        print(k, M.pipe_bench(k, 0), flush=True)
    except Exception as e:
        print(k, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02a_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --sims-per-step 1048576 > gpurun_out/r02a_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ols_lean|ols_solve" -s 6 -c 2 -f -o gpurun_out/r02a_lean \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --sims-per-step 1048576 > gpurun_out/r02a_ncu_full.log 2>&1
ls -la gpurun_out/r02a*
