#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02f_tests.log
tail -4 gpurun_out/r02f_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02f_bench.log 2> gpurun_out/r02f_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02f_bench.err
