#!/bin/bash
# Round-2b GPU session: host-path refactor (per-device contexts, progress mailbox, pipelined BYTES, in-library multi-GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02b_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02b_tests.log
tail -5 gpurun_out/r02b_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench.log 2> gpurun_out/r02b_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r02b_bench.err
