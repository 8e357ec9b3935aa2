#!/bin/bash
# 8-GPU session: in-library multi-device tests, strong scaling of one call, bench with the in-library fan-out
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02zz_smi.log
timeout 600 python -m pytest tests/test_gpu_hostpath.py -m gpu -x -q > gpurun_out/r02zz_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02zz_tests.log; tail -3 gpurun_out/r02zz_tests.log
timeout 900 python tools/strong_scaling.py > gpurun_out/r02zz_strong_scaling.json 2> gpurun_out/r02zz_strong_scaling.err
echo "strong rc=$?"; tail -3 gpurun_out/r02zz_strong_scaling.err | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02zz_bench.log 2> gpurun_out/r02zz_bench.err
echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02zz_bench_n8.log 2> gpurun_out/r02zz_bench_n8.err
echo "bench8 rc=$?"; tail -c 300 gpurun_out/r02zz_bench_n8.log
