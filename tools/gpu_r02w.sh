#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02w_smi.log
timeout 900 python -m pytest tests/test_gpu_hostpath.py tests/test_random_slope.py tests/test_mixed_groupings.py -q -m gpu > gpurun_out/r02w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_tests.log; tail -4 gpurun_out/r02w_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02w_bench_n2.json 2> gpurun_out/r02w_bench_n2.err; echo "bench2 rc=$?"; tail -c 400 gpurun_out/r02w_bench_n2.json
