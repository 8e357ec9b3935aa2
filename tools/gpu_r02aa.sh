#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hostpath.py tests/test_mixed_groupings.py tests/test_random_slope.py -q -m gpu > gpurun_out/r02aa.log 2>&1; echo "rc=$?" >> gpurun_out/r02aa.log; tail -12 gpurun_out/r02aa.log | cut -c1-250
