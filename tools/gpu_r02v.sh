#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mixed_groupings.py tests/test_random_slope.py tests/test_fit_provided_data.py -q -m gpu > gpurun_out/r02v_mixed.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_mixed.log; tail -12 gpurun_out/r02v_mixed.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_fuzz_parity.py -q -m gpu > gpurun_out/r02v_fuzz.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_fuzz.log; tail -5 gpurun_out/r02v_fuzz.log | cut -c1-300
timeout 600 python tools/fuzz_parity.py 72 372 > gpurun_out/r02v_fuzz300.log 2>&1; tail -5 gpurun_out/r02v_fuzz300.log | cut -c1-400
