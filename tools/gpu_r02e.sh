#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wide_designs.py tests/test_gpu_hostpath.py tests/test_fit_provided_data.py -m gpu -x -q > gpurun_out/r02e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02e_tests.log
tail -15 gpurun_out/r02e_tests.log
