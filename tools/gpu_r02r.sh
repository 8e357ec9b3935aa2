#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_fuzz_parity.py -q -m gpu > gpurun_out/r02r_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r02r_fuzz.log
tail -40 gpurun_out/r02r_fuzz.log | cut -c1-300
