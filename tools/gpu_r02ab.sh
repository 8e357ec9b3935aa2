#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/fuzz_parity.py 72 772 > gpurun_out/r02ab_fuzz.log 2>&1
tail -6 gpurun_out/r02ab_fuzz.log | cut -c1-400
