#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/fuzz_parity.py 72 1072 > gpurun_out/r02s_fuzz.log 2>&1
tail -30 gpurun_out/r02s_fuzz.log | cut -c1-400
