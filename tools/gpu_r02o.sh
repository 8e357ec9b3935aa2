#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_random_slope.py -x -q -m gpu > gpurun_out/r02o_slope.log 2>&1; echo "slope rc=$?" >> gpurun_out/r02o_slope.log
tail -30 gpurun_out/r02o_slope.log
timeout 2400 python -m pytest tests -q -m gpu --deselect tests/test_random_slope.py > gpurun_out/r02o_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r02o_all.log
tail -15 gpurun_out/r02o_all.log
for s in c3 glm lmm; do MCP_VARIANT_SET=$s timeout 600 python tools/bench_variants.py 2097152 > gpurun_out/r02o_$s.log 2>&1; cat gpurun_out/r02o_$s.log; done
