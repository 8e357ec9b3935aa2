#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_scenario_goldens.py tests/test_lme_knobs.py tests/test_wide_designs.py -x -q -m gpu > gpurun_out/r02l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02l_tests.log
tail -8 gpurun_out/r02l_tests.log
MCP_VARIANT_SET=c3 timeout 600 python tools/bench_variants.py 2097152 > gpurun_out/r02l_c3_variants.log 2>&1
cat gpurun_out/r02l_c3_variants.log
MCP_VARIANT_SET=glm timeout 600 python tools/bench_variants.py 2097152 > gpurun_out/r02l_glm_variants.log 2>&1
cat gpurun_out/r02l_glm_variants.log
