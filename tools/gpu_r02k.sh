#!/bin/bash
mkdir -p gpurun_out
MCP_VARIANT_SET=glm timeout 1200 python tools/bench_variants.py 2097152 > gpurun_out/r02k_glm_variants.log 2>&1
cat gpurun_out/r02k_glm_variants.log
