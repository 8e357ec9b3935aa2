#!/bin/bash
mkdir -p gpurun_out
MCP_VARIANT_SET=c3 timeout 600 python tools/bench_variants.py 2097152 > gpurun_out/r02m_c3_variants.log 2>&1
cat gpurun_out/r02m_c3_variants.log
