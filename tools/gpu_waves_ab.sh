#!/bin/bash
# A/B of the BYTES wave rule on a multi-GPU box: the in-tree library against build_variants/<name> (strong scaling of one call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in main waves128; do
  so=$PWD/mcpower_b200/libmcpower_b200.so; [ $v != main ] && so=$PWD/build_variants/$v/libmcpower_b200.so
  MCPOWER_B200_LIB=$so timeout 600 python tools/strong_scaling.py > gpurun_out/waves_ab_$v.json 2> gpurun_out/waves_ab_$v.err
  echo "== $v"; grep "bytes" gpurun_out/waves_ab_$v.err | sed 's/"lower_ms.*"waves"/"waves"/' | cut -c1-220
done
timeout 600 python -m pytest tests/test_gpu_hostpath.py tests/test_gpu_full_size.py -q -m gpu > gpurun_out/waves_ab_tests.log 2>&1; tail -3 gpurun_out/waves_ab_tests.log
