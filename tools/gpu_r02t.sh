#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_mixed_groupings.py -x -q -m gpu > gpurun_out/r02t_groupings.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_groupings.log
tail -40 gpurun_out/r02t_groupings.log | cut -c1-250
