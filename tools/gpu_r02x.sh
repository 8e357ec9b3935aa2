#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mixed_groupings.py tests/test_random_slope.py tests/test_lme_knobs.py tests/test_wide_designs.py tests/test_fit_provided_data.py -q -m gpu > gpurun_out/r02x_lmm.log 2>&1; echo "rc=$?" >> gpurun_out/r02x_lmm.log; tail -4 gpurun_out/r02x_lmm.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "lmm or every_template_width" > gpurun_out/r02x_par.log 2>&1; tail -2 gpurun_out/r02x_par.log
