#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_random_slope.py tests/test_lme_knobs.py -x -q -m gpu > gpurun_out/r02q_slope.log 2>&1; echo "slope rc=$?" >> gpurun_out/r02q_slope.log
tail -4 gpurun_out/r02q_slope.log
for s in glm lmm c3; do MCP_VARIANT_SET=$s timeout 600 python tools/bench_variants.py 2097152 > gpurun_out/r02q_$s.log 2>&1; cat gpurun_out/r02q_$s.log; done
timeout 300 python - > gpurun_out/r02q_slope_speed.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import mcpower_b200 as M
from test_random_slope import slope_spec
for label, s in (("one slope", slope_spec(30, 0.2, 0.1, 0.3)), ("two slopes", slope_spec(30, 0.2, 0.1, 0.3, second=(1, 0.08, 0.1, 0.2)))):
    M.run_batch(s, [300, 600], 2000, 7)
    for n in (20000, 100000):
        r = M.run_batch(s, [300, 600], n, 7)
        print(label, n, 'sims', round(r.kernel_ms, 1), 'ms', round(n * 2 / (r.kernel_ms * 1e-3)), 'fits/s', 'conv', r.n_converged.tolist())
PY
cat gpurun_out/r02q_slope_speed.log
