#!/bin/bash
# r02j: random-slope parity + the whole GPU suite at this HEAD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_random_slope.py -x -q -m gpu > gpurun_out/r02j_slope.log 2>&1; echo "slope rc=$?" >> gpurun_out/r02j_slope.log
tail -30 gpurun_out/r02j_slope.log
timeout 2400 python -m pytest tests -q -m gpu --deselect tests/test_random_slope.py > gpurun_out/r02j_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r02j_all.log
tail -15 gpurun_out/r02j_all.log
timeout 300 python - > gpurun_out/r02j_slope_speed.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import mcpower_b200 as M
from test_random_slope import slope_spec
s = slope_spec(30, 0.2, 0.1, 0.3)
for n in (2000, 20000, 100000):
    t = time.time(); r = M.run_batch(s, [300, 600], n, 7); dt = time.time() - t
    print(n, 'sims', round(dt * 1e3, 1), 'ms', round(n * 2 / dt), 'fits/s', 'conv', r.n_converged.tolist(), 'launches', r.gpu_launches)
PY
cat gpurun_out/r02j_slope_speed.log
