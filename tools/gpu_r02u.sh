#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fuzz_parity.py -q -m gpu > gpurun_out/r02u_fuzz.log 2>&1; echo "rc=$?" >> gpurun_out/r02u_fuzz.log; tail -5 gpurun_out/r02u_fuzz.log | cut -c1-300
timeout 1500 python tools/fuzz_parity.py 72 672 > gpurun_out/r02u_fuzz600.log 2>&1; tail -8 gpurun_out/r02u_fuzz600.log | cut -c1-400
timeout 600 python - > gpurun_out/r02u_speed.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import mcpower_b200 as M
import _cases as CS
for name in ("lmm_crossed_a", "lmm_nested_b", "lmm_crossed_nested_a", "lmm_slope_crossed"):
    c = CS.validation_cases()[name]
    s = CS.case_spec(c)
    n = c["n"]
    M.run_batch(s, [n // 2, n], 512, 7)
    for sims in (4096, 16384):
        r = M.run_batch(s, [n // 2, n], sims, 7)
        print(name, sims, 'sims', round(r.kernel_ms, 1), 'ms', round(sims * 2 / (r.kernel_ms * 1e-3)), 'fits/s', 'conv', r.n_converged.tolist(), 'waves', r.n_waves, flush=True)
PY
cat gpurun_out/r02u_speed.log
