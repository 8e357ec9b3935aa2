#!/usr/bin/env python3
"""Run one of the BASELINE.json secondary workloads once (for ncu): usage run_workload.py c3|c4|c5 [n_sims]."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mcpower_b200 as M  # noqa: E402
from mcpower_b200 import spec as S  # noqa: E402

which = sys.argv[1]
seed = M.splitmix64(2137)
grid50 = list(range(20, 1001, 20))
if which == "c3":
    spec, grid, n = dict(S.sweep_spec(), scenario=S.scenario("doomer"), crit_values={"alpha": 5e-7, "posthoc_alpha": None}), grid50, 1 << 16
elif which == "c4":
    spec, grid, n = S.logit_spec(), [50, 100, 200, 500, 1000, 2000], 1 << 15
else:
    spec, grid, n = S.lmm_spec(30, 0.2), [60, 120, 300, 600, 900], 1 << 17
if len(sys.argv) > 2:
    n = int(sys.argv[2])
M.run_batch(spec, grid, 2048, seed)
r = M.run_batch(spec, grid, n, seed)
print(which, n, len(grid), r.kernel_ms, n * len(grid) / (r.kernel_ms * 1e-3))
