#!/usr/bin/env python3
"""Developer tool: BASELINE configs[0] — OLS y ~ x1 + x2 + x1:x2, n = 100, 1600 sims, alpha 0.05 — as ONE default-size call through
the public API with a ProgressSink attached (what the reference's Python API does): median wall clock per call on the GPU next
to the CPU restatement on all host threads.  Prints one JSON object."""
import json
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import mcpower_b200 as M  # noqa: E402
from mcpower_b200 import spec as S  # noqa: E402
from oracle import orc  # noqa: E402  (the CPU side of the comparison: a bench leg, not the product)

spec = S.add_interaction(S.base_spec(2), [1, 2], 0.2)
spec["effect_sizes"][1], spec["effect_sizes"][2] = 0.3, 0.25
seed = M.splitmix64(2137)
sink = lambda cur, tot: True  # noqa: E731
for _ in range(20):
    r = M.run_batch(spec, [100], 1600, seed, sink)
gpu = []
for _ in range(300):
    t0 = time.perf_counter()
    r = M.run_batch(spec, [100], 1600, seed, sink)
    gpu.append(time.perf_counter() - t0)
cpu = []
for _ in range(30):
    t0 = time.perf_counter()
    w = orc.run_batch(spec, [100], 1600, seed, n_threads=orc.host_threads())
    cpu.append(time.perf_counter() - t0)
print(json.dumps({"workload": "config 1: y ~ x1 + x2 + x1:x2, n=100, 1600 sims, one call, sink attached",
                  "gpu_call_us_median": 1e6 * statistics.median(gpu), "gpu_call_us_p90": 1e6 * sorted(gpu)[int(0.9 * len(gpu))],
                  "gpu_kernel_us": 1e3 * r.kernel_ms, "gpu_launches": r.gpu_launches,
                  "cpu_call_us_median": 1e6 * statistics.median(cpu), "cpu_threads": orc.host_threads(),
                  "power_gpu": [float(v) for v in r.power()[0]], "identical_counters": bool((r.counters[0, :3] == M.counters_from_bytes(w["uncorrected"], w["corrected"], w["converged"], w["overall"], w["factor_excluded"])[0, :3]).all())}))
