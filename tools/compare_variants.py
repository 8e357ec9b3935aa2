#!/usr/bin/env python3
"""Device-time fits/s of one OLS workload under each forced kernel variant (mcpower_b200.force_ols_variant):
usage compare_variants.py c1|c2|c3|c3r|hl [n_sims] [variants, e.g. 1,4,6]."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mcpower_b200 as M  # noqa: E402
from mcpower_b200 import spec as S  # noqa: E402

which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["1", "4", "6"]
seed = M.splitmix64(2137)
grid50 = list(range(20, 1001, 20))
if which == "c3":
    spec, grid = dict(S.sweep_spec(), scenario=S.scenario("doomer"), crit_values={"alpha": 5e-7, "posthoc_alpha": None}), grid50
elif which == "c3r":
    spec, grid = dict(S.sweep_spec(), scenario=S.scenario("realistic")), grid50
elif which == "c2":
    spec, grid = S.sweep_spec(), grid50
elif which == "c1":
    spec = S.add_interaction(S.base_spec(2), [1, 2], 0.2)
    spec["effect_sizes"][1], spec["effect_sizes"][2] = 0.3, 0.2
    spec["target_indices"] = [1, 2, 3]
    grid = [100]
else:
    spec, grid = S.headline_spec(), [500]
ref = None
for v in variants:
    M.force_ols_variant(int(v))
    M.run_batch(spec, grid, 4096, seed)
    best = None
    for _ in range(3):
        r = M.run_batch(spec, grid, n, seed)
        best = r.kernel_ms if best is None else min(best, r.kernel_ms)
    same = "" if ref is None else ("  counters identical" if (r.counters == ref).all() else "  COUNTERS DIFFER")
    if ref is None:
        ref = r.counters
    print(f"{which} variant {v}: {n * len(grid) / (best * 1e-3):.4e} fits/s  ({best:.2f} ms, launches {r.gpu_launches}){same}", flush=True)
