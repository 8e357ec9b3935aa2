#!/usr/bin/env python3
"""Developer tool: device-time fits/s of the headline shape (and the config-2 sweep) for every library variant under
build_variants/ (tools/build_variant.sh) plus the in-tree library; every variant must produce the same counters.
usage: [MCP_VARIANT_SET=ols|glm|c3|lmm] bench_variants.py [n_sims]   — each variant runs in its own process (MCPOWER_B200_LIB)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import json, sys
sys.path.insert(0, %r)
import mcpower_b200 as M
from mcpower_b200 import spec as S
n = int(sys.argv[1])
seed = M.splitmix64(2137)
out = {}
import os
SETS = {"ols": (("headline", S.headline_spec(), [500], n), ("c2_sweep", S.sweep_spec(), list(range(20, 1001, 20)), n // 8),
                ("c1", S.add_interaction(S.base_spec(2), [1, 2], 0.2), [100], n),
                ("mix9", dict(S.headline_spec(), correlation=[0.3 ** abs(i - j) for j in range(8) for i in range(8)]), [500], n // 2)),
        "glm": (("c4_logit", S.logit_spec(), [50, 100, 200, 500, 1000, 2000], n // 16),
                ("logit_n300", S.logit_spec(), [300], n // 4)),
        "c3": (("c3_doomer", dict(S.sweep_spec(), scenario=S.scenario("doomer"), crit_values={"alpha": 5e-7, "posthoc_alpha": None}),
                list(range(20, 1001, 20)), n // 8),
               ("c3_realistic", dict(S.sweep_spec(), scenario=S.scenario("realistic")), list(range(20, 1001, 20)), n // 8)),
        "lmm": (("c5_lmm", S.lmm_spec(30, 0.2), [60, 120, 300, 600, 900], n // 4),)}
for name, spec, grid, ns in SETS[os.environ.get("MCP_VARIANT_SET", "ols")]:
    M.run_batch(spec, grid, 4096, seed)
    best, cnt = None, None
    for _ in range(3):
        r = M.run_batch(spec, grid, ns, seed)
        best = r.kernel_ms if best is None else min(best, r.kernel_ms)
        cnt = r.counters
    out[name] = {"fits_per_s": ns * len(grid) / (best * 1e-3), "ms": best, "checksum": int(cnt.sum()), "first": [int(v) for v in cnt[0, :4]]}
print(json.dumps(out))
'''


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
    libs = {"main": ROOT / "mcpower_b200" / "libmcpower_b200.so"}
    for d in sorted((ROOT / "build_variants").glob("*/libmcpower_b200.so")):
        libs[d.parent.name] = d
    ref = None
    for name, so in libs.items():
        env = dict(os.environ, MCPOWER_B200_LIB=str(so))
        res = subprocess.run([sys.executable, "-c", CHILD % str(ROOT), str(n)], env=env, capture_output=True, text=True)
        if res.returncode != 0:
            print(name, "FAILED", res.stderr[-400:], flush=True)
            continue
        out = json.loads(res.stdout.strip().splitlines()[-1])
        sig = {k: (v["checksum"], v["first"]) for k, v in out.items()}
        if ref is None:
            ref = sig
        same = "same counters" if sig == ref else "COUNTERS DIFFER"
        print(f"{name:10s} " + "  ".join(f"{k} {v['fits_per_s']:.4e} ({v['ms']:.2f} ms)" for k, v in out.items()) + "  " + same, flush=True)


if __name__ == "__main__":
    main()
