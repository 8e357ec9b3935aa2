#!/usr/bin/env python3
"""Secondary workloads (BASELINE.json configs 1-5) on one GPU: device-time fits/s and the oracle's fits/s on a
bounded sample.  Prints one JSON object; used for profiles/*_workloads.json (not the headline bench)."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mcpower_b200 as M  # noqa: E402
from mcpower_b200 import spec as S  # noqa: E402
from oracle import orc  # noqa: E402


def workloads():
    c1 = S.add_interaction(S.base_spec(2), [1, 2], 0.2)
    c1["effect_sizes"][1], c1["effect_sizes"][2] = 0.3, 0.2
    c1["target_indices"] = [1, 2, 3]
    sweep = S.sweep_spec()
    bh = dict(sweep, correction_method="benjamini_hochberg")
    doomer = dict(sweep, scenario=S.scenario("doomer"), crit_values={"alpha": 5e-7, "posthoc_alpha": None})
    grid50 = list(range(20, 1001, 20))
    return [
        ("c1_ols_interaction_n100", c1, [100], 1600, 1600),
        ("headline_ols_n500_p9", S.headline_spec(), [500], 1 << 21, 20000),
        ("c2_sweep_p11_grid50_bonferroni", sweep, grid50, 1 << 19, 2000),
        ("c2_sweep_p11_grid50_bh", bh, grid50, 1 << 19, 2000),
        ("c3_doomer_alpha5e-7_grid50", doomer, grid50, 1 << 19, 1000),
        ("c4_logit_n50_2000", S.logit_spec(), [50, 100, 200, 500, 1000, 2000], 1 << 19, 1000),
        ("c5_lmm_30clusters_grid", S.lmm_spec(30, 0.2), [60, 120, 300, 600, 900], 1 << 20, 2000),
    ]


def main():
    seed = M.splitmix64(2137)
    out = {}
    threads = orc.host_threads()
    for name, spec, grid, n_gpu, n_cpu in workloads():
        M.run_batch(spec, grid, min(n_gpu, 4096), seed)  # warm-up (module load, local-memory sizing)
        r = M.run_batch(spec, grid, n_gpu, seed)
        t0 = time.perf_counter()
        r2 = M.run_batch(spec, grid, n_gpu, seed, sim_begin=n_gpu)
        wall = time.perf_counter() - t0
        fits = n_gpu * len(grid)
        t0 = time.perf_counter()
        orc.run_batch(spec, grid, n_cpu, seed, n_threads=threads)
        cpu = time.perf_counter() - t0
        out[name] = {"n_sims": n_gpu, "n_sample_sizes": len(grid), "gpu_fits_per_s": fits / (r2.kernel_ms * 1e-3),
                     "gpu_sims_per_s": n_gpu / (r2.kernel_ms * 1e-3), "gpu_e2e_fits_per_s": fits / wall,
                     "kernel_ms": r2.kernel_ms, "cpu_oracle_fits_per_s": n_cpu * len(grid) / cpu, "cpu_threads": threads,
                     "power_first_target": [float(v) for v in r.power()[[0, -1], 0]],
                     "converged_frac": float(r.n_converged.sum() / fits)}
        print(name, json.dumps(out[name]), file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
