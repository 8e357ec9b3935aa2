#!/usr/bin/env python3
"""Strong scaling of ONE library call over the GPUs of a box (McpOptions.n_devices): fixed total work, 1 / 2 / 4 / 8 devices,
end-to-end wall clock through the public API (host spec in, host result out, ProgressSink attached) next to the device time
and the host phases the library reports (lowering, enqueueing, waiting).  Prints one JSON object (profiles/r02_strong_scaling.json).

  headline : OLS n = 500, P = 9, 3.2e7 sims in one call (COUNTS)
  c2       : BASELINE config 2 — 50-point sweep, P = 11, Bonferroni, 1e6 sims per cell — COUNTS and BYTES | COUNTS (1.1 GB)
"""
import json
import statistics
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mcpower_b200 as M  # noqa: E402
from mcpower_b200 import spec as S  # noqa: E402


def measure(spec, grid, n_sims, nd, mode, reps=3):
    seed = M.splitmix64(2137)
    sink = lambda cur, tot: True  # noqa: E731
    ref = None
    M.run_batch(spec, grid, n_sims, seed, sink, mode=mode, device=0, n_devices=nd)  # arenas, pinned pool, module load
    rows = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = M.run_batch(spec, grid, n_sims, seed, sink, mode=mode, device=0, n_devices=nd)
        wall = time.perf_counter() - t0
        rows.append({"wall_ms": 1e3 * wall, "kernel_ms": r.kernel_ms, "lower_ms": r.lower_ms, "enqueue_ms": r.enqueue_ms,
                     "wait_ms": r.wait_ms, "d2h_bytes": r.d2h_bytes, "h2d_bytes": r.h2d_bytes, "waves": r.n_waves})
        ref = r.counters.copy()
    med = sorted(rows, key=lambda x: x["wall_ms"])[len(rows) // 2]
    fits = n_sims * len(grid)
    med.update(n_devices=nd, fits_per_s_e2e=fits / (med["wall_ms"] * 1e-3), fits_per_s_device=fits / (med["kernel_ms"] * 1e-3),
               host_share=1.0 - med["kernel_ms"] / med["wall_ms"])
    return med, ref


def main():
    have = M.device_count()
    counts = [n for n in (1, 2, 4, 8) if n <= have]
    out = {"devices_visible": have, "workloads": {}}
    grid50 = list(range(20, 1001, 20))
    for name, spec, grid, n, mode in (("headline_3.2e7_sims", S.headline_spec(), [500], 32_000_000, M.MODE_COUNTS),
                                      ("c2_1e6_sims_x50_counts", S.sweep_spec(), grid50, 1_000_000, M.MODE_COUNTS),
                                      ("c2_1e6_sims_x50_bytes", S.sweep_spec(), grid50, 1_000_000, M.MODE_COUNTS | M.MODE_BYTES)):
        base, rows = None, []
        for nd in counts:
            m, cnt = measure(spec, grid, n, nd, mode)
            if base is None:
                base = (m["wall_ms"], cnt)
            m["speedup_e2e"] = base[0] / m["wall_ms"]
            m["efficiency_e2e"] = m["speedup_e2e"] / nd
            m["bit_identical_to_one_device"] = bool((cnt == base[1]).all())
            rows.append(m)
            print(name, json.dumps(m), file=sys.stderr, flush=True)
        out["workloads"][name] = rows
    print(json.dumps(out))


if __name__ == "__main__":
    main()
