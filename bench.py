#!/usr/bin/env python3
"""bench.py — OLS fits/s at BASELINE.json's metric shape (n=500, k=8 → P=9), whole job over N GPUs.

A "step" is one pass of the hot path (engine_core::run_batch twin) over one batch of synthetic
simulations: SIMS_PER_STEP sims per GPU, each generated and fitted on-chip.  Sims shard across
ranks by global sim id with the same base seed — weak scaling, no data-path
collective; the only exchange is one all-reduce of the per-cell integer counters.

  value : fits/s from device time (CUDA events on the launch stream, max over ranks)
  e2e   : fits/s through the public API (mcpower_b200.run_batch → C ABI) with host buffers: spec JSON
          parse + lowering + H2D of the plan/tables + kernel + D2H of the counters, wall clock
  roofline      : instruction-issue view (the real bound): algorithmic warp-instructions/s ÷ issue peak
  roofline_fp64 : FP64-pipe view — algorithmic FP64 flop per fit ÷ measured DFMA peak
  secondary     : BASELINE configs[1] (sample-size sweep, correlated predictors + factor, Bonferroni) fits/s
  workloads     : BASELINE configs[2..4] (doomer scenario sweep, logistic GLM, random-intercept LMM) fits/s, device time
  cpu_baseline  : the oracle (CPU restatement of the reference algorithm) on all host threads, bounded sample

`--impl reference` times the oracle alone (the Rust engine cannot be built here: no cargo/rustc).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_ROWS = 500
SIMS_PER_STEP = 1 << 22          # per GPU per step
FLOP64_PER_FIT = 6.6e4           # 3.3e4 DFMA at n=500, P=9 (identity mix skipped: structural zeros)
USER_SEED = 2137                 # configs/config.json default seed; base_seed = splitmix64(seed) (find_power.rs:259)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin: float = 0.0, t_end: float = float("inf")) -> dict:
        """Summarise the samples taken inside [t_begin, t_end] (the timed region); the sampler itself is started before
        the warm-up because nvidia-smi needs a few hundred ms to produce its first line."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (ts, r) in self.rows if t_begin <= ts <= t_end + 0.15]
        window = "timed region"
        if not inside:  # timed region shorter than the sampling period: use the samples under load (warm-up + timed)
            inside, window = [r for (_, r) in self.rows], "warm-up + timed region"
        for r in inside:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_reference(spec, n_rows, seed, target_seconds, orc):
    """Time the oracle on all host threads over a bounded sample (~target_seconds of CPU work)."""
    threads = orc.host_threads()
    n = 2000 * threads
    orc.run_batch(spec, [n_rows], 256 * threads, seed, n_threads=threads)  # warm-up
    t0 = time.perf_counter()
    orc.run_batch(spec, [n_rows], n, seed, n_threads=threads)
    dt = time.perf_counter() - t0
    n_big = int(max(n, min(4_000_000, n * target_seconds / max(dt, 1e-6))))
    t0 = time.perf_counter()
    orc.run_batch(spec, [n_rows], n_big, seed, n_threads=threads)
    dt = time.perf_counter() - t0
    return n_big / dt, threads, n_big, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sims-per-step", type=int, default=SIMS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from mcpower_b200 import spec as S
    spec = S.headline_spec()
    config = {"workload": "ols_n500_k8_p9 (BASELINE configs[1] metric shape M: 8 independent N(0,1) predictors, "
                          "effects x1=x2=0.18, targets {x1,x2}, alpha=0.05, optimistic scenario)",
              "n_rows": N_ROWS, "n_predictors": 9, "sims_per_step_per_gpu": args.sims_per_step,
              "parallelism": f"sim-id sharding x{world}, counters all-reduced",
              "l2_policy": "n/a - no input tensors: every dataset is generated on-chip from (spec, seed, sim_id)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import orc
        seed = orc.splitmix64(USER_SEED)
        threads = orc.host_threads()
        per_step = 1500 * threads
        for _ in range(args.warmup):
            orc.run_batch(spec, [N_ROWS], per_step, seed, n_threads=threads)
        t0 = time.perf_counter()
        for i in range(args.steps):
            orc.run_batch(spec, [N_ROWS], per_step, seed, sim_begin=i * per_step, n_threads=threads)
        dt = time.perf_counter() - t0
        v = per_step * args.steps / dt
        line = {"impl": "reference", "metric": "ols_fits_per_sec", "value": v, "unit": "fits/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "fits/s", "cores": threads, "kind": "port",
                                 "sample": f"{per_step} sims x n={N_ROWS} per step, oracle/mcp_oracle.c (CPU restatement of "
                                           "the reference algorithm; the Rust engine cannot be built here: no cargo/rustc)"},
                "e2e": {"value": v, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import mcpower_b200 as M

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the CUDA backend has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    seed = M.splitmix64(USER_SEED)
    per = args.sims_per_step
    launches = 0
    E_STATS = 9 * 10 // 2 + 9 + 1  # doubles of sufficient statistics per fit at P = 9 (X'X packed | X'y | y'y)

    def step(i, mode=M.MODE_COUNTS, sink=None):
        # global sim ids: step-major, rank-minor — the union over ranks is one contiguous id range
        begin = (i * world + rank) * per
        return M.run_batch(spec, [N_ROWS], per, seed, sink, mode=mode, device=local_rank, sim_begin=begin)

    def reduce_counters(r):
        """The path's only collective: one all-reduce of the integer counters per (sample size, test)."""
        if dist is None:
            return r.counters
        c = torch.from_numpy(r.counters.astype(np.int64)).cuda()
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        return c.cpu().numpy()

    def timed(n_steps, first, **kw):
        """Wall clock over n_steps public-API calls (host spec in, host result out, counters all-reduced), barrier on both sides."""
        barrier()
        t = time.perf_counter()
        last = None
        for i in range(n_steps):
            last = None  # hand the previous result's page-locked buffers back to the pool before the next call asks for them
            last = step(first + i, **kw)
            reduce_counters(last)
        barrier()
        return time.perf_counter() - t, last

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step(i)
    # ---- timed region 1: device time (value)
    barrier()
    t_begin = time.time()
    t0 = time.perf_counter()
    dev_ms, total_counters = 0.0, None
    h2d = d2h = 0
    for i in range(args.steps):
        r = step(args.warmup + i)
        dev_ms += r.kernel_ms
        launches += r.gpu_launches
        h2d, d2h = r.h2d_bytes, r.d2h_bytes
        total_counters = r.counters.copy() if total_counters is None else total_counters + r.counters
    barrier()
    wall = time.perf_counter() - t0
    # ---- timed region 2: end-to-end through the public API (host buffers in, host counters out, the all-reduce inside)
    base = args.warmup + args.steps
    e2e_wall, r_e2e = timed(args.steps, base)
    # ---- the reference's real host path: a ProgressSink is always attached (engine-py/src/lib.rs:57-71) ...
    n_reports = [0]

    def sink(cur, tot):
        n_reports[0] += 1
        return True

    step(base + args.steps, sink=sink)
    n_reports[0] = 0
    prog_wall, r_prog = timed(args.steps, base + args.steps + 1, sink=sink)
    reports_per_call = n_reports[0] / max(args.steps, 1)
    # ---- ... and the orchestrator folds the sim-major byte tensors (BYTES | COUNTS, sink attached)
    step(base + 2 * args.steps + 1, mode=M.MODE_COUNTS | M.MODE_BYTES, sink=sink)
    bytes_wall, r_bytes = timed(args.steps, base + 2 * args.steps + 2, mode=M.MODE_COUNTS | M.MODE_BYTES, sink=sink)
    clocks = sampler.stop(t_begin, time.time())  # every timed region

    t = torch.tensor([dev_ms, wall, e2e_wall, prog_wall, bytes_wall], dtype=torch.float64, device="cuda")
    cnt = torch.from_numpy(total_counters.astype(np.int64)).cuda()
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, wall, e2e_wall, prog_wall, bytes_wall = (float(v) for v in t.tolist())
    counters = cnt.cpu().numpy()
    total_fits = per * args.steps * world
    value = total_fits / (dev_ms * 1e-3)
    e2e_value = total_fits / e2e_wall

    if rank == 0:
        power = counters[0, :2] / total_fits
        dfma = M.pipe_bench(0, local_rank)
        peak_tf = 2.0 * dfma / 1e12
        ach_tf = (value / world) * FLOP64_PER_FIT / 1e12
        # The path is bound by instruction issue / datapath cycles, not by HBM (generated data never leaves the SM) and
        # not by tensor cores: the headline roofline is the issue roofline,
        # ALGORITHMIC warp-instructions per fit (2.6e5 thread-instructions) x fits/s against one issue per
        # scheduler per clock at the SM clock sampled during the timed region.  frac_essential uses the instruction count
        # derived in DESIGN.md section 3.3 (Philox + packed Horner + DFMA + widening + tails: what no implementation of
        # this arithmetic can avoid) instead of SURVEY's estimate, which is close to what the kernel executes.
        prof = json.loads((ROOT / "profiles" / "inst_per_fit.json").read_text())
        sm_clock_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        peak_issue = 148 * 4 * sm_clock_hz
        ach_issue = (value / world) * prof["algorithmic_thread_inst_per_fit"] / 32.0
        ess_issue = (value / world) * prof["essential_thread_inst_per_fit"] / 32.0
        line = {"metric": "ols_fits_per_sec", "value": value, "unit": "fits/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "fits/s", "h2d_bytes_per_step": int(r_e2e.h2d_bytes),
                        "d2h_bytes_per_step": int(r_e2e.d2h_bytes), "ms_per_step": 1e3 * e2e_wall / args.steps,
                        "note": "public API, host spec in / host counters out, the counter all-reduce inside the timed region; "
                                "byte counts are what the library copied (plan + tables up, counters down)"},
                "e2e_progress": {"value": total_fits / prog_wall, "unit": "fits/s", "ms_per_step": 1e3 * prog_wall / args.steps,
                                 "reports_per_call": reports_per_call, "gpu_launches_per_call": int(r_prog.gpu_launches),
                                 "waves_per_call": int(r_prog.n_waves),
                                 "note": "the same with a ProgressSink attached (the reference's hosts always pass one): progress "
                                         "comes from a device counter published to mapped host memory, the batch is not split"},
                "e2e_bytes": {"value": total_fits / bytes_wall, "unit": "fits/s", "ms_per_step": 1e3 * bytes_wall / args.steps,
                              "h2d_bytes_per_step": int(r_bytes.h2d_bytes), "d2h_bytes_per_step": int(r_bytes.d2h_bytes),
                              "waves_per_call": int(r_bytes.n_waves), "kernel_ms_per_call": r_bytes.kernel_ms,
                              "host_ms_per_call": {"lower": r_bytes.lower_ms, "enqueue": r_bytes.enqueue_ms, "wait": r_bytes.wait_ms},
                              "note": "BYTES | COUNTS with the sink attached: the sim-major BatchResult tensors land in page-locked "
                                      "host memory, each wave's copy overlapping the next wave's kernels"},
                "host_ms_per_call": {"lower": r_e2e.lower_ms, "enqueue": r_e2e.enqueue_ms, "wait": r_e2e.wait_ms},
                "gpu_launches": launches,
                "roofline": {"bound": "issue", "achieved": ach_issue / 1e9, "peak": peak_issue / 1e9, "unit": "Gwarp-inst/s",
                             "frac": ach_issue / peak_issue, "frac_essential": ess_issue / peak_issue,
                             "frac_executed": (value / world) * prof["warp_inst_per_fit"] / peak_issue,
                             "traffic": float(E_STATS * 8 * per),
                             "note": "per GPU, dominant kernel ols_lean_kernel (98 % of the step); achieved = algorithmic "
                                     "2.6e5 thread-instructions per fit (SURVEY 8d) / 32 x fits/s; frac_essential = the irreducible "
                                     "instruction count of DESIGN.md 3.3 (" + str(int(prof["essential_thread_inst_per_fit"])) +
                                     " thread-instructions per fit); frac_executed = what ncu counted at HEAD (" + prof["source"] +
                                     "); peak = 148 SMs x 4 schedulers x SM clock; traffic = the statistics hand-off the kernel "
                                     "writes, E x 8 B x fits per launch (E = 55 entries at P = 9), computed from the plan",
                             "executed_warp_inst_per_fit": prof["warp_inst_per_fit"]},
                "roofline_fp64": {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                                  "frac": ach_tf / peak_tf,
                                  "note": "algorithmic 6.6e4 FP64 flop per fit against the DFMA rate measured on this device "
                                          "(mcp_pipe_bench kind 0); MEASURED_PEAKS.json has no FP64 entry"},
                "power_x1_x2": [float(p) for p in power],
                "wall_ms_per_step": 1e3 * wall / args.steps}
        if world == 1 and M.device_count() > 1 and not args.no_secondary:
            # the same batch sharded over every visible GPU INSIDE one library call (McpOptions.n_devices): strong scaling of
            # one run_batch, counters summed by the library
            nd = M.device_count()
            M.run_batch(spec, [N_ROWS], per * nd, seed, device=0, n_devices=nd)
            tt = time.perf_counter()
            rm = M.run_batch(spec, [N_ROWS], per * nd, seed, device=0, n_devices=nd)
            tw = time.perf_counter() - tt
            line["in_library_multi_gpu"] = {"n_devices": nd, "sims": per * nd, "value": per * nd / (rm.kernel_ms * 1e-3),
                                            "e2e": per * nd / tw, "lower_ms": rm.lower_ms, "enqueue_ms": rm.enqueue_ms,
                                            "wait_ms": rm.wait_ms}
        if not args.no_secondary:
            # second headline: BASELINE configs[1] — sample-size sweep 20..1000 step 20, 8 correlated
            # predictors + 3-level factor, Bonferroni, at the BASELINE size of 1e6 sims per cell; fits/s = n_sims x n_ss / time
            sweep = S.sweep_spec()
            grid = list(range(20, 1001, 20))
            n2 = 1_000_000
            M.run_batch(sweep, grid, 4096, seed, device=local_rank)
            rs = M.run_batch(sweep, grid, n2, seed, device=local_rank)
            tt = time.perf_counter()
            rs = M.run_batch(sweep, grid, n2, seed, sink, device=local_rank)
            c2_wall = time.perf_counter() - tt
            rb = M.run_batch(sweep, grid, n2, seed, sink, mode=M.MODE_COUNTS | M.MODE_BYTES, device=local_rank)  # pins the result pool
            rb = None
            tt = time.perf_counter()
            rb = M.run_batch(sweep, grid, n2, seed, sink, mode=M.MODE_COUNTS | M.MODE_BYTES, device=local_rank)
            c2_bytes_wall = time.perf_counter() - tt
            line["secondary"] = {"workload": "c2_sweep_p11_grid50_bonferroni", "n_sims": n2, "n_sample_sizes": len(grid),
                                 "fits_per_s": n2 * len(grid) / (rs.kernel_ms * 1e-3),
                                 "sims_per_s": n2 / (rs.kernel_ms * 1e-3), "n_gpus": 1,
                                 "e2e_progress_fits_per_s": n2 * len(grid) / c2_wall, "e2e_progress_ms": 1e3 * c2_wall,
                                 "e2e_bytes_fits_per_s": n2 * len(grid) / c2_bytes_wall, "e2e_bytes_ms": 1e3 * c2_bytes_wall,
                                 "bytes_d2h": int(rb.d2h_bytes), "bytes_waves": int(rb.n_waves), "kernel_ms": rs.kernel_ms,
                                 "bytes_kernel_ms": rb.kernel_ms,
                                 "lower_ms": rs.lower_ms}
            # the other BASELINE configs (device time, one GPU): robustness scenario sweep, logistic GLM power curve,
            # random-intercept mixed model — the split path's kernels
            doomer = dict(sweep, scenario=S.scenario("doomer"), crit_values={"alpha": 5e-7, "posthoc_alpha": None})
            others = {}
            for name, wspec, wgrid, wn in (("c3_doomer_alpha5e-7_grid50", doomer, grid, 1 << 18),
                                           ("c4_logit_n50_2000", S.logit_spec(), [50, 100, 200, 500, 1000, 2000], 1 << 17),
                                           ("c5_lmm_30clusters_grid", S.lmm_spec(30, 0.2), [60, 120, 300, 600, 900], 1 << 19)):
                M.run_batch(wspec, wgrid, 4096, seed, device=local_rank)
                rw = M.run_batch(wspec, wgrid, wn, seed, device=local_rank)
                others[name] = {"n_sims": wn, "n_sample_sizes": len(wgrid), "fits_per_s": wn * len(wgrid) / (rw.kernel_ms * 1e-3),
                                "gpu_launches": int(rw.gpu_launches)}
            line["workloads"] = others
        if world == 1 and not args.no_cpu_baseline:
            from oracle import orc
            v, cores, n_big, dt = cpu_reference(spec, N_ROWS, seed, 12.0, orc)
            line["cpu_baseline"] = {"value": v, "unit": "fits/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_big} sims x n={N_ROWS} in {dt:.1f} s, oracle/mcp_oracle.c on all host threads"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
