"""mcpower_b200 — B200-native backend for MCPower's per-simulation hot loop (engine_core::run_batch).

`run_batch(spec, sample_sizes, n_sims, base_seed, sink)` mirrors the reference entry point
(crates/engine-core/src/batch.rs:130-146) and executes on the GPU through the C ABI of
libmcpower_b200.so (include/mcpower_b200.h).  There is no CPU fallback.
"""
from . import spec  # noqa: F401
from .engine import (MODE_BYTES, MODE_CAPTURE, MODE_COUNTS, BatchResult, Cancelled, CorrelationNotPSD,  # noqa: F401
                     CudaError, EngineError, InvalidSpec, RankDeficient, Unsupported, counters_from_bytes, crit_table,
                     build_grid, device_count, find_power, find_power_contracts, find_sample_size,
                     find_sample_size_contracts, fit_crossing, fit_data, fit_provided_data, force_ols_variant, generate_sim, lower_contracts,
                     pipe_bench, run_batch, splitmix64, wilson_ci)
