/*
 * mcpower_b200.h — C ABI of the B200-native MCPower hot-path backend (libmcpower_b200.so).
 *
 * Drop-in seam: engine_core::run_batch
 *   pub fn run_batch(spec: &SimulationSpec, sample_sizes: &[u32], n_sims: u32, base_seed: u64,
 *                    sink: Option<&dyn ProgressSink>) -> Result<BatchResult, EngineError>
 *   (reference: crates/engine-core/src/batch.rs:130-146; callers find_power.rs:337,
 *    find_sample_size.rs:213, single_core.rs, introspect.rs:7-8, bin/throughput.rs:668).
 * A Rust `engine-cuda` crate binds these symbols with `extern "C"` (INTEGRATION.md shows the stub);
 * the Python mirror in mcpower_b200/engine.py binds them with ctypes.
 *
 * Plain pointers and sizes only; no C++/torch types; nothing throws across the boundary.
 * All functions return an McpStatus code; mcp_last_error() gives the message for the calling thread.
 */
#ifndef MCPOWER_B200_H
#define MCPOWER_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors EngineError (crates/engine-core/src/spec.rs:19-29) + backend-specific codes. */
enum {
    MCP_STATUS_OK = 0,
    MCP_STATUS_INVALID_SPEC = 1,        /* EngineError::InvalidSpec(String)   */
    MCP_STATUS_CORRELATION_NOT_PSD = 2, /* EngineError::CorrelationNotPSD     */
    MCP_STATUS_RANK_DEFICIENT = 3,      /* EngineError::RankDeficient(u32)    */
    MCP_STATUS_CANCELLED = 4,           /* EngineError::Cancelled             */
    MCP_STATUS_CUDA_ERROR = 5,          /* no device / launch failure: never a silent CPU fallback */
    MCP_STATUS_UNSUPPORTED = 6          /* spec feature outside the CUDA path's scope */
};

/* Output selection (bit-or). COUNTS = per-cell counters (what aggregate_batch derives,
 * crates/engine-orchestrator/src/aggregation.rs:61-271); BYTES = the sim-major 0/1 tensors of
 * BatchResult (spec.rs:552-600); CAPTURE = per-fit squared statistics (run_batch_st_capture,
 * batch.rs:175-197, extended to every sample-size index). */
enum { MCP_MODE_COUNTS = 1, MCP_MODE_BYTES = 2, MCP_MODE_CAPTURE = 4 };

/* ProgressSink (spec.rs:38-52): report(current, total) ~50x per run, called on the thread that made the library call, while
 * the kernels run (the batch is NOT cut into small launches for it: the kernels publish their progress through a mapped
 * host word the waiting thread polls); return 0 to request cancellation.  is_cancelled: the trait's cheap poll, asked every
 * ~100 us by the same waiting thread; return nonzero to cancel.  On cancellation CTAs that have not started exit at once,
 * the partial result is discarded and the call returns MCP_STATUS_CANCELLED (batch.rs:128-129).  Neither callback may call
 * back into the library for a device the call is using (the per-device lock is held). */
typedef int (*McpProgressFn)(uint64_t current, uint64_t total, void* user);
typedef int (*McpCancelFn)(void* user);

typedef struct McpOptions {
    int32_t device;      /* first CUDA ordinal; -1 = the calling thread's current device */
    uint32_t mode;       /* MCP_MODE_*; 0 = COUNTS */
    uint64_t sim_begin;  /* first GLOBAL sim id of this call (same base_seed, disjoint id ranges = one bigger batch) */
    uint32_t wave_sims;  /* sims per kernel launch; 0 = auto (as many as the scratch arena allows) */
    int32_t n_devices;   /* 0 or 1: one device; N > 1: the call shards its sims over ordinals device .. device+N-1 (contiguous
                            global sim-id ranges, same base_seed: every output bit equals the one-device result; counters are
                            summed, byte tensors land in sim order); -1: every visible device from `device` on */
    McpCancelFn is_cancelled; /* may be NULL */
} McpOptions;

/* Counter row layout per sample-size index (width = 4T + 12 + 2F + 4C, C = n_posthoc, J = T + C; aggregation.rs:91-242):
 *   [0,T) uncorrected successes   [T,2T) corrected successes   2T converged   2T+1 overall
 *   [2T+2, +J+1) histogram of #uncorrected-significant tests per sim (targets AND posthoc contrasts)
 *   [2T+3+J, +J+1) histogram of #corrected-significant tests per sim
 *   then F: sims with factor f sparse-excluded (factor_excluded code 1)
 *   then joint Wald successes (MLE arm), fits with boundary_hit != 0 (MLE arm)
 *   then C posthoc uncorrected successes, C posthoc corrected successes
 *   then converged fits with boundary_hit == 1 (singular fits), then F: sims where the GLM separation fallback
 *   dropped factor f (factor_excluded code 2), then 5: converged fits with variance component k pinned at 0
 *   (k = intercept, slopes, extra groupings: EstimatorExtras::Mle::boundary_component_counts, result.rs:362-398). */
typedef struct McpBatchResult {
    uint32_t n_sims, n_sample_sizes, n_targets, n_factors;
    uint8_t* uncorrected;     /* [n_sims][n_ss][T]  or NULL */
    uint8_t* corrected;       /* [n_sims][n_ss][T]  or NULL */
    uint8_t* converged;       /* [n_sims][n_ss]     or NULL */
    uint8_t* overall;         /* [n_sims][n_ss]     or NULL (spec.report_overall) */
    uint8_t* factor_excluded; /* [n_sims][n_ss][F]  or NULL; 0 included, 1 sparse-excluded, 2 separation fallback */
    uint8_t* joint;           /* [n_sims][n_ss]     or NULL (MLE arm: joint_unc == joint_cor, spec.rs:574-582) */
    uint8_t* boundary_hit;    /* [n_sims][n_ss]     or NULL (MLE arm).  bits 0-1: BatchResult::boundary_hit (0 none, 1 a variance
                               * component at 0, 2 optimiser at its bound); bits 2-6: BatchResult::pinned_components bits 0-4
                               * (intercept, slopes, extra groupings of the fitted model; spec.rs:562-574) */
    double* t_sq;             /* [n_sims][n_ss][T]  or NULL (CAPTURE) */
    uint32_t counter_width;
    uint32_t gpu_launches;    /* kernels launched for this call */
    uint64_t* counters;       /* [n_ss][counter_width] */
    double kernel_ms;         /* device time of the power kernel(s), CUDA events on the launch stream */
    uint32_t n_posthoc;       /* C = total pairwise posthoc contrasts over all blocks (spec.posthoc, batch.rs:2183-2227) */
    uint32_t reserved;
    uint8_t* posthoc_unc;     /* [n_sims][n_ss][C] or NULL (BatchResult::posthoc_unc, spec.rs:557-559) */
    uint8_t* posthoc_cor;     /* [n_sims][n_ss][C] or NULL */
    uint64_t h2d_bytes;       /* bytes this call copied host -> device (plan + tables, every device) */
    uint64_t d2h_bytes;       /* bytes this call copied device -> host (counters, byte tensors, capture) */
    uint32_t n_devices;       /* devices that ran a shard of the call */
    uint32_t n_waves;         /* kernel waves over all devices (1 per device unless the scratch arena or BYTES pipelining split it) */
    double lower_ms;          /* host: validation + lowering (critical-value tables, exact factor walks) */
    double enqueue_ms;        /* host: result allocation + enqueueing every device's uploads, launches and copies */
    double wait_ms;           /* host: from the last enqueue until every device finished (kernels + copies) */
} McpBatchResult;

/* run_batch twin. spec_json = serde JSON of SimulationSpec (same field names / enum spellings). */
int mcp_run_batch(const char* spec_json, size_t spec_len, const uint32_t* sample_sizes, uint32_t n_sample_sizes,
                  uint32_t n_sims, uint64_t base_seed, const McpOptions* options, McpProgressFn progress, void* user,
                  McpBatchResult** out);
void mcp_free_result(McpBatchResult* result);

/* ---- Orchestrator entry points for ONE lowered scenario -------------------------
 * find_power (engine-orchestrator/src/find_power.rs:337 + aggregation.rs:61-271): run_batch + the per-N PowerResult
 * fold.  *json_out = JSON array of PowerResult (serde field names: n, n_sims, target_indices, contrast_pairs,
 * power_uncorrected, power_corrected, ci_uncorrected, ci_corrected, convergence_rate, boundary_hit, estimator_extras,
 * overall_significant_rate, success_counts_*, convergence_count, overall_significant_count, overall_significant_ci,
 * success_count_histogram_*, grid_warnings, posthoc, factor_exclusion_counts, factor_separation_counts); free with
 * mcp_free_string.  boundary_hit (per sim) is filled only when options->mode includes MCP_MODE_BYTES. */
int mcp_find_power(const char* spec_json, size_t spec_len, const uint32_t* sample_sizes, uint32_t n_sample_sizes,
                   uint32_t n_sims, uint64_t base_seed, const McpOptions* options, McpProgressFn progress, void* user,
                   char** json_out);

/* find_sample_size (find_sample_size.rs:83-257, grid.rs:37-330, fit.rs:45-246): cluster-aware grid over
 * [from_size, to_size] (by = fixed step, or point count when by_auto != 0; log_grid != 0 = log-spaced), one run_batch
 * over the grid, per-N fold, first-N lookups and model-based crossing fits.  *json_out = SampleSizeResult object
 * (grid_or_trace, first_achieved, first_joint_achieved, fitted, fitted_joint, first_overall_achieved, fitted_overall,
 * cluster_atom, target_power, grid_warnings). */
int mcp_find_sample_size(const char* spec_json, size_t spec_len, uint64_t from_size, uint64_t to_size, uint64_t by,
                         int32_t by_auto, int32_t log_grid, double target_power, uint32_t n_sims, uint64_t base_seed,
                         const McpOptions* options, McpProgressFn progress, void* user, char** json_out);
void mcp_free_string(char* s);

/* ---- Contract entry points -------------------------------------------------------------
 * The hosts' own payload: the `Vec<SimulationContract>` blob engine-py / engine-r forward verbatim
 * (engine-py/src/orchestrator_bridge.rs:23-30 — msgpack, `rmp_serde::to_vec_named`; format 0) or the same tree as serde
 * JSON (format 1).  Each contract is validated (engine-contract/src/validate.rs:22-50) and lowered like
 * `contract_to_simulation_spec` (engine-core/src/contract_adapter.rs:74-230) + `lower_contracts`
 * (engine-orchestrator/src/find_power.rs:251-288: factor_min_level_count forced to the configured 5, every scenario run
 * with the call seed splitmix64(user_seed), results labelled with the design_test term positions).
 * mcp_lower_contracts: *json_out = [{"label", "spec" (SimulationSpec JSON, ready for mcp_run_batch),
 * "report_target_indices", "report_contrast_pairs"}, ...].  The two runners return ScenarioResult<T>
 * (result.rs:581-584): {"scenarios": [[label, PowerResult | SampleSizeResult], ...]}.  Errors: MCP_INVALID_SPEC with
 * "invalid scenarios input: contract <i> (<label>): ..." (OrchestratorError::InvalidScenarios). */
int mcp_lower_contracts(const uint8_t* contracts, size_t len, int32_t format, char** json_out);
int mcp_find_power_contracts(const uint8_t* contracts, size_t len, int32_t format, uint64_t sample_size, uint32_t n_sims,
                             uint64_t user_seed, const McpOptions* options, McpProgressFn progress, void* user, char** json_out);
int mcp_find_sample_size_contracts(const uint8_t* contracts, size_t len, int32_t format, uint64_t from_size, uint64_t to_size,
                                   uint64_t by, int32_t by_auto, int32_t log_grid, double target_power, uint32_t n_sims,
                                   uint64_t user_seed, const McpOptions* options, McpProgressFn progress, void* user,
                                   char** json_out);

/* CrossingFit (result.rs:416-450); status: 0 fitted, 1 at_or_below_min, 2 not_reached, 3 non_monotone. */
typedef struct McpCrossingFit {
    int32_t status;
    int32_t has_ci_lo, has_ci_hi, has_n_approx;
    double n_star, ci_lo, ci_hi, max_violation;
    uint64_t n_achievable, n_min, n_approx;
} McpCrossingFit;
/* fit_crossing (fit.rs:45-120): monotonicity gate, PAVA, interpolated crossing, Wilson band inversion, probit hint. */
int mcp_fit_crossing(const uint64_t* grid, const uint64_t* counts, uint32_t n_points, uint64_t n_sims, double target_power,
                     uint64_t atom, McpCrossingFit* out);
/* build_grid (grid.rs:37-147). */
int mcp_build_grid(uint64_t from_size, uint64_t to_size, uint64_t by, int32_t by_auto, int32_t log_grid, uint64_t atom,
                   uint64_t hard_min, uint64_t* grid_out, uint32_t capacity, uint32_t* n_out, uint32_t* n_warnings);
/* wilson_ci (aggregation.rs:18-33). */
void mcp_wilson_ci(uint64_t k, uint64_t n, double* lo, double* hi);

/* introspect::capture_sim0_data twin (introspect.rs:104-155) for ANY sim id: the f32 design matrix
 * (column-major, max_n × P) and outcome the power kernel generates for (base_seed, sim_id). */
int mcp_generate_sim(const char* spec_json, size_t spec_len, uint32_t max_n, uint64_t base_seed, uint64_t sim_id,
                     int32_t device, float* x_colmajor, float* y);

/* introspect::fit_provided_data twin (introspect.rs:174-293), OLS: column-major f64 design (narrowed to
 * f32 first, like the reference), n × p.  out = beta[p], se[p], t_sq[p], sigma_sq, sst, rss, converged. */
int mcp_fit_data(const double* x_colmajor, const double* y, uint32_t n, uint32_t p, int32_t device, double* out);

/* introspect::fit_provided_data twin WITH the spec (introspect.rs:174-562): the supplied dataset goes through the shipped fit
 * code of spec.estimator — OLS (sufficient statistics + Cholesky), unclustered GLM (the IRLS of the power kernel, truth start
 * beta0 = effect_sizes), Gaussian random-intercept MLE (cluster sufficient statistics + profiled REML) — and comes back as
 * IntrospectResults: betas per design column, se and the natural (unsquared) statistic per target (spec.target_indices), the
 * squared uncorrected threshold and df_resid = N - n_predictors.  design = column-major f64 nrow x ncol (narrowed to the f32
 * data plane first, like the reference); cluster_ids may be NULL (= the spec's sizing) and must otherwise equal it. */
typedef struct McpFitInfo {
    int32_t converged;
    int32_t boundary_hit;        /* MLE: 0 none, 1 tau-hat = 0, 2 upper bound; else 0 */
    int32_t iterations;          /* GLM: IRLS passes; else 0 */
    int32_t n_targets;
    double crit_sq_uncorrected;  /* t_crit_sq_uncorrected[0] (t for OLS, z for GLM / MLE) */
    double df_resid;
    double sigma_sq_hat;         /* OLS: residual variance; MLE: sigma²-hat at theta-hat; GLM: NaN */
    double tau_sq_hat;           /* MLE: the variance component theta² sigma²; else NaN */
    double deviance;             /* GLM; else NaN */
    double deviance_null;        /* GLM; else NaN */
    double slope_var_hat[2];     /* MLE with random slopes: variances of the slope effects (Sigma-hat = sigma² Lambda Lambda'); else NaN */
    double slope_corr_hat[3];    /* ... and the correlations (slope0, intercept), (slope1, intercept), (slope1, slope0); else NaN */
    double extra_var_hat[2];     /* MLE with extra groupings: variance of each grouping's random intercept; else NaN */
    int32_t pinned_components;   /* MLE: bit k set iff variance component k (intercept, slopes, extra groupings) sat at 0 */
} McpFitInfo;
int mcp_fit_provided_data(const char* spec_json, size_t spec_len, const double* design_colmajor, uint32_t nrow, uint32_t ncol,
                          const double* outcome, const uint32_t* cluster_ids, int32_t device, double* betas /* [ncol] */,
                          double* se /* [n_targets] */, double* statistic /* [n_targets] */, McpFitInfo* info);

/* CritValueTable::build_with_tukey_k twin (critvals.rs:827-1012; host f64).
 * t_unc_sq[n_ss], corr_sq[n_ss*T], posthoc_sq[n_ss], overall[n_ss]; any pointer may be NULL. */
int mcp_crit_table(const char* spec_json, size_t spec_len, const uint32_t* sample_sizes, uint32_t n_sample_sizes,
                   uint32_t n_predictors, double* t_unc_sq, double* corr_sq, double* posthoc_sq, double* overall);

/* rng::splitmix64_finalize (rng.rs:30): hosts derive base_seed = splitmix64(user_seed) (find_power.rs:259). */
uint64_t mcp_splitmix64(uint64_t z);

/* Measured pipe throughput (thread-operations per second on `device`) for the roofline denominators:
 * kind 0 = DFMA, 1 = FMUL+FADD pair (counts 2 ops), 2 = Philox round mix (2 IMAD.WIDE + 2 LOP3 → counts 4 ops),
 * 3 = F2F.F64.F32 + DADD (counts 2 ops), 4 = DADD. */
int mcp_pipe_bench(int32_t kind, int32_t device, double* thread_ops_per_s);

/* Test hook (not part of the reference seam): pin the OLS power kernel — 1 one thread per study, 4 lane quad solving by
 * call, 5 lean generate+accumulate / solve pair, 6 general split pair; 0 = the library's own choice.  Every variant computes
 * the same bits; the parity tests run each OLS case under all of them.  Process-wide; returns the previous setting. */
int mcp_test_force_ols_variant(int variant);

int mcp_device_count(void);
const char* mcp_last_error(void);
const char* mcp_version(void);

#ifdef __cplusplus
}
#endif
#endif
