// plan.h — the device-side lowering of one `SimulationSpec` + sample-size grid.
//
// Host code (lower.hpp) fills a DevicePlan once per mcp_run_batch call; kernels read it
// from global/constant memory.  Everything that does not depend on the per-sim RNG
// (critical values, exclusion-variant tables, exact factor walks, heteroskedasticity
// moments) is precomputed on the host, mirroring the reference's "host builds the
// CritValueTable, the loop only compares" split (crates/engine-core/src/critvals.rs:1-9).
#pragma once
#include <stdint.h>

#define MCP_MAX_K 24          // continuous (non-factor) predictors
#define MCP_MAX_EXTRA_LEVELS 96  // levels of the extra groupings together (the dense part of the mixed-model fit)
#define MCP_MAX_SLOPES 2      // random slopes on the primary grouping (q = 1 + slopes <= 3 random effects per cluster)
#define MCP_COOP_MAX_PROB 0.65  // marginal-swap probability below which the scenario kernels take the warp-cooperative marginal pass
#define MCP_MAX_P 32          // total design columns incl. intercept
#define MCP_MAX_T 32          // marginal targets + contrasts
#define MCP_MAX_PH 32         // posthoc pairwise contrasts (all blocks)
#define MCP_MAX_FACTORS 6
#define MCP_MAX_LEVELS 48     // Σ factor_n_levels
#define MCP_MAX_INTER 16
#define MCP_MAX_INTER_COMP 48
#define MCP_MAX_MENU 8
#define MCP_MAX_TEMPLATE_P 16 // kernels with register-resident X'X exist for P <= this
// Wider designs (17 .. 32 columns: many factor levels) run on the split path's kernels instantiated at two padded widths: the
// plan carries the real P, the padding columns are generated as zeros with zero effects, and every fit variant keeps only the
// real columns — i.e. the fit goes through the runtime-p solvers the sparse-factor exclusion already uses (reduced fits on kept
// columns, batch.rs:970-1013).  Every real entry of X'X / X'y is the same fma chain as in an exact-width kernel.
#define MCP_WIDE_P_A 24
#define MCP_WIDE_P_B 32
static inline constexpr int mcp_template_p(int P) { return P <= MCP_MAX_TEMPLATE_P ? P : (P <= MCP_WIDE_P_A ? MCP_WIDE_P_A : MCP_WIDE_P_B); }

// Staging geometry of the lane-quad kernels (kernels_lean.cuh, kernels_q4.cuh).  Per study, four finished rows are staged in shared
// memory as Z = [z_0 .. z_{M-1} | z_N, z_{N-1}, .. z_{N-M+1}] (M = ceil((P+1)/2) slots per half; the high half sits `hoff` doubles
// behind the low one, a row takes `row` doubles, a study `study`).  The three strides decide the bank behaviour of the kernel's
// shared-memory traffic — 16-byte stores of a finished row by the 4 lanes x 2 studies of a quarter-warp, 16-byte and (odd M)
// 8-byte loads of the two halves by quarter- / half-warps — and are chosen per M so that every one of them is conflict-free
// (tools/stage_geometry.py enumerates the candidates; ncu showed 2-way conflicts on the 8-byte loads at M = 5, the headline
// shape, and 4-way conflicts on the stores at M = 7, 8 with the former 2·MH / 8·MH + 2 rule).
struct McpStageGeom {
    int hoff, row, study;
};
constexpr McpStageGeom mcp_stage_geom(int M) {
    return M <= 1 ? McpStageGeom{2, 4, 18}
         : M == 2 ? McpStageGeom{2, 4, 22}
         : M == 4 ? McpStageGeom{4, 10, 40}
         : M == 6 ? McpStageGeom{6, 12, 50}
         : M == 9 ? McpStageGeom{10, 20, 82}
         : M >= 10 ? McpStageGeom{(M + 1) & ~1, 2 * ((M + 1) & ~1), 8 * ((M + 1) & ~1) + 2}  // padded wide widths: not tuned
                  : McpStageGeom{8, 20, 82};  // M = 3, 5, 7, 8
}

enum McpDist { D_NORMAL = 0, D_BINARY = 1, D_RIGHT_SKEWED = 2, D_LEFT_SKEWED = 3, D_HIGH_KURTOSIS = 4, D_UNIFORM = 5,
               D_UPLOADED_FACTOR = 6, D_UPLOADED_BINARY = 7, D_UPLOADED_DATA = 8 };  // engine-contract/src/distribution.rs:23-25
enum McpRes { R_NORMAL = 0, R_RIGHT_SKEWED = 1, R_LEFT_SKEWED = 2, R_HIGH_KURTOSIS = 3, R_UNIFORM = 4 };
enum McpOutcome { O_CONTINUOUS = 0, O_BINARY = 1, O_COUNT = 2 };
enum McpEstimator { E_OLS = 0, E_GLM = 1, E_MLE = 2 };
enum McpCorrection { C_NONE = 0, C_BONFERRONI = 1, C_HOLM = 2, C_BH = 3, C_TUKEY = 4 };

// One "fit variant" = the set of kept design columns for a sparse-factor-exclusion mask
// (batch.rs:2502-2562) intersected with the test design (`fit_columns`).
struct McpVariant {
    int32_t p_red;                  // kept column count
    int32_t valid;                  // 0 → df < 1 for some N or never reachable
    int8_t kept[MCP_MAX_P];         // kept[r] = full column index
    int8_t remap[MCP_MAX_P];        // full → reduced index or -1
};

struct DevicePlan {
    // ---- shape
    int32_t k, n_fd, n_factors, n_inter, P, T, n_marg, n_contrasts;
    int32_t n_ss, max_n;
    int32_t estimator, outcome, link_probit, correction;
    int32_t report_overall, min_level_count;
    // ---- design
    double L[MCP_MAX_K * MCP_MAX_K];   // optimistic path: lower Cholesky of correlation, row-major L[j][m]
    double corr[MCP_MAX_K * MCP_MAX_K];// base correlation, column-major (scenario path)
    int32_t l_identity;                // L == I structurally (skip the mix, bit-exact: adds of ±0)
    int32_t all_normal;                // every marginal is Normal on the optimistic path
    int32_t var_type[MCP_MAX_K];
    uint8_t var_pinned[MCP_MAX_K];
    double var_param[MCP_MAX_K];
    int32_t f_levels[MCP_MAX_FACTORS];
    int32_t f_off[MCP_MAX_FACTORS];    // offset into f_props
    int32_t f_col[MCP_MAX_FACTORS];    // first dummy column
    int32_t f_sampled[MCP_MAX_FACTORS];
    double f_props[MCP_MAX_LEVELS];
    int32_t n_sampled_factors;         // sampled-factor uniforms consumed per row on the sequential stream
    int32_t seq_per_row;               // all per-row words of the sequential stream: n_sampled_factors + bootstrap
    // ---- uploaded-data arms (data_gen.rs:553-604 strict bootstrap, 1156-1175 / 1262-1277 NORTA)
    int32_t bootstrap;                 // bootstrap_frame_map non-empty: one source-row pick per design row
    int32_t un_rows, un_cols;          // upload_normal shape (sorted standardised values, column j = non-factor column j)
    int32_t ud_rows, ud_cols;          // upload_data shape (the raw frame)
    int32_t frame_map[MCP_MAX_K + MCP_MAX_FACTORS];  // frame column copied into non-factor column c / factor c − k, −1 = none
    double var_param_bs[MCP_MAX_K];    // var_params with the bootstrap arm's unwrap_or(0.0) default (data_gen.rs:577)
    const double* upload_normal;       // [un_rows][un_cols]
    const double* upload_data;         // [ud_rows][ud_cols]
    int8_t col_factor[MCP_MAX_P];      // design column j is a dummy of factor col_factor[j] (−1: not a dummy) ...
    int8_t col_level[MCP_MAX_P];       // ... and indicates level col_level[j] (1..levels−1)
    int32_t inter_ptr[MCP_MAX_INTER + 1];
    int32_t inter_cols[MCP_MAX_INTER_COMP];
    double beta[MCP_MAX_P];
    uint32_t beta_nonzero_mask;        // bit j set ⇔ beta[j] != 0
    int32_t n_nz;                      // the non-zero effects, ascending column order: the linear predictor only visits these ...
    int8_t nz_col[MCP_MAX_P];          // ... (beta_j == 0 terms add ±0 to a +0-seeded sum: skipping them is bit-exact)
    double nz_beta[MCP_MAX_P];
    int32_t target[MCP_MAX_T];         // marginal target columns
    int32_t contrast[2 * MCP_MAX_T];   // (pos, neg) pairs
    int32_t n_ph;                      // posthoc pairwise contrasts, flat over blocks (posthoc.rs:40-148)
    int32_t ph_contrast[2 * MCP_MAX_PH]; // (pos, neg) columns, neg < 0 → the factor's reference level
    // ---- residual / outcome
    int32_t residual_dist, residual_pinned;
    double residual_df;
    // ---- cluster (random intercept only)
    int32_t has_cluster, sizing_fixed_size, sizing_value, n_clusters;
    double tau_squared;
    // ---- scenario knobs (spec.rs:117-176)
    int32_t optimistic;
    double heterogeneity, hsk_gamma, corr_noise_sd, dist_change_prob, resid_change_prob;
    int coop_marginals;  // lane-quad scenario kernels: deal the warp's non-Normal (study, column) blocks to all 32 lanes (kernels_q4.cuh)
    int32_t n_new_dists, new_dists[MCP_MAX_MENU];
    int32_t n_resid_dists, resid_dists[MCP_MAX_MENU];
    int32_t has_lme, re_dist;          // scenario.lme (spec.rs:175): random-effect distribution (a McpRes code), ...
    double re_df, icc_noise_sd;        // ... its df, and the sd of the per-study jitter added to tau² (data_gen.rs:1065-1075)
    int32_t re_words;                  // sequential-stream words one random-effect draw consumes (draw_residual, data_gen.rs:1289-1320)
    // ---- random slopes on the primary grouping (cluster.slopes[k], cluster_slope_design_cols[k]; data_gen.rs:163-213, 836-842)
    int32_t n_slopes;                  // 0, 1 or 2 (MCP_MAX_SLOPES)
    int32_t slope_col[2];              // design column each slope effect multiplies
    double slope_variance[2], slope_corr[2];  // tau_k², corr(intercept, slope k): D = diag(tau) R diag(tau) ...
    double slope_corr12;               // ... and corr(slope 0, slope 1) (slopes[1].corr_with[0], generation.rs:225-240)
    // ---- extra groupings, random intercepts only (cluster.extra_groupings[g]; generation.rs:243-430, data_gen.rs:214-236, 818-835)
    int32_t n_extra;                   // 0, 1 or 2
    int32_t extra_nested[2];           // 0 = Crossed { n_clusters }, 1 = NestedWithin { n_per_parent }
    int32_t extra_count[2];            // n_clusters resp. n_per_parent
    int32_t extra_levels[2];           // levels within max_n rows (extra_n_levels_at): the length of the grouping's draw block
    double extra_tau_squared[2];
    // ---- cluster-level predictors (spec.between_var_indices, data_gen.rs:605-640)
    int32_t n_between;                 // 0 unless the spec has a cluster
    int8_t between_cols[MCP_MAX_P];    // design columns (>= 1) broadcast from the cluster's representative row
    float* between_tab;                // wave scratch [cluster][between column][study in wave]
    uint32_t between_stride;           // studies per wave (the table's innermost extent)
    int32_t apply_hsk, hsk_driver;     // driver < 0 → linear predictor
    double hsk_center, hsk_inv_std, hsk_inv_norm;
    // ---- variants + thresholds
    int32_t n_variants;                // 1 << n_factors when exclusion is live, else 1
    int32_t base_variant;              // variant used when no factor is excluded (fit_columns may reduce it)
    int32_t exact_walk;                // every factor uses the deterministic walk
    // ---- device pointers (filled by the host before upload)
    const uint32_t* ss;                // [n_ss]
    const McpVariant* variants;        // [n_variants]
    const double* crit_unc;            // [n_variants][n_ss]
    const double* crit_corr;           // [n_variants][n_ss][T]
    const double* crit_overall;        // [n_variants][n_ss]
    const double* crit_joint;          // [n_variants][n_ss] joint Wald chi-square threshold (MLE arm) or NULL
    const double* crit_ph_unc;         // [n_variants][n_ss] posthoc-alpha threshold or NULL
    const double* crit_ph_corr;        // [n_variants][n_ss][n_ph] per-block correction rows, concatenated
    const uint8_t* levels;             // [n_factors][max_n] exact-walk levels (NULL if none exact)
    const uint32_t* exact_counts;      // [n_ss][MCP_MAX_LEVELS] per-level counts of the first ss[g] rows when every factor
                                       // uses the exact walk (sim-invariant; consumed by ols_solve_kernel) or NULL
    const double* t3_table;            // [2048] censored t(3) PPF (marginals/t3.rs)
    // ---- outputs
    unsigned long long* counters;      // [n_ss][counter_width] (COUNTS)
    uint8_t* out_unc;                  // [n_sims][n_ss][T]  (BYTES; may be NULL)
    uint8_t* out_cor;
    uint8_t* out_conv;                 // [n_sims][n_ss]
    uint8_t* out_overall;              // [n_sims][n_ss] or NULL
    uint8_t* out_fx;                   // [n_sims][n_ss][n_factors] or NULL
    uint8_t* out_joint;                // [n_sims][n_ss] joint Wald decision (MLE arm) or NULL
    uint8_t* out_bh;                   // [n_sims][n_ss] boundary_hit code (MLE arm) or NULL
    uint8_t* out_ph_unc;               // [n_sims][n_ss][n_ph] or NULL
    uint8_t* out_ph_cor;
    double* out_tsq;                   // [n_sims][n_ss][T] capture or NULL
    uint32_t out_stride;               // 0: the out_* tensors are sim-major, [sim][grid point][..]; S > 0: they are one wave's
                                       // grid-point-major staging, [grid point][S sims][..] — the studies of a warp then write
                                       // neighbouring bytes — and the host transposes the wave into the sim-major tensors
    int* error_flag;                   // device int: 0 ok, 2 = CorrelationNotPSD raised by a sim
    int32_t counter_width;             // 4T + 6 + n_factors + 4 n_ph
    int32_t off_hist_cor, off_fx, off_joint, off_bh, off_ph_unc, off_ph_cor, off_sing, off_fsep, off_pin;  // counter-row offsets (below)
    int32_t smem_counters;             // counters staged in shared memory (fits)
    // ---- progress / cancellation mailbox (ProgressSink, spec.rs:38-52); all NULL / 0 when no sink is attached
    unsigned int* prog_counter;        // device counter of finished work units (OLS: studies generated; GLM / MLE: fits)
    volatile unsigned int* prog_host;  // mapped pinned host word: the counter's value whenever it crosses a multiple of prog_step
    uint32_t prog_step;                // work units between two published values (~1/50 of the call)
    const int* cancel_flag;            // device int set by the host on cancellation: CTAs that have not started yet exit at once
};

// counter row layout (per sample-size index): what aggregate_batch derives
// (crates/engine-orchestrator/src/aggregation.rs:91-242); J = T + n_ph tests enter the joint histograms
//   [0,T)                    uncorrected successes per target
//   [T,2T)                   corrected successes per target
//   2T                       converged fits
//   2T+1                     overall-test successes
//   [2T+2, 2T+3+J)           histogram of #uncorrected-significant tests per sim (targets + posthoc contrasts, J+1 buckets)
//   [off_hist_cor, +J+1)     histogram of #corrected-significant tests per sim
//   [off_fx, +F)             sims with factor f sparse-excluded (factor_excluded code 1: factor_exclusion_counts)
//   off_joint                joint Wald successes (MLE arm; 0 otherwise)
//   off_bh                   fits with boundary_hit != 0 (MLE arm; 0 otherwise)
//   [off_ph_unc, +n_ph)      posthoc uncorrected successes per contrast
//   [off_ph_cor, +n_ph)      posthoc corrected successes per contrast
//   off_sing                 converged fits pinned at the variance boundary (boundary_hit == 1: singular_count)
//   off_pin + k, k < 5       converged fits whose variance component k sat at 0 (pinned_components bit k, spec.rs:566-574; order
//                            intercept, slopes, extra groupings of the FITTED model)
//   [off_fsep, +F)           sims where the GLM separation fallback dropped factor f (code 2: factor_separation_counts)
#define MCP_MAX_PIN 5  // variance components a mixed-model fit reports pins for: intercept, two slopes, two extra groupings
static inline int mcp_counter_width(int T, int F, int n_ph) { return 4 * T + 7 + 2 * F + 4 * n_ph + MCP_MAX_PIN; }
