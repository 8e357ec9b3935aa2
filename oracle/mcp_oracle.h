/*
 * mcp_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of MCPower's per-simulation hot loop
 * (reference: crates/engine-core/src/{rng,data_gen,batch,critvals,correction,
 * distributions,scenarios,correlation}.rs).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (mcpower_b200/csrc) never links, includes or calls it.
 *
 * Parity pinning (see oracle/README.md): bit-exact against the reference's
 * RNG goldens (tests/golden_rng.rs, tests/rng_byte_identity.rs) and the saved
 * sim-0 design matrices/outcomes in validation/data/<case>.rds; <=1e-10 relative
 * against the frozen R lm goldens.  Sub-paths with no reference vector are
 * listed as "unpinned" in DESIGN.md.
 */
#ifndef MCP_ORACLE_H
#define MCP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enum codes shared with the Python test harness (tests/orc.py) */
enum { ORC_DIST_NORMAL = 0, ORC_DIST_BINARY = 1, ORC_DIST_RIGHT_SKEWED = 2, ORC_DIST_LEFT_SKEWED = 3,
       ORC_DIST_HIGH_KURTOSIS = 4, ORC_DIST_UNIFORM = 5,
       /* engine-contract/src/distribution.rs:23-25 */
       ORC_DIST_UPLOADED_FACTOR = 6, ORC_DIST_UPLOADED_BINARY = 7, ORC_DIST_UPLOADED_DATA = 8 };
enum { ORC_RES_NORMAL = 0, ORC_RES_RIGHT_SKEWED = 1, ORC_RES_LEFT_SKEWED = 2, ORC_RES_HIGH_KURTOSIS = 3,
       ORC_RES_UNIFORM = 4 };
enum { ORC_OUT_CONTINUOUS = 0, ORC_OUT_BINARY = 1, ORC_OUT_COUNT = 2 };
enum { ORC_EST_OLS = 0, ORC_EST_GLM = 1, ORC_EST_MLE = 2 };
enum { ORC_CORR_NONE = 0, ORC_CORR_BONFERRONI = 1, ORC_CORR_HOLM = 2, ORC_CORR_BH = 3, ORC_CORR_TUKEY = 4 };
enum { ORC_OK = 0, ORC_ERR_INVALID_SPEC = 1, ORC_ERR_NOT_PSD = 2, ORC_ERR_RANK = 3, ORC_ERR_CANCELLED = 4 };

typedef struct {
    int32_t n_non_factor, n_factor_dummies, n_factors, n_interactions;
    const double* correlation;        /* k*k column-major (spec.rs:255) */
    const int32_t* var_types;         /* k */
    const uint8_t* var_pinned;        /* k or NULL */
    const double* var_params;         /* k */
    const int32_t* factor_n_levels;   /* n_factors */
    const double* factor_proportions; /* sum(levels) */
    const int32_t* factor_sampled;    /* n_factors: -1 inherit, 0 exact, 1 sampled; or NULL */
    const int32_t* inter_ptr;         /* n_interactions+1 CSR offsets */
    const int32_t* inter_cols;
    const double* effect_sizes;       /* P */
    int32_t n_marginal_targets;
    const int32_t* target_indices;
    int32_t n_contrasts;
    const int32_t* contrast_pairs;    /* 2*n_contrasts (pos, neg) */
    int32_t n_fit_columns;
    const int32_t* fit_columns;
    int32_t correction_method;
    double alpha;
    double posthoc_alpha;             /* NaN = inherit alpha */
    int32_t hsk_driver;               /* -1 = linear predictor */
    int32_t residual_dist, residual_pinned, outcome_kind, link_probit, estimator;
    int32_t report_overall, factor_min_level_count;
    int32_t has_cluster, sizing_fixed_size, sizing_value;
    double tau_squared;
    /* scenario block (spec.rs:117-176) */
    double heterogeneity, hsk_ratio, corr_noise_sd, dist_change_prob;
    int32_t n_new_dists;
    const int32_t* new_dists;
    double resid_change_prob;
    int32_t n_resid_dists;
    const int32_t* resid_dists;
    double residual_df;
    int32_t sampled_factor_proportions;
    int32_t has_lme, re_dist;
    double re_df, icc_noise_sd;
    /* posthoc blocks (spec.rs:90-103, batch.rs:2183-2227): pairwise contrasts of all blocks, flat in block order,
     * c = e_pos - e_neg (neg < 0: the factor's reference level); per block its level count and C(k,2) */
    int32_t n_posthoc;
    const int32_t* posthoc_pairs;           /* 2*n_posthoc (pos, neg) */
    int32_t n_posthoc_blocks;
    const int32_t* posthoc_block_levels;    /* n_posthoc_blocks */
    const int32_t* posthoc_block_ncontrasts;
    /* uploaded-data arms (spec.rs upload_* fields; contract_adapter.rs:359-455).  NORTA: upload_normal is the
     * row-major (un_rows, un_cols) table whose column j holds the sorted standardised values of non-factor
     * column j.  Strict bootstrap: upload_data is the row-major (ud_rows, ud_cols) frame and frame_map[c]
     * (c < k: non-factor column c; c >= k: factor c-k) the frame column copied into it, -1 = none;
     * n_frame_map == 0 means "not bootstrap". */
    const double* upload_normal;
    int32_t un_rows, un_cols;
    const double* upload_data;
    int32_t ud_rows, ud_cols;
    int32_t n_frame_map;
    const int32_t* frame_map;
    int32_t n_var_params;                   /* entries of var_params the spec really carried (bootstrap binary: unwrap_or(0.0)) */
    /* cluster-level predictors (spec.between_var_indices, data_gen.rs:605-640): design columns whose value on every row of a
     * cluster is the value on the cluster's representative (first) row */
    int32_t n_between;
    const int32_t* between_cols;
    /* random slopes on the primary grouping (cluster.slopes[k] + cluster_slope_design_cols[k]; data_gen.rs:163-213, 836-842):
     * RE covariance D = diag(tau) R diag(tau) over [intercept, slope_0, slope_1], R from re_correlation_from_slopes
     * (engine-contract generation.rs:225-240) */
    int32_t n_slopes;               /* 0, 1 or 2 */
    int32_t slope_col[2];           /* design column each slope multiplies */
    double slope_variance[2], slope_corr[2]; /* corr_with_intercept */
    double slope_corr12;            /* slopes[1].corr_with[0] */
    /* extra groupings, random intercepts only (cluster.extra_groupings[g]; generation.rs:243-430, data_gen.rs:214-236, 818-835) */
    int32_t n_extra;                /* 0, 1 or 2 */
    int32_t extra_nested[2];        /* 0 = Crossed { n_clusters }, 1 = NestedWithin { n_per_parent } */
    int32_t extra_count[2];         /* n_clusters resp. n_per_parent */
    double extra_tau_squared[2];
} OrcSpec;

/* ---- RNG primitives (rng.rs:36-128,154-178,354-364) ---- */
uint64_t orc_splitmix64(uint64_t z);
uint64_t orc_pcg_mix64(uint64_t base_seed, uint64_t sim_id);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float orc_u32_to_unit_f32(uint32_t w);
float orc_norm_inv_cdf_f32(float u);
/* sequential stream draws: kind 0 = next_normal, 1 = next_uniform; returns f32 bit patterns */
void orc_seq_stream(uint64_t base_seed, uint64_t sim_id, int kind, int n, uint32_t* out_bits);
void orc_seq_categorical(uint64_t base_seed, uint64_t sim_id, const double* probs, int n_probs, int n, int32_t* out);
void orc_planar_normals(uint64_t base_seed, uint64_t sim_id, uint32_t cls, uint32_t col, int n, uint32_t* out_bits);
void orc_sample_t(uint64_t base_seed, uint64_t sim_id, double df, int n, double* out);
void orc_draw_residual(uint64_t base_seed, uint64_t sim_id, int dist, double df, int n, double* out);

/* per-study lower Cholesky factor of the (perturbed, PSD-repaired) correlation, column-major k x k */
/* Random intercept + one random slope by profiled REML on supplied data (ids = cluster of each row).  vc = {var_intercept,
 * var_slope, corr, sigma²}. */
int orc_fit_lmm_slope(const double* xc, const double* yv, const int32_t* ids, int n, int p, int slope_col, double* betas,
                      double* se, double* vc, int* converged, int* boundary, int* n_eval);
/* ... with n_slopes (1 or 2) slopes, q = 1 + n_slopes.  vc = {variances[q] (intercept first), correlations (1,0) [(2,0) (2,1)],
 * sigma²}. */
int orc_fit_lmm_slopes(const double* xc, const double* yv, const int32_t* ids, int n, int p, int n_slopes, const int32_t* slope_cols,
                       double* betas, double* se, double* vc, int* converged, int* boundary, int* n_eval);
/* General structure: primary grouping (ids[0], intercept + n_slopes slopes) plus n_extra intercept-only groupings (ids[1 + g],
 * n x 1 each, 0-based levels).  vc = {primary variances[q], primary correlations, extra variances[n_extra], sigma²}. */
int orc_fit_lmm_general(const double* xc, const double* yv, const int32_t* ids, int n, int p, int n_slopes, const int32_t* slope_cols,
                        int n_extra, double* betas, double* se, double* vc, int* converged, int* boundary, int* n_eval);
/* level of row i in extra grouping g (ClusterSpec::extra_level_of_row) and the number of levels within the first n rows */
int orc_extra_level_of_row(const OrcSpec* s, int g, int i);
int orc_extra_levels_at(const OrcSpec* s, int g, int n);
int orc_design_factor(const OrcSpec* s, uint64_t base_seed, uint64_t sim_id, double* l_colmajor);

/* ---- transforms (distributions.rs:32-56, data_gen.rs:1106-1182, marginals/t3.rs) ---- */
double orc_exp(double x);
double orc_log(double x);
double orc_phi(double z);
double orc_apply_marginal(double z, int dist, double param);
/* the uploaded-data marginals as well (data_gen.rs:1156-1180): `table` = upload_normal (rows x cols, row-major), column `col` */
double orc_apply_marginal_upload(double z, int dist, double param, const double* table, int rows, int cols, int col);
double orc_t3_lookup(double percentile);
double orc_poisson_quantile(double z, double lambda);

/* ---- critical values (critvals.rs) ---- */
double orc_norm_ppf(double p);
double orc_t_ppf(double p, double df);
double orc_chi2_ppf(double p, double k);
double orc_f_ppf(double p, double dfn, double dfd);
double orc_q_tukey_ppf(double p, double k, double df);
/* per-N rows: t_unc[n_ss], corr[n_ss*m], posthoc[n_ss], overall[n_ss]; returns ORC_* */
int orc_build_crit_table(const OrcSpec* spec, const uint32_t* sample_sizes, int n_ss, int n_predictors,
                         double* t_unc_sq, double* corr_sq, double* posthoc_sq, double* overall);

/* ---- corrections (correction.rs:16-115) ---- */
void orc_apply_correction(int method, const double* t_sq, int t, double crit_unc, const double* crit_sq, uint8_t* out);

/* ---- data generation (data_gen.rs:87-925): x col-major f32 [max_n x P], y f32 ---- */
int orc_generate_sim(const OrcSpec* spec, uint64_t base_seed, uint64_t sim_id, int max_n, float* x, float* y,
                     int32_t* cluster_ids /* max_n or NULL */);

/* ---- OLS on supplied data (introspect.rs:174-293): col-major f64 design, narrowed to f32 first ---- */
int orc_fit_ols(const double* x_colmajor, const double* y, int n, int p, double* betas, double* se, double* t_sq,
                double* sigma_sq, double* sst, double* rss, int* converged);

/* ---- GLM IRLS on supplied data; family 0 logit, 1 probit, 2 poisson ---- */
int orc_fit_glm(const double* x_colmajor, const double* y, int n, int p, int family, const double* beta_start,
                double* betas, double* se, double* deviance, double* deviance_null, int* converged, int* iters);

/* ---- random-intercept LMM (profiled REML) on supplied data; ids = 0-based cluster index per row ---- */
int orc_fit_lmm(const double* x_colmajor, const double* y, const int32_t* ids, int n, int p, double* betas, double* se,
                double* theta, double* sigma_sq, int* converged, int* boundary);

/* ---- the batch driver (batch.rs:199-776, 795-1213) ----
 * Outputs (caller-allocated, sim-major like spec.rs:552-600):
 *   uncorrected/corrected [n_sims][n_ss][T], converged [n_sims][n_ss],
 *   overall [n_sims][n_ss] (or NULL), factor_excluded [n_sims][n_ss][n_factors] (or NULL),
 *   t_sq_capture [n_sims][n_ss][T] f64 (or NULL; NaN where not converged/excluded).
 * sim ids run over [sim_begin, sim_begin + n_sims).  n_threads<=1 → serial. */
int orc_run_batch(const OrcSpec* spec, const uint32_t* sample_sizes, int n_ss, uint32_t n_sims, uint64_t sim_begin,
                  uint64_t base_seed, int n_threads, uint8_t* uncorrected, uint8_t* corrected, uint8_t* converged,
                  uint8_t* overall, uint8_t* factor_excluded, double* t_sq_capture,
                  uint8_t* joint /* [n_sims][n_ss] or NULL (MLE joint Wald) */, uint8_t* boundary_hit /* or NULL */,
                  uint8_t* posthoc_unc /* [n_sims][n_ss][n_posthoc] or NULL */, uint8_t* posthoc_cor);
/* per-N posthoc correction rows, [n_ss][n_posthoc]: one table per block (alpha = posthoc alpha, family = the block's
 * contrasts, Tukey k = the block's levels), concatenated (batch.rs:2252-2405) */
int orc_posthoc_corr_rows(const OrcSpec* spec, const uint32_t* sample_sizes, int n_ss, int n_predictors, double* out);

const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
