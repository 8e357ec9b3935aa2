// kernels.cuh — the sm_100a hot path: one simulated study per thread, generated and
// fitted entirely on-chip.
//
// Replaces engine_core::batch::run_one_sim (crates/engine-core/src/batch.rs:795-1213, OLS arm)
// and generate_sim_data (data_gen.rs:87-925).  Mapping:
//   * thread  = one simulation (global sim id → Philox key, rng.rs:36-69); a warp is 32
//     independent studies in lock-step, so there is no cross-lane reduction, no idle lane
//     in the P×P solve, and the integer outputs are independent of launch geometry.
//   * rows are produced four at a time (one Philox4x32 block per column per 4 rows,
//     rng.rs:130-137) into a per-thread f32 slab in shared memory — the reference's f32
//     "generation plane" — then widened to f64 registers row by row (batch.rs:828-851).
//   * X'X, X'y, y'y, Σy live in registers for the whole study (P(P+1)/2 + P + 2 doubles,
//     P is a template parameter); nested sample sizes reuse the running sums
//     (batch.rs:871-917) and re-solve at every grid point.
//   * the inverse-normal tail branch (4.85 % of draws, rng.rs:354-364) is deferred: the
//     central polynomial runs branch-free for every draw and tail lanes are patched from a
//     per-thread bit mask afterwards, so a warp does not pay the tail path for every draw.
// Generated data never touches HBM; the only global writes are counters (COUNTS mode) or the
// 0/1 byte tensors of BatchResult (spec.rs:552-600).
#pragma once
#include <cuda_runtime.h>

#include "device_math.cuh"
#include "plan.h"

namespace mcp {

// c_plan (the __constant__ DevicePlan, one copy per translation unit) is declared in device_math.cuh.

#ifndef MCP_BLOCK
#define MCP_BLOCK 128
#endif
#ifndef MCP_MIN_BLOCKS
#define MCP_MIN_BLOCKS 1
#endif
#ifndef MCP_UNROLL_TASKS
#define MCP_UNROLL_TASKS 1
#endif
#define MCP_FNZ 1e-30   /* lib.rs:46 FLOAT_NEAR_ZERO */
#define MCP_EPS_RANK 1e-12 /* batch.rs:41 */

__device__ __forceinline__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }
constexpr int kUnrollTasks = MCP_UNROLL_TASKS;

// ------------------------------------------------------------------------------------------
// Per-fit results shared by the templated (registers) and generic (local memory) solvers.
struct FitOut {
    double sigma_sq, sst, rss, df_resid;
    int converged;
};

// Cholesky + solves on a packed lower-triangular system held in registers (all indices static).
// The operation order here is the parity contract of the test oracle's solver: bit-identical t².
template <int PT>
struct RegSolver {
    static constexpr int NACC = PT * (PT + 1) / 2;
    double L[NACC];
    double inv_d[PT];
    double beta[PT];
    FitOut out;

    __device__ __forceinline__ void factor_and_solve(const double (&A)[NACC], const double (&b)[PT], double yty, double sum_y,
                                                     uint32_t n_rows) {
        int ok = 1;
        double dmin = __longlong_as_double(0x7ff0000000000000LL), dmax = 0.0;
#pragma unroll
        for (int i = 0; i < PT; i++) {
#pragma unroll
            for (int j = 0; j <= i; j++) {
                double s = A[tri(i, j)];
#pragma unroll
                for (int k = 0; k < j; k++) s = fma(-L[tri(i, k)], L[tri(j, k)], s);
                if (i == j) {
                    if (!(s > 0.0) || isinf(s)) {
                        ok = 0;
                        s = 1.0;
                    }
                    const double d = sqrt(s);
                    L[tri(i, i)] = d;
                    inv_d[i] = 1.0 / d;
                    dmin = d < dmin ? d : dmin;
                    dmax = d > dmax ? d : dmax;
                } else {
                    L[tri(i, j)] = s * inv_d[j];
                }
            }
        }
        if (ok && !(dmin / dmax >= MCP_EPS_RANK)) ok = 0;
        double w[PT];
#pragma unroll
        for (int i = 0; i < PT; i++) {
            double s = b[i];
#pragma unroll
            for (int k = 0; k < i; k++) s = fma(-L[tri(i, k)], w[k], s);
            w[i] = s * inv_d[i];
        }
#pragma unroll
        for (int i = PT - 1; i >= 0; i--) {
            double s = w[i];
#pragma unroll
            for (int k = i + 1; k < PT; k++) s = fma(-L[tri(k, i)], beta[k], s);
            beta[i] = s * inv_d[i];
        }
        double rss = yty;
#pragma unroll
        for (int i = 0; i < PT; i++) rss = fma(-w[i], w[i], rss);
        const double n = (double)n_rows;
        out.converged = ok;
        out.rss = rss;
        out.sst = yty - sum_y * sum_y / n;
        out.df_resid = n - (double)PT;
        out.sigma_sq = rss / out.df_resid;
    }

    // t² of c'β for c = e_pos − e_neg (neg < 0 → e_pos): (c'β)² / (σ² ‖L⁻¹c‖²)   (posthoc.rs:3-8)
    __device__ __forceinline__ double t_sq(int pos, int neg) const {
        double u[PT];
        double q = 0.0, est = 0.0;
#pragma unroll
        for (int i = 0; i < PT; i++) {
            double s = (i == pos ? 1.0 : 0.0) - (i == neg ? 1.0 : 0.0);
#pragma unroll
            for (int k = 0; k < i; k++) s = fma(-L[tri(i, k)], u[k], s);
            u[i] = s * inv_d[i];
            q = fma(u[i], u[i], q);
            if (i == pos) est += beta[i];
        }
        if (neg >= 0) {
            double bn = 0.0;
#pragma unroll
            for (int i = 0; i < PT; i++)
                if (i == neg) bn = beta[i];
            est = est - bn;
        }
        const double var = out.sigma_sq * q;
        if (!(var > MCP_FNZ) || isinf(var)) return __longlong_as_double(0x7ff8000000000000LL);
        return est * est / var;
    }

    // Var(beta_pos) = sigma² ‖L⁻¹ e_pos‖² (fit_suff_stats_t_sq's var_diag; the introspection twins report its square root)
    __device__ __forceinline__ double var_of(int pos) const {
        double u[PT];
        double q = 0.0;
#pragma unroll
        for (int i = 0; i < PT; i++) {
            double s = i == pos ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; k++) s = fma(-L[tri(i, k)], u[k], s);
            u[i] = s * inv_d[i];
            q = fma(u[i], u[i], q);
        }
        return out.sigma_sq * q;
    }
};

// Generic solver on local-memory arrays (runtime p): sparse-factor-exclusion / test-design
// reduced fits (batch.rs:970-1013).  Cold path.
template <int PMAX>
struct MemSolver {
    double L[PMAX * (PMAX + 1) / 2];
    double inv_d[PMAX];
    double beta[PMAX];
    double w[PMAX];
    FitOut out;
    int p;

    __device__ __noinline__ void factor_and_solve(const double* A, const double* b, int p_, double yty, double sum_y, uint32_t n_rows) {
        p = p_;
        int ok = 1;
        double dmin = __longlong_as_double(0x7ff0000000000000LL), dmax = 0.0;
        for (int i = 0; i < p; i++) {
            for (int j = 0; j <= i; j++) {
                double s = A[tri(i, j)];
                for (int k = 0; k < j; k++) s = fma(-L[tri(i, k)], L[tri(j, k)], s);
                if (i == j) {
                    if (!(s > 0.0) || isinf(s)) {
                        ok = 0;
                        s = 1.0;
                    }
                    const double d = sqrt(s);
                    L[tri(i, i)] = d;
                    inv_d[i] = 1.0 / d;
                    dmin = d < dmin ? d : dmin;
                    dmax = d > dmax ? d : dmax;
                } else {
                    L[tri(i, j)] = s * inv_d[j];
                }
            }
        }
        if (ok && !(dmin / dmax >= MCP_EPS_RANK)) ok = 0;
        for (int i = 0; i < p; i++) {
            double s = b[i];
            for (int k = 0; k < i; k++) s = fma(-L[tri(i, k)], w[k], s);
            w[i] = s * inv_d[i];
        }
        for (int i = p - 1; i >= 0; i--) {
            double s = w[i];
            for (int k = i + 1; k < p; k++) s = fma(-L[tri(k, i)], beta[k], s);
            beta[i] = s * inv_d[i];
        }
        double rss = yty;
        for (int i = 0; i < p; i++) rss = fma(-w[i], w[i], rss);
        const double n = (double)n_rows;
        out.converged = ok;
        out.rss = rss;
        out.sst = yty - sum_y * sum_y / n;
        out.df_resid = n - (double)p;
        out.sigma_sq = rss / out.df_resid;
    }

    __device__ __noinline__ double t_sq(int pos, int neg) {
        int start = pos;
        if (neg >= 0 && neg < start) start = neg;
        double q = 0.0;
        for (int i = start; i < p; i++) {
            double s = (i == pos ? 1.0 : 0.0) - (i == neg ? 1.0 : 0.0);
            for (int k = start; k < i; k++) s = fma(-L[tri(i, k)], w[k], s);  // w reused as u
            w[i] = s * inv_d[i];
            q = fma(w[i], w[i], q);
        }
        const double est = neg >= 0 ? beta[pos] - beta[neg] : beta[pos];
        const double var = out.sigma_sq * q;
        if (!(var > MCP_FNZ) || isinf(var)) return __longlong_as_double(0x7ff8000000000000LL);
        return est * est / var;
    }

    __device__ __noinline__ double var_of(int pos) {  // sigma² ‖L⁻¹ e_pos‖²; clobbers w like t_sq
        double q = 0.0;
        for (int i = pos; i < p; i++) {
            double s = i == pos ? 1.0 : 0.0;
            for (int k = pos; k < i; k++) s = fma(-L[tri(i, k)], w[k], s);
            w[i] = s * inv_d[i];
            q = fma(w[i], w[i], q);
        }
        return out.sigma_sq * q;
    }
};

// ------------------------------------------------------------------------------------------
// Significance decisions + corrections (batch.rs:1089-1142, correction.rs:16-115) and output.
// `all_t` lives in local memory (T is a runtime count); executed once per fit.
__device__ __forceinline__ bool gt_crit(double t, double c) { return !(t != t) && t > c; }

// order[] = indices sorted by t² descending, NaN last, ties by ascending index (stable insertion sort)
static __device__ __noinline__ void sort_desc(const double* t_sq, int n, int8_t* order) {
    for (int i = 0; i < n; i++) order[i] = (int8_t)i;
    for (int i = 1; i < n; i++) {
        const int v = order[i];
        const double b = t_sq[v];
        const bool b_nan = b != b;
        int j = i - 1;
        while (j >= 0) {
            const double a = t_sq[order[j]];
            const bool a_nan = a != a;
            bool move;
            if (a_nan && b_nan) move = false;
            else if (a_nan) move = true;
            else if (b_nan) move = false;
            else move = a < b || (a == b && signbit(a) && !signbit(b));
            if (!move) break;
            order[j + 1] = order[j];
            j--;
        }
        order[j + 1] = (int8_t)v;
    }
}

static __device__ __noinline__ void apply_correction(int method, const double* t_sq, int T, double crit_unc, const double* __restrict__ crit_row,
                                              uint8_t* out) {
    for (int i = 0; i < T; i++) out[i] = 0;
    if (method == C_NONE) {
        for (int i = 0; i < T; i++) out[i] = gt_crit(t_sq[i], crit_unc);
    } else if (method == C_BONFERRONI || method == C_TUKEY) {
        for (int i = 0; i < T; i++) out[i] = gt_crit(t_sq[i], crit_row[i]);
    } else {
        int8_t order[MCP_MAX_T];
        sort_desc(t_sq, T, order);
        if (method == C_HOLM) {
            for (int k = 0; k < T; k++) {
                if (gt_crit(t_sq[order[k]], crit_row[k])) out[order[k]] = 1;
                else break;
            }
        } else {
            int last = -1;
            for (int k = 0; k < T; k++)
                if (gt_crit(t_sq[order[k]], crit_row[k])) last = k;
            for (int k = 0; k <= last; k++) out[order[k]] = 1;
        }
    }
}

// ProgressSink plumbing (spec.rs:38-52, batch.rs:2149-2177).  The host never cuts a batch into small launches to report
// progress: one lane per warp adds its finished work units to a device counter, and the warp whose add crosses a multiple of
// prog_step publishes the new value to a mapped pinned host word that the waiting host thread polls (~50 stores per call
// travel over PCIe).  Cancellation is the reverse direction: the host sets a device flag, CTAs that have not started exit.
__device__ __forceinline__ bool mcp_cancelled() {
    const int* f = c_plan.cancel_flag;
    return f != nullptr && *reinterpret_cast<const volatile int*>(f) != 0;
}
__device__ __forceinline__ void mcp_report_units(unsigned int units) {  // call from ONE lane per warp
    unsigned int* ctr = c_plan.prog_counter;
    if (ctr == nullptr || units == 0u) return;
    const unsigned int old = atomicAdd(ctr, units), now = old + units, step = c_plan.prog_step;
    if (now / step != old / step) {
        *c_plan.prog_host = now;
        __threadfence_system();
    }
}
// units finished by the warp whose first unit has index `first`, `per_warp` units per warp, `total` units in the launch
__device__ __forceinline__ void mcp_report_warp(uint32_t first, uint32_t per_warp, uint32_t total) {
    __syncwarp();
    if ((threadIdx.x & 31u) == 0u && first < total) mcp_report_units(total - first < per_warp ? total - first : per_warp);
}

// Warp-aggregated counter add: one atomic per cell per convergence group.  The lanes that execute this call together need not be
// the whole warp, and — after divergent per-fit searches of different length — lanes that sit in different iterations of the
// caller's loops can be scheduled together at this instruction: the group is therefore split by the cell the lanes count into
// (match.any on the address) before the ballot is folded.
__device__ __forceinline__ void count_flag(unsigned int* cnt_s, unsigned long long* cnt_g, int idx, bool flag) {
    const unsigned am = __activemask();
    const unsigned long long cell = cnt_s ? (unsigned long long)(cnt_s + idx) : (unsigned long long)(cnt_g + idx);
    const unsigned peers = __match_any_sync(am, cell);
    const unsigned b = __ballot_sync(am, flag) & peers;
    if (b != 0u && (threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) {
        if (cnt_s) atomicAdd(&cnt_s[idx], (unsigned)__popc(b));
        else atomicAdd(&cnt_g[idx], (unsigned long long)__popc(b));
    }
}

struct SimOut {
    unsigned int* cnt_s;  // shared-memory counters or NULL
    uint64_t sim_local;   // index into the byte tensors
};

// Emit one fit: decisions, corrections, overall F, counters/bytes.  all_t = T squared statistics (NaN = untestable).
// ov_given < 0: compute the overall F test from (sst, rss); otherwise 0/1 decided by the caller (GLM: deviance LRT).
static // joint_sig / boundary: MLE arm only (joint Wald chi-square decision, boundary_hit code; spec.rs:562-582).
__device__ __noinline__ void emit_fit(const SimOut& so, int g, int variant, const FitOut& fit, int p_red, double* all_t,
                                      const uint8_t* excl_flags, int ov_given = -1, int joint_sig = 0, int boundary = 0,
                                      const double* ph_t = nullptr) {
    const DevicePlan& pl = c_plan;
    const int T = pl.T;
    const size_t vrow = (size_t)variant * pl.n_ss + g;
    const double crit_unc = pl.crit_unc[vrow];
    uint8_t unc[MCP_MAX_T], cor[MCP_MAX_T];
    int n_unc = 0, n_cor = 0;
    if (fit.converged) {
        for (int t = 0; t < T; t++) {
            unc[t] = gt_crit(all_t[t], crit_unc);
            n_unc += unc[t];
        }
        apply_correction(pl.correction, all_t, T, crit_unc, pl.crit_corr + vrow * T, cor);
        for (int t = 0; t < T; t++) n_cor += cor[t];
    } else {
        for (int t = 0; t < T; t++) {
            unc[t] = 0;
            cor[t] = 0;
            all_t[t] = __longlong_as_double(0x7ff8000000000000LL);
        }
    }
    // ---- posthoc pairwise contrasts (evaluate_posthoc, posthoc.rs:40-148): uncorrected at the posthoc-alpha threshold,
    // corrected with the main correction method over the flat contrast set and the per-block threshold rows
    const int NPH = ph_t ? pl.n_ph : 0;
    uint8_t ph_unc[MCP_MAX_PH], ph_cor[MCP_MAX_PH];
    if (NPH > 0) {
        for (int c = 0; c < NPH; c++) ph_unc[c] = ph_cor[c] = 0;
        if (fit.converged) {
            const double ph_crit = pl.crit_ph_unc[vrow];
            for (int c = 0; c < NPH; c++) ph_unc[c] = gt_crit(ph_t[c], ph_crit);
            apply_correction(pl.correction, ph_t, NPH, ph_crit, pl.crit_ph_corr + vrow * NPH, ph_cor);
        }
        for (int c = 0; c < NPH; c++) {
            n_unc += ph_unc[c];
            n_cor += ph_cor[c];
        }
    }
    uint8_t ov = 0;
    if (ov_given >= 0) {
        ov = (uint8_t)(pl.report_overall ? ov_given : 0);
    } else if (pl.report_overall && fit.converged) {
        const double dfn = (double)(p_red - 1), dfd = fit.df_resid;
        if (dfn >= 1.0 && fit.sst > 1e-10 && fit.rss > 0.0 && !isinf(fit.rss) && fit.rss == fit.rss && !isinf(fit.sst) &&
            fit.sst == fit.sst) {
            const double f = ((fit.sst - fit.rss) / dfn) / (fit.rss / dfd);
            if (f == f && !isinf(f) && f > pl.crit_overall[vrow]) ov = 1;
        }
    }
    // ---- counters (what aggregate_batch derives, aggregation.rs:91-242)
    unsigned int* cs = so.cnt_s ? so.cnt_s + (size_t)g * pl.counter_width : nullptr;
    unsigned long long* cg = pl.counters + (size_t)g * pl.counter_width;
    for (int t = 0; t < T; t++) {
        count_flag(cs, cg, t, unc[t] != 0);
        count_flag(cs, cg, T + t, cor[t] != 0);
    }
    count_flag(cs, cg, 2 * T, fit.converged != 0);
    count_flag(cs, cg, 2 * T + 1, ov != 0);
    if (cs) {  // joint histograms pool the targets and the posthoc contrasts (aggregation.rs:186-217)
        atomicAdd(&cs[2 * T + 2 + n_unc], 1u);
        atomicAdd(&cs[pl.off_hist_cor + n_cor], 1u);
    } else {
        atomicAdd(&cg[2 * T + 2 + n_unc], 1ull);
        atomicAdd(&cg[pl.off_hist_cor + n_cor], 1ull);
    }
    for (int f = 0; f < pl.n_factors; f++) {
        count_flag(cs, cg, pl.off_fx + f, excl_flags && excl_flags[f] == 1);
        count_flag(cs, cg, pl.off_fsep + f, excl_flags && excl_flags[f] == 2);
    }
    count_flag(cs, cg, pl.off_joint, joint_sig != 0);
    // boundary = code (bits 0-1: 0 none, 1 a variance component at 0, 2 optimiser at its bound) | pinned-component mask << 2
    const int bcode = boundary & 3, bpins = boundary >> 2;
    count_flag(cs, cg, pl.off_bh, bcode != 0);
    count_flag(cs, cg, pl.off_sing, fit.converged && bcode == 1);
    if (pl.has_cluster)
        for (int k = 0; k < MCP_MAX_PIN; k++) count_flag(cs, cg, pl.off_pin + k, fit.converged && ((bpins >> k) & 1));
    for (int c = 0; c < NPH; c++) {
        count_flag(cs, cg, pl.off_ph_unc + c, ph_unc[c] != 0);
        count_flag(cs, cg, pl.off_ph_cor + c, ph_cor[c] != 0);
    }
    // ---- BatchResult byte tensors (spec.rs:552-600), sim-major
    // cell index: sim-major, or (waves of a multi-point BYTES / CAPTURE call) the grid-point-major staging layout
    const size_t cell = pl.out_stride ? (size_t)g * pl.out_stride + so.sim_local : so.sim_local * pl.n_ss + g;
    if (pl.out_unc) {
        for (int t = 0; t < T; t++) {
            pl.out_unc[cell * T + t] = unc[t];
            pl.out_cor[cell * T + t] = cor[t];
        }
        pl.out_conv[cell] = (uint8_t)fit.converged;
        if (pl.out_overall) pl.out_overall[cell] = ov;
        if (pl.out_joint) pl.out_joint[cell] = (uint8_t)joint_sig;
        if (pl.out_bh) pl.out_bh[cell] = (uint8_t)boundary;
        if (pl.out_fx)
            for (int f = 0; f < pl.n_factors; f++) pl.out_fx[cell * pl.n_factors + f] = excl_flags ? excl_flags[f] : 0;
    }
    if (pl.out_ph_unc && NPH > 0) {
        for (int c = 0; c < NPH; c++) {
            pl.out_ph_unc[cell * NPH + c] = ph_unc[c];
            pl.out_ph_cor[cell * NPH + c] = ph_cor[c];
        }
    }
    if (pl.out_tsq) {
        for (int t = 0; t < T; t++) pl.out_tsq[cell * T + t] = all_t[t];
    }
}

// ------------------------------------------------------------------------------------------
// Per-sim generation context.
struct SimCtx {
    uint32_t k0, k1;       // Philox key words (rng.rs:62-64)
    SeqRng seq;            // counter-addressed view of the sequential stream
    double tau;            // sqrt(tau²) of the random intercept
    double tau_extra[2];   // ... and of the extra groupings' intercepts
    uint32_t seq_row_base; // first sequential-stream word of the per-row factor draws
    const double* utab = nullptr;  // optional table of the study's cluster effects u_c (shared memory), else drawn per row
    uint32_t sim_local = 0;        // the study's index inside the launch (addresses per-wave scratch)
    double sl[5];                  // rows 1, 2 of chol(D) of this study's random-effect covariance (random slopes): L10 L11 L20 L21 L22
};

// Residuals of the four rows of Philox block `blk` from the CLASS_RESID slots (data_gen.rs:412-484); slot0 = the slot-0
// normals already generated.  ONE Philox call per extra slot serves all four rows and the four inverse-normal chains
// interleave; each row's chi-square is accumulated in slot order, as in the oracle.
static __device__ __noinline__ float4 transform_residual4(int dist, double df_in, float4 slot0, uint32_t blk, uint32_t k0, uint32_t k1) {
    const double df = df_in > 3.0 ? df_in : 3.0;
    long long dfi = llround(df);
    if (dfi < 1) dfi = 1;
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
    const bool is_t = dist == R_HIGH_KURTOSIS;
    if (!is_t) {  // right/left skewed: slots 0..df-1, slot 0 is the draw we already have
        const double g0 = (double)slot0.x, g1 = (double)slot0.y, g2 = (double)slot0.z, g3 = (double)slot0.w;
        c0 += g0 * g0;
        c1 += g1 * g1;
        c2 += g2 * g2;
        c3 += g3 * g3;
    }
    for (long long s = is_t ? 0 : 1; s < dfi; s++) {
        const U4 w = philox4x32_10(blk, (uint32_t)(is_t ? 1 + s : s), 2u, 0u, k0, k1);
        const double g0 = (double)norm_inv_from_word(w.x), g1 = (double)norm_inv_from_word(w.y);
        const double g2 = (double)norm_inv_from_word(w.z), g3 = (double)norm_inv_from_word(w.w);
        c0 += g0 * g0;
        c1 += g1 * g1;
        c2 += g2 * g2;
        c3 += g3 * g3;
    }
    if (is_t) {
        const double scale = 1.0 / sqrt(df / (df - 2.0));
        const double d0 = sqrt(c0 / df), d1 = sqrt(c1 / df), d2 = sqrt(c2 / df), d3 = sqrt(c3 / df);
        const double z0 = (double)slot0.x, z1 = (double)slot0.y, z2 = (double)slot0.z, z3 = (double)slot0.w;
        return make_float4((float)((d0 <= 0.0 ? z0 : z0 / d0) * scale), (float)((d1 <= 0.0 ? z1 : z1 / d1) * scale),
                           (float)((d2 <= 0.0 ? z2 : z2 / d2) * scale), (float)((d3 <= 0.0 ? z3 : z3 / d3) * scale));
    }
    const double scale = 1.0 / sqrt(2.0 * df);
    const double sign = dist == R_LEFT_SKEWED ? -1.0 : 1.0;
    return make_float4((float)(sign * (c0 - df) * scale), (float)(sign * (c1 - df) * scale), (float)(sign * (c2 - df) * scale),
                       (float)(sign * (c3 - df) * scale));
}

// Per-thread view of the shared-memory slab.  Element `off` of thread t lives at smem[off * STRIDE + t]:
// consecutive threads hit consecutive banks, so every per-thread access (static or dynamic offset) is
// conflict-free (the first version used a 36-float thread stride: 58 % of shared wavefronts were conflicts).
template <int STRIDE>
struct Slab {
    float* p;
    __device__ __forceinline__ float& operator[](int off) const { return p[off * STRIDE]; }
};

// ------------------------------------------------------------------------------------------
// Scenario (non-optimistic) path: the per-sim design draw of populate_design (data_gen.rs:990-1093):
// perturbed correlation → PSD repair → Cholesky, marginal swaps, residual swap, per-study β jitter.
// Lives in local memory (runtime k); only instantiated in the SCEN kernels.
template <int PT>
struct ScenDesign {
    static constexpr int KT = PT > 1 ? PT - 1 : 1;
    double L[KT * KT];  // row-major lower Cholesky factor of this sim's correlation
    int vt[KT];         // this sim's marginal types
    int residual_dist;
    int use_het;
    double beta_eff[PT];
    double tau;         // sqrt of this study's effective tau² (base + ICC jitter, clamped at 0; data_gen.rs:1065-1075)
    double tau_extra[2];  // the same per extra grouping: one independent jitter draw each, declaration order (data_gen.rs:1077-1091)
    int ok;
};

// Column-major LLT, left-looking (same order as the host's correlation_llt / data_gen.rs:55-75).
static __device__ __noinline__ int llt_colmajor(double* a, int n) {
    for (int j = 0; j < n; j++) {
        double diag = a[j * n + j];
        for (int m = 0; m < j; m++) diag -= a[m * n + j] * a[m * n + j];
        if (!(diag > 0.0) || isinf(diag)) return -1;
        const double ljj = sqrt(diag);
        a[j * n + j] = ljj;
        for (int i = j + 1; i < n; i++) {
            double s = a[j * n + i];
            for (int m = 0; m < j; m++) s -= a[m * n + i] * a[m * n + j];
            a[j * n + i] = s / ljj;
        }
    }
    for (int j = 0; j < n; j++)
        for (int i = 0; i < j; i++) a[j * n + i] = 0.0;
    return 0;
}

// Cyclic Jacobi eigen-decomposition (row-major a destroyed; eigenvalues w, eigenvectors = columns of v).
// Stand-in for faer's self_adjoint_eigen (correlation.rs:120-127); the oracle runs the same rotations.
static __device__ __noinline__ void jacobi_eigh(double* a, int n, double* w, double* v) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) v[i * n + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = a[p * n + q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                // The matrices live in local memory (runtime n): issue the loads of four rows / columns — and of the
                // eigenvector update, which touches a different array — before their stores, so that eight to sixteen
                // accesses are in flight instead of two.  Same element arithmetic, same results.
                for (int k0 = 0; k0 < n; k0 += 4) {
                    double xp[4], xq[4], vp[4], vq[4];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (k0 + u < n) {
                            xp[u] = a[(k0 + u) * n + p];
                            xq[u] = a[(k0 + u) * n + q];
                            vp[u] = v[(k0 + u) * n + p];
                            vq[u] = v[(k0 + u) * n + q];
                        }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (k0 + u < n) {
                            a[(k0 + u) * n + p] = c * xp[u] - s * xq[u];
                            a[(k0 + u) * n + q] = s * xp[u] + c * xq[u];
                            v[(k0 + u) * n + p] = c * vp[u] - s * vq[u];
                            v[(k0 + u) * n + q] = s * vp[u] + c * vq[u];
                        }
                }
                for (int k0 = 0; k0 < n; k0 += 4) {
                    double xp[4], xq[4];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (k0 + u < n) {
                            xp[u] = a[p * n + k0 + u];
                            xq[u] = a[q * n + k0 + u];
                        }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (k0 + u < n) {
                            a[p * n + k0 + u] = c * xp[u] - s * xq[u];
                            a[q * n + k0 + u] = s * xp[u] + c * xq[u];
                        }
                }
            }
    }
    for (int i = 0; i < n; i++) w[i] = a[i * n + i];
}

template <int PT>
static __device__ __noinline__ void scen_populate(ScenDesign<PT>& sd, uint64_t base_seed, uint64_t sim_id) {
    constexpr int KT = ScenDesign<PT>::KT;
    const DevicePlan& pl = c_plan;
    const int k = pl.k;
    SeqRng sr;
    sr.init(pcg_mix64(base_seed, sim_id ^ 0x5C5C5C5C5C5C5C5CULL));  // scenarios.rs:19,50
    double a[KT * KT];  // column-major working matrix
    // ---- perturb_correlation (scenarios.rs:67-120)
    if (pl.corr_noise_sd == 0.0) {
        for (int j = 0; j < k; j++)
            for (int i = 0; i < k; i++) a[j * k + i] = pl.corr[j * MCP_MAX_K + i];
    } else {
        double noise[KT * KT];
        for (int j = 0; j < k; j++)
            for (int i = 0; i < k; i++) noise[j * k + i] = (double)sr.next_normal() * pl.corr_noise_sd;
        for (int j = 0; j < k; j++)
            for (int i = 0; i < k; i++) {
                const double sym = 0.5 * (noise[j * k + i] + noise[i * k + j]);
                const double v = pl.corr[j * MCP_MAX_K + i] + sym;
                a[j * k + i] = v < -0.8 ? -0.8 : (v > 0.8 ? 0.8 : v);
            }
        for (int i = 0; i < k; i++) a[i * k + i] = 1.0;
    }
    // ---- psd_repair_and_factor (correlation.rs:82-189)
    sd.ok = 1;
    if (k > 0) {
        double cp[KT * KT];
        for (int i = 0; i < k * k; i++) cp[i] = a[i];
        if (llt_colmajor(cp, k) == 0) {
            for (int i = 0; i < k * k; i++) a[i] = cp[i];
        } else {
            double sym[KT * KT], vec[KT * KT], lam[KT], d[KT];
            for (int j = 0; j < k; j++)
                for (int i = 0; i < k; i++) sym[i * k + j] = i >= j ? a[j * k + i] : a[i * k + j];
            jacobi_eigh(sym, k, lam, vec);
            for (int q = 0; q < k; q++)
                if (!(lam[q] > 1e-10)) lam[q] = 1e-10;
            for (int i = 0; i < k; i++)
                for (int j = 0; j < k; j++) {
                    double acc = 0.0;
                    for (int q = 0; q < k; q++) acc += vec[i * k + q] * lam[q] * vec[j * k + q];
                    cp[i * k + j] = acc;  // repaired, row-major
                }
            for (int i = 0; i < k; i++) d[i] = cp[i * k + i] > 0.0 ? sqrt(cp[i * k + i]) : 1.0;
            for (int j = 0; j < k; j++)
                for (int i = 0; i < k; i++) a[j * k + i] = cp[i * k + j] / (d[i] * d[j]);
            for (int i = 0; i < k; i++) a[i * k + i] = 1.0;
            if (llt_colmajor(a, k) != 0) sd.ok = 0;
        }
    }
    for (int j = 0; j < k; j++)
        for (int m = 0; m < k; m++) sd.L[j * KT + m] = m <= j ? a[m * k + j] : 0.0;
    // ---- perturb_var_types (scenarios.rs:127-156)
    for (int j = 0; j < k; j++) sd.vt[j] = pl.var_type[j];
    if (pl.dist_change_prob > 0.0 && pl.n_new_dists > 0) {
        for (int j = 0; j < k; j++) {
            const double u = (double)sr.next_uniform();
            if (sd.vt[j] == D_NORMAL && !pl.var_pinned[j] && u < pl.dist_change_prob) {
                int q = (int)((double)sr.next_uniform() * (double)pl.n_new_dists);
                if (q > pl.n_new_dists - 1) q = pl.n_new_dists - 1;
                sd.vt[j] = pl.new_dists[q];
            }
        }
    }
    // ---- pick_residual (scenarios.rs:163-186)
    sd.residual_dist = pl.residual_dist;
    if (pl.resid_change_prob > 0.0 && pl.n_resid_dists > 0) {
        const double u = (double)sr.next_uniform();
        if (u < pl.resid_change_prob && pl.residual_dist == R_NORMAL && !pl.residual_pinned) {
            int q = (int)((double)sr.next_uniform() * (double)pl.n_resid_dists);
            if (q > pl.n_resid_dists - 1) q = pl.n_resid_dists - 1;
            sd.residual_dist = pl.resid_dists[q];
        }
    }
    // ---- effective tau² of the study (data_gen.rs:1065-1075): base + N(0, icc_noise_sd²) drawn from the scenario stream AFTER
    // the design draws above, clamped at 0
    {
        double tsq = pl.has_cluster ? pl.tau_squared : 0.0;
        if (pl.has_lme && pl.icc_noise_sd > 0.0) {
            const double jitter = (double)sr.next_normal() * pl.icc_noise_sd;
            tsq = tsq + jitter;
            tsq = tsq > 0.0 ? tsq : 0.0;
        }
        sd.tau = sqrt(tsq > 0.0 ? tsq : 0.0);
        for (int g = 0; g < 2; g++) {
            double te = g < pl.n_extra ? pl.extra_tau_squared[g] : 0.0;
            if (g < pl.n_extra && pl.has_lme && pl.icc_noise_sd > 0.0) {
                te = te + (double)sr.next_normal() * pl.icc_noise_sd;
                te = te > 0.0 ? te : 0.0;
            }
            sd.tau_extra[g] = sqrt(te > 0.0 ? te : 0.0);
        }
    }
    // ---- per-study effect draw (data_gen.rs:709-751), domain-separated stream (scenarios.rs:28)
    const double het = pl.heterogeneity > 0.0 ? pl.heterogeneity : 0.0;
    sd.use_het = het > MCP_FNZ;
    if (sd.use_het) {
        SeqRng hr;
        hr.init(pcg_mix64(base_seed, sim_id ^ 0x4848484848484848ULL));
        const double int_sd = pl.outcome == O_CONTINUOUS ? 0.0 : het;
        for (int j = 0; j < PT; j++) {
            const double bj = pl.beta[j];
            const double sj = j == 0 ? int_sd : het * fabs(bj);
            const double raw = bj + (double)hr.next_normal() * sj;
            if (j == 0 || bj == 0.0) sd.beta_eff[j] = raw;
            else if (bj > 0.0) sd.beta_eff[j] = raw > 0.0 ? raw : 0.0;
            else sd.beta_eff[j] = raw < 0.0 ? raw : 0.0;
        }
    }
}

// Random effect of cluster c: u_c = f32(draw_residual(re_dist, re_df) · tau) (data_gen.rs:146-161, 1289-1320).  The draws sit
// at the head of the study's sequential stream, re_words words each, so cluster c's draw is counter-addressed at c · re_words.
// One draw_residual(re_dist, re_df) (data_gen.rs:1289-1320) whose first word is sequential-stream word w0.
static __device__ __noinline__ double re_draw(SimCtx& cx, uint32_t w0) {
    const DevicePlan& pl = c_plan;
    const int dist = pl.has_lme ? pl.re_dist : R_NORMAL;
    double d;
    if (dist == R_NORMAL) {
        d = (double)norm_inv_from_word(cx.seq.word_at(w0));
    } else if (dist == R_UNIFORM) {
        d = (2.0 * (double)word_to_unit(cx.seq.word_at(w0)) - 1.0) * MCP_SQRT3;
    } else {
        const double df = pl.re_df > 3.0 ? pl.re_df : 3.0;
        long long dfi = llround(df);
        if (dfi < 1) dfi = 1;
        const bool is_t = dist == R_HIGH_KURTOSIS;
        const double z = is_t ? (double)norm_inv_from_word(cx.seq.word_at(w0)) : 0.0;  // sample_t draws z first (distributions.rs:78-97)
        double chi2 = 0.0;
        for (long long i = 0; i < dfi; i++) {
            const double g = (double)norm_inv_from_word(cx.seq.word_at(w0 + (is_t ? 1u : 0u) + (uint32_t)i));
            chi2 += g * g;
        }
        if (is_t) {
            const double scale = 1.0 / sqrt(df / (df - 2.0));
            const double denom = sqrt(chi2 / df);
            d = (denom <= 0.0 ? z : z / denom) * scale;
        } else {
            const double centered = (chi2 - df) * (1.0 / sqrt(2.0 * df));
            d = dist == R_LEFT_SKEWED ? -centered : centered;
        }
    }
    return d;
}
static __device__ __noinline__ double cluster_effect(SimCtx& cx, uint32_t c) {
    return (double)(float)(re_draw(cx, c * (uint32_t)c_plan.re_words) * cx.tau);
}
// Level of row i in extra grouping g (ClusterSpec::extra_level_of_row, generation.rs:370-430): one lexicographic stride chain,
// primary fastest, extras in declaration order; nested levels are parent · n_per_parent + within.
__device__ __forceinline__ int extra_level_of_row(const DevicePlan& pl, int g, int i) {
    const int bl = pl.extra_count[g], sv = pl.sizing_value > 0 ? pl.sizing_value : 1;
    if (!pl.sizing_fixed_size) {
        const int stride = g == 0 ? sv : sv * pl.extra_count[0];
        const int within = (i / stride) % bl;
        return pl.extra_nested[g] ? (i % sv) * bl + within : within;
    }
    return (i / sv) * bl + (i % sv) % bl;
}
// Sum of the extra groupings' intercept effects of one row (data_gen.rs:214-236, 818-835): grouping g's draws sit behind the
// primary's (intercepts, then slopes) and behind the earlier groupings' blocks on the sequential stream.
static __device__ __noinline__ double extra_effects(SimCtx& cx, int row, double u_re) {
    const DevicePlan& pl = c_plan;
    uint32_t base = (uint32_t)pl.n_clusters * (uint32_t)(1 + pl.n_slopes);
    for (int g = 0; g < pl.n_extra; g++) {
        const uint32_t lvl = (uint32_t)extra_level_of_row(pl, g, row);
        u_re += (double)(float)(re_draw(cx, (base + lvl) * (uint32_t)pl.re_words) * cx.tau_extra[g]);
        base += (uint32_t)pl.extra_levels[g];
    }
    return u_re;
}
// Random-slope effects of cluster c (data_gen.rs:163-213): z_1 (, z_2) are drawn behind ALL the intercept draws, cluster by
// cluster; z_0 is recovered from the f32 intercept effect; u_k = f32(sum_{j <= k} L[k][j] z_j) with L = chol(diag(tau) R diag(tau)).
static __device__ __noinline__ void cluster_slope_effects(SimCtx& cx, uint32_t c, double u0, double (&uk)[2]) {
    const DevicePlan& pl = c_plan;
    const uint32_t nsl = (uint32_t)pl.n_slopes, w1 = ((uint32_t)pl.n_clusters + c * nsl) * (uint32_t)pl.re_words;
    const double z0 = cx.tau > 0.0 ? u0 / cx.tau : 0.0;
    const double z1 = re_draw(cx, w1);
    double u = 0.0;
    u += cx.sl[0] * z0;
    u += cx.sl[1] * z1;
    uk[0] = (double)(float)u;
    uk[1] = 0.0;
    if (nsl == 2u) {
        const double z2 = re_draw(cx, w1 + (uint32_t)pl.re_words);
        u = 0.0;
        u += cx.sl[2] * z0;
        u += cx.sl[3] * z1;
        u += cx.sl[4] * z2;
        uk[1] = (double)(float)u;
    }
}
// chol_lower(D) (data_gen.rs:55-75), D = diag(tau) R diag(tau) with this study's tau0 (the ICC jitter makes it per study)
static __device__ __noinline__ void slope_factor(SimCtx& cx) {
    const DevicePlan& pl = c_plan;
    const int q = 1 + pl.n_slopes;
    double tv[3], R[3][3], D[3][3], L[3][3];
    tv[0] = cx.tau;
    for (int k = 0; k < pl.n_slopes; k++) tv[1 + k] = sqrt(pl.slope_variance[k] > 0.0 ? pl.slope_variance[k] : 0.0);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            R[i][j] = i == j ? 1.0 : 0.0;
            L[i][j] = 0.0;
        }
    for (int k = 0; k < pl.n_slopes; k++) R[0][k + 1] = R[k + 1][0] = pl.slope_corr[k];
    if (pl.n_slopes == 2) R[1][2] = R[2][1] = pl.slope_corr12;
    for (int i = 0; i < q; i++)
        for (int j = 0; j < q; j++) D[i][j] = tv[i] * R[i][j] * tv[j];
    for (int j = 0; j < q; j++) {
        double diag = D[j][j];
        for (int k = 0; k < j; k++) diag -= L[j][k] * L[j][k];
        const double ljj = sqrt(diag > 0.0 ? diag : 0.0);
        L[j][j] = ljj;
        for (int i = j + 1; i < q; i++)
            if (ljj > 0.0) {
                double sacc = D[i][j];
                for (int k = 0; k < j; k++) sacc -= L[i][k] * L[j][k];
                L[i][j] = sacc / ljj;
            }
    }
    cx.sl[0] = L[1][0];
    cx.sl[1] = L[1][1];
    cx.sl[2] = L[2][0];
    cx.sl[3] = L[2][1];
    cx.sl[4] = L[2][2];
}

// Cluster-level predictors (data_gen.rs:605-640): on every non-representative row the marked columns take the value of the
// cluster's representative row (FixedClusters: row i mod n_clusters; FixedSize: the cluster's first row).  The representative
// values of a study are parked in a per-wave table in HBM; rows are generated four at a time — and by four lanes at once in the
// quad kernels — so the stores of a block's representative rows (between_store) and the loads of its other rows (between_load)
// are separate passes with a quad barrier between them.
template <typename SlabT>
static __device__ __noinline__ void between_store(const SlabT slab, int b, uint32_t sim_local) {
    const DevicePlan& pl = c_plan;
    const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
    for (int r = 0; r < 4; r++) {
        const int row = 4 * b + r;
        if (row >= pl.max_n) break;
        const int rep = pl.sizing_fixed_size ? row - row % cs : row % cs;
        if (row != rep) continue;
        const int c = pl.sizing_fixed_size ? row / cs : row;
        for (int q = 0; q < pl.n_between; q++)
            pl.between_tab[((size_t)c * pl.n_between + q) * pl.between_stride + sim_local] = slab[4 * (pl.between_cols[q] - 1) + r];
    }
}
template <typename SlabT>
static __device__ __noinline__ void between_load(const SlabT slab, int b, uint32_t sim_local) {
    const DevicePlan& pl = c_plan;
    const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
    for (int r = 0; r < 4; r++) {
        const int row = 4 * b + r;
        if (row >= pl.max_n) break;
        const int rep = pl.sizing_fixed_size ? row - row % cs : row % cs;
        if (row == rep) continue;
        const int c = pl.sizing_fixed_size ? row / cs : rep;
        for (int q = 0; q < pl.n_between; q++)
            slab[4 * (pl.between_cols[q] - 1) + r] = __ldcg(&pl.between_tab[((size_t)c * pl.n_between + q) * pl.between_stride + sim_local]);
    }
}

// Strict bootstrap (data_gen.rs:553-604) for rows 4b..4b+3: ONE source row per design row — picked by the sequential-stream
// word after the row's factor draws — is copied into every mapped column (the values generated before are discarded) and
// the mapped factors are re-expanded to reference-coded dummies.  Out of line and self-contained (its own view of the
// sequential stream), so that the synthetic path keeps its register footprint.
template <typename SlabT>
static __device__ __noinline__ void bootstrap_block(const SlabT slab, int b, uint32_t k0, uint32_t k1, uint32_t seq_row_base) {
    const DevicePlan& pl = c_plan;
    SeqRng seq;
    seq.k0 = k0;
    seq.k1 = k1;
    seq.pos = 0;
    seq.buf_block = 0xffffffffu;
    for (int r = 0; r < 4; r++) {
        const int row = 4 * b + r;
        if (row >= pl.max_n) break;
        const uint32_t word = seq.word_at(seq_row_base + (uint32_t)row * (uint32_t)pl.seq_per_row + (uint32_t)pl.n_sampled_factors);
        const double pick = (double)word_to_unit(word) * (double)pl.ud_rows;
        int src = (int)pick;  // pick ∈ [0, ud_rows): truncation = Rust's `as usize`
        if (src > pl.ud_rows - 1) src = pl.ud_rows - 1;
        const double* __restrict__ fr = pl.upload_data + (size_t)src * pl.ud_cols;
        for (int j = 0; j < pl.k; j++) {
            const int fc = pl.frame_map[j];
            if (fc < 0) continue;
            double v = fr[fc];
            // the ORIGINAL contract type decides, not the scenario-perturbed one (data_gen.rs:570-573)
            if (pl.var_type[j] == D_UPLOADED_BINARY) v = v + pl.var_param_bs[j] >= 0.5 ? 1.0 : 0.0;
            slab[4 * j + r] = (float)v;
        }
        for (int f = 0; f < pl.n_factors; f++) {
            const int fc = pl.frame_map[MCP_MAX_K + f];
            if (fc < 0) continue;
            const double lv = round(fr[fc]);
            const int level = !(lv > 0.0) ? 0 : (lv > 1e9 ? 1000000000 : (int)lv);  // saturating `as usize`, NaN → 0
            const int c0 = pl.f_col[f] - 1, nd = pl.f_levels[f] - 1;
            for (int d = 0; d < nd; d++) slab[4 * (c0 + d) + r] = (level == d + 1) ? 1.0f : 0.0f;
        }
    }
}

// Generate rows 4b..4b+3 of the f32 plane into the thread's slab:
//   slab[4*(j-1) + r] = X[4b+r][j] for design columns j = 1..P-1,  slab[4*(P-1) + r] = residual draw.
// marginal transforms of one generated block on the mixed f32 values, two rows of a column per call (same kind: their chains
// interleave)
template <int PT, bool SCEN, typename SlabT>
__device__ __forceinline__ void gen_marginals(const SlabT slab, const ScenDesign<PT>* sd) {
    const DevicePlan& pl = c_plan;
#pragma unroll 1
    for (int j = 0; j < pl.k; j++) {
        const int vt = SCEN ? sd->vt[j] : pl.var_type[j];
        if (vt == D_NORMAL) continue;
        const float2 lo = apply_marginal2(slab[4 * j + 0], slab[4 * j + 1], vt | (j << 8), pl.var_param[j], pl.t3_table);
        const float2 hi = apply_marginal2(slab[4 * j + 2], slab[4 * j + 3], vt | (j << 8), pl.var_param[j], pl.t3_table);
        slab[4 * j + 0] = lo.x;
        slab[4 * j + 1] = lo.y;
        slab[4 * j + 2] = hi.x;
        slab[4 * j + 3] = hi.y;
    }
}

// gen_block_pre: everything up to and including the factor / bootstrap block (+ the stores of cluster-level predictors);
// gen_block_post: the loads of cluster-level predictors and the interaction columns.  gen_block = both, for one thread per study.
// PHASE 0 = all of it; 1 = up to and including the Cholesky mix (the caller runs the marginal transforms itself: the lane-quad
// kernels' warp-cooperative pass, kernels_q4.cuh); 2 = the rest (factor block, bootstrap, cluster-level stores).
template <int PT, bool SCEN, int PHASE = 0, typename SlabT>
__device__ __forceinline__ void gen_block_pre(int b, const SlabT slab, SimCtx& cx, const ScenDesign<PT>* sd) {
    constexpr int KT = ScenDesign<PT>::KT;
    const DevicePlan& pl = c_plan;
    const int k = pl.k;
    const uint32_t k0 = cx.k0, k1 = cx.k1;
    const int residual_dist = SCEN ? sd->residual_dist : pl.residual_dist;
    if (PHASE != 2) {
    // ---- A. normals for the continuous block (CLASS_XNORM=1, column j) and residual slot 0 (CLASS_RESID=2)
    // One "column task" = one Philox block → 4 draws → central polynomial → slab column.  Tasks are independent;
    // the loop is unrolled by 3 so that three Philox chains and six packed Horner chains are in flight per thread
    // (the kernel runs at 2 warps per scheduler, so latency must be hidden by ILP inside the thread).
    unsigned long long tmask = 0ull;
    const bool uniform_slot0 = (pl.outcome == O_CONTINUOUS && residual_dist == R_UNIFORM) ||
                               (pl.outcome == O_BINARY && !pl.link_probit);
    const int n_tasks = uniform_slot0 ? k : k + 1;
#pragma unroll kUnrollTasks
    for (int t = 0; t < n_tasks; t++) {
        const bool is_x = t < k;
        const U4 w = philox4x32_10((uint32_t)b, is_x ? (uint32_t)t : 0u, is_x ? 1u : 2u, 0u, k0, k1);
        const int c = is_x ? t : PT - 1;
        float4 z;
        float v0, v1, v2, v3;  // packed f32x2 polynomial: FMUL2 + FFMA2(acc, 1, c), bit-identical to mul + add (device_math.cuh)
        ni_central_pair_packed(w.x, w.y, z.x, z.y, v0, v1);
        ni_central_pair_packed(w.z, w.w, z.z, z.w, v2, v3);
        if (t < 15 || !is_x) {
            unsigned tb = 0u;
            if (fabsf(v0) > MCP_NI_BP) { z.x = v0; tb |= 1u; }
            if (fabsf(v1) > MCP_NI_BP) { z.y = v1; tb |= 2u; }
            if (fabsf(v2) > MCP_NI_BP) { z.z = v2; tb |= 4u; }
            if (fabsf(v3) > MCP_NI_BP) { z.w = v3; tb |= 8u; }
            tmask |= (unsigned long long)tb << (is_x ? 4 * t : 60);
        } else {
            if (fabsf(v0) > MCP_NI_BP) z.x = ni_tail(v0);
            if (fabsf(v1) > MCP_NI_BP) z.y = ni_tail(v1);
            if (fabsf(v2) > MCP_NI_BP) z.z = ni_tail(v2);
            if (fabsf(v3) > MCP_NI_BP) z.w = ni_tail(v3);
        }
        slab[4 * c + 0] = z.x;
        slab[4 * c + 1] = z.y;
        slab[4 * c + 2] = z.z;
        slab[4 * c + 3] = z.w;
    }
    if (uniform_slot0) {
        const U4 w = philox4x32_10((uint32_t)b, 0u, 2u, 0u, k0, k1);
        float4 z;
        if (pl.outcome == O_CONTINUOUS) {  // Uniform residual: (2u-1)*sqrt(3) (data_gen.rs:485-501)
            z.x = (float)((2.0 * (double)word_to_unit(w.x) - 1.0) * MCP_SQRT3);
            z.y = (float)((2.0 * (double)word_to_unit(w.y) - 1.0) * MCP_SQRT3);
            z.z = (float)((2.0 * (double)word_to_unit(w.z) - 1.0) * MCP_SQRT3);
            z.w = (float)((2.0 * (double)word_to_unit(w.w) - 1.0) * MCP_SQRT3);
        } else {  // logit: the raw uniform is compared against sigmoid(eta) (data_gen.rs:390-398)
            z.x = word_to_unit(w.x);
            z.y = word_to_unit(w.y);
            z.z = word_to_unit(w.z);
            z.w = word_to_unit(w.w);
        }
        slab[4 * (PT - 1) + 0] = z.x;
        slab[4 * (PT - 1) + 1] = z.y;
        slab[4 * (PT - 1) + 2] = z.z;
        slab[4 * (PT - 1) + 3] = z.w;
    }
    // ---- B. deferred tail branch of the inverse-normal kernel
    while (tmask) {
        const int bit = __ffsll((long long)tmask) - 1;
        tmask &= tmask - 1ull;
        const int c = bit >> 2;
        const int idx = (c == 15 ? (PT - 1) : c) * 4 + (bit & 3);
        slab[idx] = ni_tail(slab[idx]);
    }
    // ---- B'. t / chi-square residuals need extra CLASS_RESID slots
    if (pl.outcome == O_CONTINUOUS && residual_dist != R_NORMAL && residual_dist != R_UNIFORM) {
        const float4 e4 = transform_residual4(residual_dist, pl.residual_df,
                                              make_float4(slab[4 * (PT - 1) + 0], slab[4 * (PT - 1) + 1], slab[4 * (PT - 1) + 2], slab[4 * (PT - 1) + 3]),
                                              (uint32_t)b, k0, k1);
        slab[4 * (PT - 1) + 0] = e4.x;
        slab[4 * (PT - 1) + 1] = e4.y;
        slab[4 * (PT - 1) + 2] = e4.z;
        slab[4 * (PT - 1) + 3] = e4.w;
    }
    // ---- C. Cholesky mix + marginal transforms (data_gen.rs:334-363); skipped when L = I and all Normal
    if (SCEN || !(pl.l_identity && pl.all_normal)) {
        for (int r = 0; r < 4; r++) {
            double zr[PT];  // zr[j] = z of continuous column j (j < k)
#pragma unroll
            for (int j = 0; j < PT - 1; j++) zr[j] = j < k ? (double)slab[4 * j + r] : 0.0;
#pragma unroll
            for (int j = PT - 2; j >= 0; j--) {
                if (j < k) {
                    double a = 0.0;
#pragma unroll
                    for (int m = 0; m <= j; m++) a = a + (SCEN ? sd->L[j * KT + m] : pl.L[j * MCP_MAX_K + m]) * zr[m];
                    slab[4 * j + r] = (float)a;
                }
            }
        }
        if (PHASE == 0 && (SCEN || !pl.all_normal)) gen_marginals<PT, SCEN>(slab, sd);
    }
    }  // PHASE != 2
    if (PHASE == 1) return;
    // ---- D. factor dummies (data_gen.rs:509-551) and interactions (643-649), in the f32 plane
    if (pl.n_factors > 0) {
        for (int r = 0; r < 4; r++) {
            const int row = 4 * b + r;
            if (row >= pl.max_n) break;
            int sidx = 0;
            for (int f = 0; f < pl.n_factors; f++) {
                const int nl = pl.f_levels[f];
                int lvl;
                if (!pl.f_sampled[f]) {
                    lvl = pl.levels[(size_t)f * pl.max_n + row];
                } else {  // next_categorical on the sequential stream (rng.rs:112-127), counter-addressed
                    const uint32_t widx = cx.seq_row_base + (uint32_t)row * (uint32_t)pl.seq_per_row + (uint32_t)sidx;
                    const double u = (double)word_to_unit(cx.seq.word_at(widx));
                    const double* pr = pl.f_props + pl.f_off[f];
                    double ca = 0.0;
                    lvl = nl - 1;
                    for (int q = 0; q < nl; q++) {
                        ca += pr[q];
                        if (u < ca) { lvl = q; break; }
                    }
                    sidx++;
                }
                const int c0 = pl.f_col[f] - 1;
                for (int d = 0; d < nl - 1; d++) slab[4 * (c0 + d) + r] = (lvl == d + 1) ? 1.0f : 0.0f;
            }
        }
    }
    // strict bootstrap: after the factor block, before the interactions (data_gen.rs:553-604)
    if (pl.bootstrap) bootstrap_block(slab, b, cx.seq.k0, cx.seq.k1, cx.seq_row_base);
    if (pl.n_between > 0) between_store(slab, b, cx.sim_local);
}

template <int PT, bool SCEN, typename SlabT>
__device__ __forceinline__ void gen_block_post(int b, const SlabT slab, SimCtx& cx) {
    const DevicePlan& pl = c_plan;
    const int k = pl.k;
    if (pl.n_between > 0) between_load(slab, b, cx.sim_local);
    if (PT > MCP_MAX_TEMPLATE_P) {  // padded width: the columns behind the real design are zeros (plan.h, mcp_template_p)
        for (int j = pl.P; j < PT; j++)
            for (int r = 0; r < 4; r++) slab[4 * (j - 1) + r] = 0.0f;
    }
    if (pl.n_inter > 0) {
        for (int r = 0; r < 4; r++) {
            if (4 * b + r >= pl.max_n) break;
            const int ib = k + pl.n_fd;
            for (int j = 0; j < pl.n_inter; j++) {
                double prod = 1.0;
                for (int q = pl.inter_ptr[j]; q < pl.inter_ptr[j + 1]; q++) prod *= (double)slab[4 * (pl.inter_cols[q] - 1) + r];
                slab[4 * (ib + j) + r] = (float)prod;
            }
        }
    }
}

template <int PT, bool SCEN, typename SlabT>
__device__ __forceinline__ void gen_block(int b, const SlabT slab, SimCtx& cx, const ScenDesign<PT>* sd) {
    gen_block_pre<PT, SCEN, 0>(b, slab, cx, sd);
    gen_block_post<PT, SCEN>(b, slab, cx);
}

// Per-study generation context: Philox key, sequential stream, tau (the scenario design's when there is one), and where
// the per-row words of the sequential stream start (behind the cluster draws).
template <int PT>
__device__ __forceinline__ void init_sim_ctx(SimCtx& cx, uint64_t key, uint32_t sim_local, const ScenDesign<PT>* sd) {
    const DevicePlan& pl = c_plan;
    cx.k0 = (uint32_t)key;
    cx.k1 = (uint32_t)(key >> 32);
    cx.seq.init(key);
    cx.tau = !pl.has_cluster ? 0.0 : (sd ? sd->tau : sqrt(pl.tau_squared > 0.0 ? pl.tau_squared : 0.0));
    cx.seq_row_base = pl.has_cluster ? ((uint32_t)pl.n_clusters * (uint32_t)(1 + pl.n_slopes) + (uint32_t)(pl.extra_levels[0] + pl.extra_levels[1])) *
                                           (uint32_t)pl.re_words
                                     : 0u;
    for (int g = 0; g < 2; g++)
        cx.tau_extra[g] = g >= pl.n_extra ? 0.0 : (sd ? sd->tau_extra[g] : sqrt(pl.extra_tau_squared[g] > 0.0 ? pl.extra_tau_squared[g] : 0.0));
    cx.sim_local = sim_local;
    if (pl.n_slopes > 0) slope_factor(cx);
}

// MCP_ROW_LP_ALL: the linear predictor visits every column, as the reference's loop does (data_gen.rs:781-795), instead of testing
// a non-zero-effect mask per column (which ptxas turns into DMUL + DADD + two selects): config 3 realistic +2.4 %, doomer +0.5 %
#ifndef MCP_ROW_LP_ALL
#define MCP_ROW_LP_ALL 1
#endif
// Widen row r of the slab to f64 registers and assemble y (data_gen.rs:781-922).  Returns y as f64 of the f32 value.
template <int PT, bool SCEN, typename SlabT>
__device__ __forceinline__ double row_xy(int b, int r, const SlabT slab, SimCtx& cx, double (&xr)[PT], const ScenDesign<PT>* sd) {
    const DevicePlan& pl = c_plan;
    const int row = 4 * b + r;
    xr[0] = 1.0;
#pragma unroll
    for (int j = 1; j < PT; j++) xr[j] = (double)slab[4 * (j - 1) + r];
    double lp = 0.0;
#if MCP_ROW_LP_ALL
    // every column, as the reference's loop does (data_gen.rs:781-795): two FP64 instructions per column on registers
#pragma unroll
    for (int j = 0; j < PT; j++) lp = lp + xr[j] * pl.beta[j];
#else
    const uint32_t bm = pl.beta_nonzero_mask;  // β_j == 0 terms add ±0 to a +0-seeded sum: skipping them is bit-exact
#pragma unroll
    for (int j = 0; j < PT; j++)
        if (bm & (1u << j)) lp = lp + xr[j] * pl.beta[j];
#endif
    float ef = slab[4 * (PT - 1) + r];
    double u_re = 0.0;
    if (pl.has_cluster) {  // u_c = f32(normal(word c) * tau), drawn first on the sequential stream (data_gen.rs:146-161)
        const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
        const uint32_t cid = pl.sizing_fixed_size ? (uint32_t)(row / cs) : (uint32_t)(row % cs);
        // rows of a trailing partial cluster (FixedSize, max_n not a multiple of cluster_size) have no drawn effect
        if (cx.utab) u_re = cx.utab[cid];
        else if (cid < (uint32_t)pl.n_clusters) u_re = cluster_effect(cx, cid);
        const double u0 = u_re;
        if (pl.n_extra > 0) u_re = extra_effects(cx, row, u_re);  // extras first (data_gen.rs:818-835), then the primary slopes
        if (pl.n_slopes > 0 && cid < (uint32_t)pl.n_clusters) {  // data_gen.rs:836-842: u_re += u_k · x[slope column k]
            double uk[2];
            cluster_slope_effects(cx, cid, u0, uk);
            for (int k = 0; k < pl.n_slopes; k++) u_re += uk[k] * (double)slab[4 * (pl.slope_col[k] - 1) + r];
        }
    }
    if (SCEN) {
        // per-study jittered effects drive y; the clean predictor drives heteroskedasticity (data_gen.rs:781-809)
        const double lp_clean = lp;
        if (sd->use_het) {
            double le = 0.0;
#pragma unroll
            for (int j = 0; j < PT; j++) le = le + xr[j] * sd->beta_eff[j];
            lp = le;
        }
        if (pl.apply_hsk) {  // Var(e_i) = exp(gamma z_i) / exp(gamma^2/2) (data_gen.rs:798-809, 876-888)
            const double drv = pl.hsk_driver < 0 ? lp_clean : (double)slab[4 * (pl.hsk_driver - 1) + r];
            const double zz = (drv - pl.hsk_center) * pl.hsk_inv_std;
            const double mult = own_exp(pl.hsk_gamma * zz) * pl.hsk_inv_norm;
            ef = (float)((double)ef * sqrt(mult));
            return (double)(float)((lp + u_re) + (double)ef);
        }
    }
    if (pl.outcome == O_CONTINUOUS) return (double)(float)(lp + (double)ef + u_re);
    const double eta = lp + u_re;
    if (pl.outcome == O_BINARY) {
        const double thr = pl.link_probit ? eta : sigmoid_stable(eta);
        return (double)ef < thr ? 1.0 : 0.0;
    }
    return (double)(float)poisson_quantile((double)ef, own_exp(eta));
}

// ------------------------------------------------------------------------------------------
// One OLS fit at a grid point, from sufficient statistics handed over THROUGH MEMORY (st = X'X packed lower
// triangle [NACC], X'y [PT], element i at st[i * stride]: a local array, or the [entry][sim] buffer in HBM).  Deliberately not inlined: the solver needs ~2x the registers of the streaming loops,
// and inlining it makes ptxas spill the live accumulators inside the hot row loop (measured: +70 % instructions,
// 1.6 GB of local-memory write-back per 0.5 M studies).  As a call, the spill happens once per fit at the call site.
template <int PT>
static __device__ __noinline__ void ols_fit_and_emit(const double* __restrict__ st, double yty, double sum_y, uint32_t n_rows,
                                                     const SimOut& so, int g, const uint32_t* lvl_counts, uint8_t* excl,
                                                     bool track_excl, size_t stride = 1) {
    constexpr int NACC = PT * (PT + 1) / 2;
    const DevicePlan& pl = c_plan;
    int variant = pl.base_variant;
    if (track_excl) {  // batch.rs:2485-2495
        int maskbits = 0;
        for (int f = 0; f < pl.n_factors; f++) {
            uint32_t mn = 0xffffffffu;
            for (int q = 0; q < pl.f_levels[f]; q++) mn = min(mn, lvl_counts[pl.f_off[f] + q]);
            excl[f] = mn < (uint32_t)pl.min_level_count;
            maskbits |= (int)excl[f] << f;
        }
        variant = maskbits;
    }
    const McpVariant& vr = pl.variants[variant];
    double all_t[MCP_MAX_T];
    bool full = false;
    if constexpr (PT <= MCP_MAX_TEMPLATE_P) full = vr.p_red == PT;  // padded widths always fit the kept (real) columns
    if constexpr (PT <= MCP_MAX_TEMPLATE_P) if (full) {
        double acc[NACC], xty[PT];
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = st[i * stride];
#pragma unroll
        for (int i = 0; i < PT; i++) xty[i] = st[(NACC + i) * stride];
        RegSolver<PT> sv;
        sv.factor_and_solve(acc, xty, yty, sum_y, n_rows);
        for (int t = 0; t < pl.n_marg; t++) all_t[t] = sv.t_sq(pl.target[t], -1);
        for (int t = 0; t < pl.n_contrasts; t++) all_t[pl.n_marg + t] = sv.t_sq(pl.contrast[2 * t], pl.contrast[2 * t + 1]);
        double ph_t[MCP_MAX_PH];
        for (int c = 0; c < pl.n_ph; c++) ph_t[c] = sv.t_sq(pl.ph_contrast[2 * c], pl.ph_contrast[2 * c + 1]);
        emit_fit(so, g, variant, sv.out, PT, all_t, excl, -1, 0, 0, pl.n_ph > 0 ? ph_t : nullptr);
    }
    if (!full) {
        // reduced model: gather kept rows/cols of the sufficient statistics (batch.rs:970-1013)
        double Ar[NACC], br[PT];
        const int pr = vr.p_red;
        for (int ri = 0; ri < pr; ri++) {
            br[ri] = st[(NACC + vr.kept[ri]) * stride];
            for (int rj = 0; rj <= ri; rj++) Ar[tri(ri, rj)] = st[tri(vr.kept[ri], vr.kept[rj]) * stride];
        }
        MemSolver<PT> ms;
        ms.factor_and_solve(Ar, br, pr, yty, sum_y, n_rows);
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int t = 0; t < pl.n_marg; t++) {
            const int c = vr.remap[pl.target[t]];
            all_t[t] = c >= 0 ? ms.t_sq(c, -1) : nan;
        }
        for (int t = 0; t < pl.n_contrasts; t++) {
            const int pc = vr.remap[pl.contrast[2 * t]], nc = vr.remap[pl.contrast[2 * t + 1]];
            all_t[pl.n_marg + t] = (pc >= 0 && nc >= 0) ? ms.t_sq(pc, nc) : nan;
        }
        // posthoc under exclusion (batch.rs:1163-1203): a contrast touching a dropped column becomes the zero contrast,
        // whose se² fails the > 1e-30 guard — never significant
        double ph_t[MCP_MAX_PH];
        for (int c = 0; c < pl.n_ph; c++) {
            const int pos = pl.ph_contrast[2 * c], neg = pl.ph_contrast[2 * c + 1];
            const int pc = vr.remap[pos], nc = neg >= 0 ? vr.remap[neg] : -1;
            ph_t[c] = (pc >= 0 && (neg < 0 || nc >= 0)) ? ms.t_sq(pc, nc) : nan;
        }
        emit_fit(so, g, variant, ms.out, pr, all_t, excl, -1, 0, 0, pl.n_ph > 0 ? ph_t : nullptr);
    }
}

// ------------------------------------------------------------------------------------------
// The OLS power kernel.
//   PT  : total design columns P (compile-time → register-resident sufficient statistics)
// Dynamic shared memory: [MCP_BLOCK] per-thread slabs of 4*SP floats (SP = P|1 keeps the per-thread
// LDS.128/STS.128 conflict-free), then the optional block-level counters.
template <int PT, bool SCEN>
__global__ void __launch_bounds__(MCP_BLOCK, MCP_MIN_BLOCKS) ols_power_kernel(uint64_t base_seed, uint64_t sim_begin, uint32_t n_sims) {
    constexpr int NACC = PT * (PT + 1) / 2;
    constexpr int SP = PT | 1;
    const DevicePlan& pl = c_plan;
    if (mcp_cancelled()) return;
    extern __shared__ __align__(16) float smem[];
    const Slab<MCP_BLOCK> slab{smem + threadIdx.x};
    unsigned int* cnt_s = nullptr;
    if (pl.smem_counters) {
        cnt_s = reinterpret_cast<unsigned int*>(smem + (size_t)MCP_BLOCK * 4 * SP);
        for (int i = threadIdx.x; i < pl.n_ss * pl.counter_width; i += MCP_BLOCK) cnt_s[i] = 0u;
        __syncthreads();
    }
    const uint32_t sim = blockIdx.x * MCP_BLOCK + threadIdx.x;
    if (sim < n_sims) {
        const uint64_t key = pcg_mix64(base_seed, sim_begin + sim);
        const int max_n = pl.max_n;
        SimOut so{cnt_s, (uint64_t)sim};
        ScenDesign<PT> sdv;
        ScenDesign<PT>* sd = nullptr;
        if (SCEN) {
            scen_populate<PT>(sdv, base_seed, sim_begin + sim);
            sd = &sdv;
            if (!sdv.ok) atomicExch(pl.error_flag, 2);  // EngineError::CorrelationNotPSD (data_gen.rs:1049)
        }
        SimCtx cx;
        init_sim_ctx<PT>(cx, key, sim, sd);

        double acc[NACC], xty[PT];
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = 0.0;
#pragma unroll
        for (int i = 0; i < PT; i++) xty[i] = 0.0;
        double yty = 0.0, sum_y = 0.0;
        uint32_t n_rows = 0;

        uint32_t lvl_counts[MCP_MAX_LEVELS];
        uint8_t excl[MCP_MAX_FACTORS];
        const bool track_excl = pl.n_factors > 0 && pl.min_level_count > 0;
        if (track_excl)
            for (int i = 0; i < MCP_MAX_LEVELS; i++) lvl_counts[i] = 0;
        for (int f = 0; f < MCP_MAX_FACTORS; f++) excl[f] = 0;

        int g = 0;
        uint32_t n_next = pl.ss[0];
        const int n_blocks = (max_n + 3) >> 2;

        for (int b = 0; b < n_blocks; b++) {
            gen_block<PT, SCEN>(b, slab, cx, sd);
            // ---- E. row pass: widen, linear predictor, y, sufficient statistics, fits at grid points
#pragma unroll 1
            for (int r = 0; r < 4; r++) {
                const int row = 4 * b + r;
                if (row >= max_n) break;
                double xr[PT];
                const double yd = row_xy<PT, SCEN>(b, r, slab, cx, xr, sd);
#pragma unroll
                for (int a = 0; a < PT; a++) {
#pragma unroll
                    for (int c = 0; c <= a; c++) acc[tri(a, c)] = fma(xr[a], xr[c], acc[tri(a, c)]);
                    xty[a] = fma(xr[a], yd, xty[a]);
                }
                yty = fma(yd, yd, yty);
                sum_y += yd;
                n_rows++;
                if (track_excl) {  // batch.rs:2451-2496
                    for (int f = 0; f < pl.n_factors; f++) {
                        const int c0 = pl.f_col[f] - 1, nl = pl.f_levels[f];
                        int lvl = 0;
                        for (int d = 0; d < nl - 1; d++)
                            if (slab[4 * (c0 + d) + r] == 1.0f) { lvl = d + 1; break; }
                        lvl_counts[pl.f_off[f] + lvl]++;
                    }
                }
                if ((uint32_t)(row + 1) == n_next) {
                    // ---- fit at grid point g (through memory, see ols_fit_and_emit)
                    double st[NACC + PT];
#pragma unroll
                    for (int i = 0; i < NACC; i++) st[i] = acc[i];
#pragma unroll
                    for (int i = 0; i < PT; i++) st[NACC + i] = xty[i];
                    ols_fit_and_emit<PT>(st, yty, sum_y, n_rows, so, g, lvl_counts, excl, track_excl);
                    g++;
                    n_next = g < pl.n_ss ? pl.ss[g] : 0xffffffffu;
                }
            }
        }
    }
    mcp_report_warp((blockIdx.x * MCP_BLOCK + threadIdx.x) & ~31u, 32u, n_sims);
    if (cnt_s) {
        __syncthreads();
        for (int i = threadIdx.x; i < pl.n_ss * pl.counter_width; i += MCP_BLOCK) {
            const unsigned v = cnt_s[i];
            if (v) atomicAdd(&pl.counters[i], (unsigned long long)v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K0: dump the generated design of one simulation (introspect::capture_sim0_data, introspect.rs:104-155)
// through exactly the same generation functions as the power kernel.  One thread; parity/debug only.
template <int PT, bool SCEN>
__global__ void dgp_dump_kernel(uint64_t base_seed, uint64_t sim_id, float* __restrict__ x_out /* col-major max_n × P */,
                                float* __restrict__ y_out) {
    constexpr int SP = PT | 1;
    const DevicePlan& pl = c_plan;
    __shared__ __align__(16) float slab_mem[4 * SP];
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Slab<1> slab{slab_mem};
    const uint64_t key = pcg_mix64(base_seed, sim_id);
    const int max_n = pl.max_n;
    ScenDesign<PT> sdv;
    ScenDesign<PT>* sd = nullptr;
    if (SCEN) {
        scen_populate<PT>(sdv, base_seed, sim_id);
        sd = &sdv;
        if (!sdv.ok) atomicExch(pl.error_flag, 2);
    }
    SimCtx cx;
    init_sim_ctx<PT>(cx, key, 0u, sd);
    for (int b = 0; b < (max_n + 3) / 4; b++) {
        gen_block<PT, SCEN>(b, slab, cx, sd);
        for (int r = 0; r < 4; r++) {
            const int row = 4 * b + r;
            if (row >= max_n) break;
            double xr[PT];
            const double yd = row_xy<PT, SCEN>(b, r, slab, cx, xr, sd);
            x_out[row] = 1.0f;
            for (int j = 1; j < PT; j++) x_out[(size_t)j * max_n + row] = slab[4 * (j - 1) + r];
            y_out[row] = (float)yd;
        }
    }
}

// K0b: OLS on supplied bytes (introspect::fit_provided_data, introspect.rs:174-293): narrow to f32, widen, same
// sufficient statistics + solver as the power kernel.  One thread.
template <int PT>
__global__ void fit_data_kernel(const double* __restrict__ x /* col-major n × P */, const double* __restrict__ y, int n,
                                double* __restrict__ out /* beta[P], se[P], t_sq[P], sigma_sq, sst, rss, converged */) {
    constexpr int NACC = PT * (PT + 1) / 2;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double acc[NACC], xty[PT];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < PT; i++) xty[i] = 0.0;
    double yty = 0.0, sum_y = 0.0;
    for (int row = 0; row < n; row++) {
        double xr[PT];
#pragma unroll
        for (int j = 0; j < PT; j++) xr[j] = (double)(float)x[(size_t)j * n + row];
        const double yd = (double)(float)y[row];
#pragma unroll
        for (int a = 0; a < PT; a++) {
#pragma unroll
            for (int c = 0; c <= a; c++) acc[tri(a, c)] = fma(xr[a], xr[c], acc[tri(a, c)]);
            xty[a] = fma(xr[a], yd, xty[a]);
        }
        yty = fma(yd, yd, yty);
        sum_y += yd;
    }
    RegSolver<PT> sv;
    sv.factor_and_solve(acc, xty, yty, sum_y, (uint32_t)n);
    for (int t = 0; t < PT; t++) {
        const double t2 = sv.t_sq(t, -1);
        double b = 0.0;
#pragma unroll
        for (int i = 0; i < PT; i++)
            if (i == t) b = sv.beta[i];
        out[t] = b;
        out[PT + t] = sqrt(b * b / t2);
        out[2 * PT + t] = t2;
    }
    out[3 * PT] = sv.out.sigma_sq;
    out[3 * PT + 1] = sv.out.sst;
    out[3 * PT + 2] = sv.out.rss;
    out[3 * PT + 3] = (double)sv.out.converged;
}

}  // namespace mcp
