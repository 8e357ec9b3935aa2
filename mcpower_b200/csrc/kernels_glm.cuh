// kernels_glm.cuh — batched IRLS power kernel (logit / probit / Poisson), sm_100a.
//
// Replaces the unclustered GLM arm of run_one_sim (crates/engine-core/src/batch.rs:1562-2135) and
// glmm::glm_irls_fit (call site batch.rs:1969-1981).  Two kernels per wave (the quad kernel in PLANE mode: one study per lane quad;
// glm_fit_kernel: one (study, sample size) fit per thread):
//   1. the study's max_n rows are generated once (same gen_block/row_xy as the OLS kernel, so the data
//      plane is bit-identical to the reference fixtures) and parked in an HBM scratch plane laid out
//      [row][column][sim-in-wave] — the 32 studies of a warp read consecutive floats, so every IRLS pass
//      is fully coalesced;
//   2. for every sample size (the reference refits each (sim, N) on the row prefix from the truth start,
//      batch.rs:1563-1577) IRLS runs with per-lane convergence masks: X'WX / X'Wz / deviance live in
//      registers (P is a template parameter), one Cholesky per iteration, no step-halving, W clamp,
//      |beta| > 30 ⇒ not converged (batch.rs:1942-1944, 4550-4555);
//   3. Wald z², corrections, the deviance LRT against chi2(p-1), sparse-factor exclusion and the separation
//      fallback (drop the sparsest included factor, cold refit; batch.rs:1994-2028) are decided in-kernel.
// IRLS constants (25 iterations, |ddev|/(|dev|+0.1) < 1e-8, weight clamp 1e-10) follow R's glm.fit: the crate's
// own values are not in the reference tree.  Every operation mirrors the test oracle's glm_irls.
#pragma once
#include "kernels.cuh"
#include "kernels_q4.cuh"
#include <type_traits>

namespace mcp {

#define MCP_GLM_MAXIT 25
#define MCP_GLM_TOL 1e-8
#define MCP_GLM_WCLAMP 1e-10
#define MCP_GLM_BETA_CAP 30.0

// family 0 logit, 1 probit, 2 poisson(log)
__device__ __forceinline__ void glm_row(int family, double eta, double y, double& w, double& z, double& dev) {
    if (family == 0) {
        const double e = own_exp(-fabs(eta));
        const double inv = 1.0 / (1.0 + e);
        const double mu = eta >= 0.0 ? inv : e * inv;
        w = mu * (1.0 - mu);
        if (w < MCP_GLM_WCLAMP) w = MCP_GLM_WCLAMP;
        z = eta + (y - mu) / w;
        dev = y > 0.5 ? -2.0 * own_log(mu) : -2.0 * own_log(1.0 - mu);
    } else if (family == 1) {
        const double mu = phi(eta);
        double d = own_exp(-0.5 * eta * eta) * 0.3989422804014327;
        if (d < 1e-300) d = 1e-300;
        double v = mu * (1.0 - mu);
        if (v < MCP_GLM_WCLAMP) v = MCP_GLM_WCLAMP;
        w = d * d / v;
        z = eta + (y - mu) / d;
        dev = y > 0.5 ? -2.0 * own_log(mu) : -2.0 * own_log(1.0 - mu);
    } else {
        const double mu = own_exp(eta);
        w = mu;
        if (w < MCP_GLM_WCLAMP) w = MCP_GLM_WCLAMP;
        z = eta + (y - mu) / w;
        dev = y > 0.0 ? 2.0 * (y * own_log(y / mu) - (y - mu)) : 2.0 * mu;
    }
}

// One IRLS row, branch-free (selects only), per family: the same arithmetic as glm_row, so that two rows evaluated back to
// back form one basic block and their exp / division / log chains interleave.
template <int FAM>
__device__ __forceinline__ void glm_row_sel(double eta, double y, double& w, double& z, double& dev) {
    if (FAM == 0) {
        const double e = own_exp_sel(-fabs(eta));
        const double inv = 1.0 / (1.0 + e);
        const double mu = eta >= 0.0 ? inv : e * inv;
        double ww = mu * (1.0 - mu);
        ww = ww < MCP_GLM_WCLAMP ? MCP_GLM_WCLAMP : ww;
        w = ww;
        z = eta + (y - mu) / ww;
        dev = -2.0 * own_log_sel(y > 0.5 ? mu : 1.0 - mu);
    } else if (FAM == 1) {
        const double mu = phi_sel(eta);
        double d = own_exp_sel(-0.5 * eta * eta) * 0.3989422804014327;
        d = d < 1e-300 ? 1e-300 : d;
        double v = mu * (1.0 - mu);
        v = v < MCP_GLM_WCLAMP ? MCP_GLM_WCLAMP : v;
        w = d * d / v;
        z = eta + (y - mu) / d;
        dev = -2.0 * own_log_sel(y > 0.5 ? mu : 1.0 - mu);
    } else {
        const double mu = own_exp_sel(eta);
        double ww = mu;
        ww = ww < MCP_GLM_WCLAMP ? MCP_GLM_WCLAMP : ww;
        w = ww;
        z = eta + (y - mu) / ww;
        const double dpos = 2.0 * (y * own_log_sel(y / mu) - (y - mu));  // selected away for y == 0 (log 0 → −inf, no trap)
        dev = y > 0.0 ? dpos : 2.0 * mu;
    }
}

struct GlmOut {
    int converged, iters;
    double deviance, deviance_null;
};

__device__ __forceinline__ double glm_null_deviance(int family, double s0, double s1, double n) {
    const double ybar = s0 / n;
    if (family == 2) return ybar > 0.0 ? 2.0 * (s1 - s0 * own_log(ybar)) : 0.0;
    const double n1 = s0, n0 = n - s0;
    double dn = 0.0;
    if (n1 > 0.0) dn += n1 * own_log(ybar);
    if (n0 > 0.0) dn += n0 * own_log(1.0 - ybar);
    return -2.0 * dn;
}

// Full-model IRLS with register-resident X'WX (all indices static).  `plane` points at this thread's element of
// row 0 / column 0; element (row, col) is plane[(row * PT + col) * stride]; columns 0..PT-2 = X[:,1..], PT-1 = y.
template <int PT>
__device__ __forceinline__ void glm_irls_reg(const float* __restrict__ plane, size_t stride, int n, int family, bool truth_start,
                                             RegSolver<PT>& sv, double (&beta)[PT], GlmOut& out) {
    constexpr int NACC = PT * (PT + 1) / 2;
    const DevicePlan& pl = c_plan;
#pragma unroll
    for (int j = 0; j < PT; j++) beta[j] = truth_start ? pl.beta[j] : 0.0;
    double dev_prev = __longlong_as_double(0x7ff0000000000000LL), dev = dev_prev, s0 = 0.0, s1 = 0.0;
    int conv = 0, it = 0;
    bool active = true;  // per-lane convergence mask: finished lanes idle until the warp's slowest study is done
    for (it = 0; it < MCP_GLM_MAXIT && active; it++) {
        double A[NACC], b[PT];
#pragma unroll
        for (int i = 0; i < NACC; i++) A[i] = 0.0;
#pragma unroll
        for (int i = 0; i < PT; i++) b[i] = 0.0;
        dev = 0.0;
        // Two rows per iteration (three and four were measured slower, profiles/r02_glm_kernel.md).  Every accumulator still sees
        // the rows in order (row i, then row i + 1), so the sums are bit-identical to a one-row loop; the two rows' exp / division / log chains are independent and interleave.  Rows
        // i + 2 / i + 3 are fetched (coalesced, L2 / HBM) while rows i / i + 1 are processed.  The family is uniform over the
        // launch: one of the three instantiations runs.
        auto pass = [&](auto fam_tag) {
            constexpr int FAM = decltype(fam_tag)::value;
            float nx0[PT], nx1[PT];
#pragma unroll
            for (int j = 0; j < PT; j++) {
                nx0[j] = n > 0 ? plane[(size_t)j * stride] : 0.0f;
                nx1[j] = n > 1 ? plane[(size_t)(PT + j) * stride] : 0.0f;
            }
            int i = 0;
            for (; i + 1 < n; i += 2) {
                float c0[PT], c1[PT];
#pragma unroll
                for (int j = 0; j < PT; j++) {
                    c0[j] = nx0[j];
                    c1[j] = nx1[j];
                }
                if (i + 2 < n) {
                    const float* rowp = plane + (size_t)(i + 2) * PT * stride;
#pragma unroll
                    for (int j = 0; j < PT; j++) nx0[j] = rowp[(size_t)j * stride];
                }
                if (i + 3 < n) {
                    const float* rowp = plane + (size_t)(i + 3) * PT * stride;
#pragma unroll
                    for (int j = 0; j < PT; j++) nx1[j] = rowp[(size_t)j * stride];
                }
                double x0[PT], x1[PT];
                x0[0] = 1.0;
                x1[0] = 1.0;
#pragma unroll
                for (int j = 1; j < PT; j++) {
                    x0[j] = (double)c0[j - 1];
                    x1[j] = (double)c1[j - 1];
                }
                const double y0 = (double)c0[PT - 1], y1 = (double)c1[PT - 1];
                double eta0 = 0.0, eta1 = 0.0;
#pragma unroll
                for (int j = 0; j < PT; j++) {
                    eta0 = fma(x0[j], beta[j], eta0);
                    eta1 = fma(x1[j], beta[j], eta1);
                }
                double w0, z0, d0, w1, z1, d1;
                glm_row_sel<FAM>(eta0, y0, w0, z0, d0);
                glm_row_sel<FAM>(eta1, y1, w1, z1, d1);
                dev += d0;
                dev += d1;
                if (it == 0) {
                    s0 += y0;
                    if (FAM == 2 && y0 > 1.0) s1 += y0 * own_log(y0);
                    s0 += y1;
                    if (FAM == 2 && y1 > 1.0) s1 += y1 * own_log(y1);
                }
#pragma unroll
                for (int a = 0; a < PT; a++) {
                    const double wa = w0 * x0[a];
#pragma unroll
                    for (int c = 0; c <= a; c++) A[tri(a, c)] = fma(wa, x0[c], A[tri(a, c)]);
                    b[a] = fma(wa, z0, b[a]);
                }
#pragma unroll
                for (int a = 0; a < PT; a++) {
                    const double wa = w1 * x1[a];
#pragma unroll
                    for (int c = 0; c <= a; c++) A[tri(a, c)] = fma(wa, x1[c], A[tri(a, c)]);
                    b[a] = fma(wa, z1, b[a]);
                }
            }
            if (i < n) {  // odd row count: the last row alone (it sits in nx0)
                double x0[PT];
                x0[0] = 1.0;
#pragma unroll
                for (int j = 1; j < PT; j++) x0[j] = (double)nx0[j - 1];
                const double y0 = (double)nx0[PT - 1];
                double eta0 = 0.0;
#pragma unroll
                for (int j = 0; j < PT; j++) eta0 = fma(x0[j], beta[j], eta0);
                double w0, z0, d0;
                glm_row_sel<FAM>(eta0, y0, w0, z0, d0);
                dev += d0;
                if (it == 0) {
                    s0 += y0;
                    if (FAM == 2 && y0 > 1.0) s1 += y0 * own_log(y0);
                }
#pragma unroll
                for (int a = 0; a < PT; a++) {
                    const double wa = w0 * x0[a];
#pragma unroll
                    for (int c = 0; c <= a; c++) A[tri(a, c)] = fma(wa, x0[c], A[tri(a, c)]);
                    b[a] = fma(wa, z0, b[a]);
                }
            }
        };
        if (family == 0) pass(std::integral_constant<int, 0>{});
        else if (family == 1) pass(std::integral_constant<int, 1>{});
        else pass(std::integral_constant<int, 2>{});
        if (isinf(dev) || dev != dev) {
            active = false;
            break;
        }
        if (it > 0 && fabs(dev - dev_prev) / (fabs(dev) + 0.1) < MCP_GLM_TOL) {
            sv.factor_and_solve(A, b, 0.0, 0.0, (uint32_t)n);  // X'WX at the converged beta → Wald covariance
            conv = sv.out.converged;
            active = false;
            break;
        }
        dev_prev = dev;
        sv.factor_and_solve(A, b, 0.0, 0.0, (uint32_t)n);
        if (!sv.out.converged) {
            active = false;
            break;
        }
        bool bad = false;
#pragma unroll
        for (int j = 0; j < PT; j++) {
            beta[j] = sv.beta[j];
            if (isinf(beta[j]) || beta[j] != beta[j] || fabs(beta[j]) > MCP_GLM_BETA_CAP) bad = true;
        }
        if (bad) {
            active = false;
            break;
        }
    }
    out.converged = conv;
    out.iters = it;
    out.deviance = dev;
    out.deviance_null = glm_null_deviance(family, s0, s1, (double)n);
}

// Reduced-model IRLS (exclusion / separation fallback / test design): runtime p on local memory.  Cold path.
template <int PT>
static __device__ __noinline__ void glm_irls_mem(const float* __restrict__ plane, size_t stride, int n, int family, bool truth_start,
                                          const McpVariant& vr, MemSolver<PT>& ms, double* beta, GlmOut& out) {
    constexpr int NACC = PT * (PT + 1) / 2;
    const DevicePlan& pl = c_plan;
    const int p = vr.p_red;
    for (int j = 0; j < p; j++) beta[j] = truth_start ? pl.beta[vr.kept[j]] : 0.0;
    double dev_prev = __longlong_as_double(0x7ff0000000000000LL), dev = dev_prev, s0 = 0.0, s1 = 0.0;
    int conv = 0, it = 0;
    double A[NACC], b[PT], xr[PT], btmp[PT];
    for (it = 0; it < MCP_GLM_MAXIT; it++) {
        for (int i = 0; i < p * (p + 1) / 2; i++) A[i] = 0.0;
        for (int i = 0; i < p; i++) b[i] = 0.0;
        dev = 0.0;
        for (int i = 0; i < n; i++) {
            const float* rowp = plane + (size_t)i * PT * stride;
            double eta = 0.0;
            for (int j = 0; j < p; j++) {
                const int c = vr.kept[j];
                xr[j] = c == 0 ? 1.0 : (double)rowp[(size_t)(c - 1) * stride];
                eta = fma(xr[j], beta[j], eta);
            }
            const double yi = (double)rowp[(size_t)(PT - 1) * stride];
            double w, z, di;
            glm_row(family, eta, yi, w, z, di);
            dev += di;
            if (it == 0) {
                s0 += yi;
                if (family == 2 && yi > 1.0) s1 += yi * own_log(yi);
            }
            for (int a = 0; a < p; a++) {
                const double wa = w * xr[a];
                for (int c = 0; c <= a; c++) A[tri(a, c)] = fma(wa, xr[c], A[tri(a, c)]);
                b[a] = fma(wa, z, b[a]);
            }
        }
        if (isinf(dev) || dev != dev) break;
        if (it > 0 && fabs(dev - dev_prev) / (fabs(dev) + 0.1) < MCP_GLM_TOL) {
            for (int j = 0; j < p; j++) btmp[j] = beta[j];
            ms.factor_and_solve(A, b, p, 0.0, 0.0, (uint32_t)n);
            for (int j = 0; j < p; j++) ms.beta[j] = btmp[j];
            conv = ms.out.converged;
            break;
        }
        dev_prev = dev;
        ms.factor_and_solve(A, b, p, 0.0, 0.0, (uint32_t)n);
        if (!ms.out.converged) break;
        bool bad = false;
        for (int j = 0; j < p; j++) {
            beta[j] = ms.beta[j];
            if (isinf(beta[j]) || beta[j] != beta[j] || fabs(beta[j]) > MCP_GLM_BETA_CAP) bad = true;
        }
        if (bad) break;
    }
    for (int j = 0; j < p; j++) ms.beta[j] = beta[j];
    out.converged = conv;
    out.iters = it;
    out.deviance = dev;
    out.deviance_null = glm_null_deviance(family, s0, s1, (double)n);
}

// Decisions for one GLM fit: Wald z² vs the (df-free) normal thresholds, corrections, deviance LRT
// (dev_null − dev > chi2(p_red − 1), batch.rs:2094-2126).
static __device__ __noinline__ void emit_fit_glm(const SimOut& so, int g, int variant, const GlmOut& go, double* all_t, const uint8_t* excl_flags) {
    const DevicePlan& pl = c_plan;
    FitOut fo;
    fo.converged = go.converged;
    fo.df_resid = 1.0;
    fo.sigma_sq = 1.0;
    fo.rss = 0.0;
    fo.sst = 0.0;
    int ov = 0;
    const bool finite = !(isinf(go.deviance) || go.deviance != go.deviance || isinf(go.deviance_null) || go.deviance_null != go.deviance_null);
    if (go.converged && finite) {
        const double lrt = go.deviance_null - go.deviance;
        if (!(isinf(lrt) || lrt != lrt) && lrt > pl.crit_overall[(size_t)variant * pl.n_ss + g]) ov = 1;
    }
    emit_fit(so, g, variant, fo, pl.variants[variant].p_red, all_t, excl_flags, ov);
}

// introspect::fit_provided_data twin, GLM arm (introspect.rs:395-447): ONE supplied dataset through the same register-resident
// IRLS as the power kernel, from the same spec-derived truth start (beta0 = effect_sizes).  plane = [row][column] f32, columns
// x_1 .. x_{P-1}, y (stride 1).  out: beta[P] | Var(beta)[P] | z²[P] | converged, boundary (0), iterations, sigma² (NaN),
// tau² (NaN), deviance, null deviance.
template <int PT>
__global__ void glm_fit_data_kernel(const float* __restrict__ plane, int n, double* __restrict__ out) {
    static_assert(PT <= MCP_MAX_TEMPLATE_P, "exact widths only");
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const DevicePlan& pl = c_plan;
    const int family = pl.outcome == O_COUNT ? 2 : (pl.link_probit ? 1 : 0);
    RegSolver<PT> sv;
    double beta[PT];
    GlmOut go;
    glm_irls_reg<PT>(plane, 1, n, family, /*truth_start=*/true, sv, beta, go);
#pragma unroll
    for (int j = 0; j < PT; j++) sv.beta[j] = beta[j];
    sv.out.sigma_sq = 1.0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int t = 0; t < PT; t++) {
        double b = 0.0;
#pragma unroll
        for (int i = 0; i < PT; i++)
            if (i == t) b = beta[i];
        out[t] = b;
        out[PT + t] = go.converged ? sv.var_of(t) : nan;
        out[2 * PT + t] = go.converged ? sv.t_sq(t, -1) : nan;
    }
    out[3 * PT] = (double)go.converged;
    out[3 * PT + 1] = 0.0;
    out[3 * PT + 2] = (double)go.iters;
    out[3 * PT + 3] = nan;
    out[3 * PT + 4] = nan;
    out[3 * PT + 5] = go.deviance;
    out[3 * PT + 6] = go.deviance_null;
    for (int i = 7; i <= 13; i++) out[3 * PT + i] = nan;
}

// Stage 1 is the lane-quad kernel in PLANE mode (kernels_q4.cuh): it generates the study's max_n rows once — the same gen_block /
// row_xy as the OLS kernels, so the data plane is bit-identical to the reference fixtures — and parks the f32 plane in HBM,
// [row][column][study-in-wave].

// Stage 2 — one (study, sample size) fit per thread, grid = (blocks of studies, n_ss): the fits of one study at its different
// sample sizes are independent (each restarts from the truth start on its own row prefix, batch.rs:1563-1577), so they run
// concurrently, the threads of a warp all fit the same n (same row-loop trip count), and this kernel carries no generation
// code.  blockIdx.y counts DOWN the grid so the longest fits are scheduled first.
template <int PT, int MINB>
__global__ void __launch_bounds__(MCP_BLOCK, MINB) glm_fit_kernel(uint32_t n_sims, const float* __restrict__ scratch) {
    const DevicePlan& pl = c_plan;
    if (mcp_cancelled()) return;
    extern __shared__ __align__(16) float smem[];
    unsigned int* cnt_s = nullptr;
    if (pl.smem_counters) {
        cnt_s = reinterpret_cast<unsigned int*>(smem);
        for (int i = threadIdx.x; i < pl.n_ss * pl.counter_width; i += MCP_BLOCK) cnt_s[i] = 0u;
        __syncthreads();
    }
    const uint32_t sim = blockIdx.x * MCP_BLOCK + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * MCP_BLOCK;
    const int g = pl.n_ss - 1 - (int)blockIdx.y;
    if (sim < n_sims) {
        SimOut so{cnt_s, (uint64_t)sim};
        const float* plane = scratch + sim;
        const int family = pl.outcome == O_COUNT ? 2 : (pl.link_probit ? 1 : 0);
        uint32_t lvl_counts[MCP_MAX_LEVELS];
        uint8_t excl[MCP_MAX_FACTORS];
        const bool track_excl = pl.n_factors > 0 && pl.min_level_count > 0;
        for (int i = 0; i < MCP_MAX_LEVELS; i++) lvl_counts[i] = 0;
        const int n = (int)pl.ss[g];
        for (int f = 0; f < MCP_MAX_FACTORS; f++) excl[f] = 0;
        if (track_excl) {  // per-level counts over rows [0, n) (batch.rs:2451-2496)
            for (int i = 0; i < n; i++) {
                const float* rowp = plane + (size_t)i * PT * stride;
                for (int f = 0; f < pl.n_factors; f++) {
                    const int c0 = pl.f_col[f] - 1, nl = pl.f_levels[f];
                    int lvl = 0;
                    for (int d = 0; d < nl - 1; d++)
                        if (rowp[(size_t)(c0 + d) * stride] == 1.0f) { lvl = d + 1; break; }
                    lvl_counts[pl.f_off[f] + lvl]++;
                }
            }
            for (int f = 0; f < pl.n_factors; f++) {
                uint32_t mn = 0xffffffffu;
                for (int q = 0; q < pl.f_levels[f]; q++) mn = min(mn, lvl_counts[pl.f_off[f] + q]);
                excl[f] = mn < (uint32_t)pl.min_level_count;
            }
        }
        GlmOut go;
        double all_t[MCP_MAX_T];
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int t = 0; t < pl.T; t++) all_t[t] = nan;
        int variant = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            variant = pl.base_variant;
            if (track_excl) {
                variant = 0;
                for (int f = 0; f < pl.n_factors; f++) variant |= (excl[f] != 0 ? 1 : 0) << f;
            }
            const McpVariant& vr = pl.variants[variant];
            bool full = false;
            if constexpr (PT <= MCP_MAX_TEMPLATE_P) full = vr.p_red == PT;  // padded widths always fit the kept (real) columns
            if constexpr (PT <= MCP_MAX_TEMPLATE_P) if (full) {
                RegSolver<PT> sv;
                double beta[PT];
                glm_irls_reg<PT>(plane, stride, n, family, attempt == 0, sv, beta, go);
                if (go.converged) {
#pragma unroll
                    for (int j = 0; j < PT; j++) sv.beta[j] = beta[j];
                    sv.out.sigma_sq = 1.0;
                    for (int t = 0; t < pl.n_marg; t++) all_t[t] = sv.t_sq(pl.target[t], -1);
                }
            }
            if (!full) {
                MemSolver<PT> ms;
                double beta[PT];
                glm_irls_mem<PT>(plane, stride, n, family, attempt == 0, vr, ms, beta, go);
                if (go.converged) {
                    ms.out.sigma_sq = 1.0;
                    for (int t = 0; t < pl.n_marg; t++) {
                        const int c = vr.remap[pl.target[t]];
                        all_t[t] = c >= 0 ? ms.t_sq(c, -1) : nan;
                    }
                }
            }
            if (go.converged || attempt == 1) break;
            if (pl.min_level_count == 0) break;
            // separation fallback: drop the included factor with the smallest min level count (batch.rs:2572-2597)
            int best = -1;
            uint32_t best_min = 0;
            for (int f = 0; f < pl.n_factors; f++) {
                if (excl[f] == 0) {
                    uint32_t mn = 0xffffffffu;
                    for (int q = 0; q < pl.f_levels[f]; q++) mn = min(mn, lvl_counts[pl.f_off[f] + q]);
                    if (best < 0 || mn < best_min) { best = f; best_min = mn; }
                }
            }
            if (best < 0) break;
            excl[best] = 2;
        }
        emit_fit_glm(so, g, variant, go, all_t, excl);
    }
    mcp_report_warp((blockIdx.x * MCP_BLOCK + threadIdx.x) & ~31u, 32u, n_sims);  // unit = one fit; the host divides by n_ss
    if (cnt_s) {
        __syncthreads();
        for (int i = threadIdx.x; i < pl.n_ss * pl.counter_width; i += MCP_BLOCK) {
            const unsigned v = cnt_s[i];
            if (v) atomicAdd(&pl.counters[i], (unsigned long long)v);
        }
    }
}

}  // namespace mcp
