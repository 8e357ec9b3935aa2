// kernels_lmm.cuh — random-intercept linear mixed model power kernel (profiled REML), sm_100a.
//
// Replaces the MLE arm of run_one_sim (crates/engine-core/src/batch.rs:1214-1561) for the random-intercept case
// (BASELINE config 5).  The reference fits through glmm::fit_on + BOBYQA over theta (crates not in the tree); for
// one random intercept the profiled REML deviance is a closed form in theta = tau/sigma over cluster-level
// sufficient statistics (scripts/lme_handcompute_theta1.py:13-36):
//     A(theta) = X'X − Σ_c s_c Sx_c Sx_c',  b(theta) = X'y − Σ_c s_c Sx_c Sy_c,  c(theta) = y'y − Σ_c s_c Sy_c²,
//     s_c = theta² / (1 + theta² n_c),   dev = Σ_c log(1 + theta² n_c) + log|A| + (n − p) log sigma²(theta).
// One simulated study per thread.  X'X / X'y / y'y stream in registers exactly as in the OLS kernel; the per-cluster
// sums Sx_c, Sy_c live in an HBM scratch plane [cluster][column][sim] (coalesced across the studies of a warp) and
// grow with the row prefix (cluster ids are N-independent in both sizing regimes, generation.rs cluster_of_row).
// Both sizing regimes produce at most two distinct cluster sizes per prefix, so the cluster sums are folded into
// two groups once per fit and every deviance evaluation is O(p²).  theta is found by Brent's fmin on [0, 30];
// z-based Wald thresholds, corrections, the joint Wald chi-square and the boundary flag are decided in-kernel.
// Every operation mirrors the test oracle's lmm_fit.
#pragma once
#include "kernels.cuh"
#include "kernels_q4.cuh"

namespace mcp {

#define MCP_LMM_THETA_MAX 30.0
#define MCP_LMM_TOL 1e-10
#define MCP_LMM_MAXIT 200
// resident CTAs per SM the fit kernel is compiled for.  The REML search is latency-bound (serial Brent steps, six logs per
// evaluation), so occupancy pays until the spills bite; measured at P = 3 (config 5): unbounded (168 registers) 1.15e8,
// 4 → 1.21e8, 6 → 1.27e8, 8 → 1.32e8, 10 → 1.37e8, 12 → 1.33e8 fits/s
#define MCP_LMM_FIT_MINB(PT) ((PT) <= 3 ? 10 : ((PT) <= 4 ? 8 : ((PT) <= 6 ? 6 : 4)))

template <int PT>
struct LmmCtx {
    static constexpr int NACC = PT * (PT + 1) / 2;
    int p, ng;
    uint32_t n_rows;
    double xtx[NACC], xty[PT], yty;
    double grp_n[2], grp_m[2];
    double M[2][NACC], v[2][PT], w[2];
    double A[NACC], b[PT];
    MemSolver<PT> ms;
};

template <int PT>
static __device__ __noinline__ double lmm_dev(LmmCtx<PT>& c, double theta) {
    const int p = c.p, nt = p * (p + 1) / 2;
    const double t2 = theta * theta;
    double s[2] = {0.0, 0.0};
    for (int g = 0; g < c.ng; g++) s[g] = t2 / (1.0 + t2 * c.grp_n[g]);
    for (int e = 0; e < nt; e++) {
        double a = c.xtx[e];
        for (int g = 0; g < c.ng; g++) a = fma(-s[g], c.M[g][e], a);
        c.A[e] = a;
    }
    for (int j = 0; j < p; j++) {
        double a = c.xty[j];
        for (int g = 0; g < c.ng; g++) a = fma(-s[g], c.v[g][j], a);
        c.b[j] = a;
    }
    double cc = c.yty;
    for (int g = 0; g < c.ng; g++) cc = fma(-s[g], c.w[g], cc);
    c.ms.factor_and_solve(c.A, c.b, p, cc, 0.0, c.n_rows);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (!c.ms.out.converged || !(c.ms.out.rss > 0.0) || isinf(c.ms.out.rss)) return inf;
    double ldv = 0.0;
    for (int g = 0; g < c.ng; g++) ldv += c.grp_m[g] * own_log(1.0 + t2 * c.grp_n[g]);
    double lda = 0.0;
    for (int j = 0; j < p; j++) lda += own_log(c.ms.L[tri(j, j)]);
    return ldv + 2.0 * lda + c.ms.out.df_resid * own_log(c.ms.out.sigma_sq);
}

// The same profiled deviance for the FULL model (p == PT), value only, on register copies of the sufficient statistics:
// every loop has static bounds, nothing goes through local memory, and the solve stops at what the deviance needs (Cholesky,
// forward substitution, rss).  Operation for operation the arithmetic of lmm_dev + MemSolver::factor_and_solve, so the Brent
// search sees bit-identical values; the final evaluation at theta-hat still runs lmm_dev, which leaves L / beta / sigma².
template <int PT>
struct LmmRegStats {
    static constexpr int NACC = PT * (PT + 1) / 2;
    int ng;
    double n_rows;
    double xtx[NACC], xty[PT], yty;
    double grp_n[2], grp_m[2];
    double M[2][NACC], v[2][PT], w[2];
    __device__ __forceinline__ void load(const LmmCtx<PT>& c) {
        ng = c.ng;
        n_rows = (double)c.n_rows;
        yty = c.yty;
#pragma unroll
        for (int e = 0; e < NACC; e++) {
            xtx[e] = c.xtx[e];
            M[0][e] = c.M[0][e];
            M[1][e] = c.M[1][e];
        }
#pragma unroll
        for (int j = 0; j < PT; j++) {
            xty[j] = c.xty[j];
            v[0][j] = c.v[0][j];
            v[1][j] = c.v[1][j];
        }
#pragma unroll
        for (int g = 0; g < 2; g++) {
            grp_n[g] = c.grp_n[g];
            grp_m[g] = c.grp_m[g];
            w[g] = c.w[g];
        }
    }
    __device__ __forceinline__ double dev(double theta) const {
        const double t2 = theta * theta;
        double s[2] = {0.0, 0.0};
#pragma unroll
        for (int g = 0; g < 2; g++)
            if (g < ng) s[g] = t2 / (1.0 + t2 * grp_n[g]);
        double A[NACC], b[PT];
#pragma unroll
        for (int e = 0; e < NACC; e++) {
            double a = xtx[e];
#pragma unroll
            for (int g = 0; g < 2; g++)
                if (g < ng) a = fma(-s[g], M[g][e], a);
            A[e] = a;
        }
#pragma unroll
        for (int j = 0; j < PT; j++) {
            double a = xty[j];
#pragma unroll
            for (int g = 0; g < 2; g++)
                if (g < ng) a = fma(-s[g], v[g][j], a);
            b[j] = a;
        }
        double cc = yty;
#pragma unroll
        for (int g = 0; g < 2; g++)
            if (g < ng) cc = fma(-s[g], w[g], cc);
        // MemSolver::factor_and_solve, p = PT: Cholesky–Banachiewicz, forward substitution, rss
        double L[NACC], inv_d[PT], wv[PT];
        int ok = 1;
        double dmin = __longlong_as_double(0x7ff0000000000000LL), dmax = 0.0;
#pragma unroll
        for (int i = 0; i < PT; i++) {
#pragma unroll
            for (int j = 0; j <= i; j++) {
                double sacc = A[tri(i, j)];
#pragma unroll
                for (int k = 0; k < j; k++) sacc = fma(-L[tri(i, k)], L[tri(j, k)], sacc);
                if (i == j) {
                    if (!(sacc > 0.0) || isinf(sacc)) {
                        ok = 0;
                        sacc = 1.0;
                    }
                    const double d = sqrt(sacc);
                    L[tri(i, i)] = d;
                    inv_d[i] = 1.0 / d;
                    dmin = d < dmin ? d : dmin;
                    dmax = d > dmax ? d : dmax;
                } else {
                    L[tri(i, j)] = sacc * inv_d[j];
                }
            }
        }
        if (ok && !(dmin / dmax >= MCP_EPS_RANK)) ok = 0;
#pragma unroll
        for (int i = 0; i < PT; i++) {
            double sacc = b[i];
#pragma unroll
            for (int k = 0; k < i; k++) sacc = fma(-L[tri(i, k)], wv[k], sacc);
            wv[i] = sacc * inv_d[i];
        }
        double rss = cc;
#pragma unroll
        for (int i = 0; i < PT; i++) rss = fma(-wv[i], wv[i], rss);
        const double df_resid = n_rows - (double)PT;
        const double sigma_sq = rss / df_resid;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        if (!ok || !(rss > 0.0) || isinf(rss)) return inf;
        double ldv = 0.0;
#pragma unroll
        for (int g = 0; g < 2; g++)
            if (g < ng) ldv += grp_m[g] * own_log(1.0 + t2 * grp_n[g]);
        double lda = 0.0;
#pragma unroll
        for (int j = 0; j < PT; j++) lda += own_log(L[tri(j, j)]);
        return ldv + 2.0 * lda + df_resid * own_log(sigma_sq);
    }
};

// Brent's fmin on [a, b] (golden section + parabolic interpolation) over any deviance evaluator, followed by the evaluation
// at theta = 0 that decides the boundary case.  Written as a small state machine around ONE call site of the evaluator: the
// register-resident evaluator is inlined, and three copies of it (first point, search loop, theta = 0) in divergent warps were
// the kernel's top stall (instruction fetch).  Same evaluations in the same order as the textbook loop.
template <typename F>
__device__ __forceinline__ double lmm_fmin_over(F&& dev_at, double a, double b, double tol, double& fmin_out, double& f0_out) {
    const double cgold = 0.3819660112501051, eps = 1.4901161193847656e-08;
    double v = a + cgold * (b - a), w = v, x = v, e = 0.0, d = 0.0;
    double fx = 0.0, fv = 0.0, fw = 0.0;
    double u = x;
    int phase = 0, it = 0;  // phase 0: first point, 1: search, 2: theta = 0
    for (;;) {
        const double fu = dev_at(u);
        if (phase == 2) {
            f0_out = fu;
            break;
        }
        if (phase == 0) {
            fx = fv = fw = fu;
            phase = 1;
        } else {
            if (fu <= fx) {
                if (u >= x) a = x;
                else b = x;
                v = w; fv = fw; w = x; fw = fx; x = u; fx = fu;
            } else {
                if (u < x) a = u;
                else b = u;
                if (fu <= fw || w == x) { v = w; fv = fw; w = u; fw = fu; }
                else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
            }
            it++;
        }
        const double xm = 0.5 * (a + b), tol1 = eps * fabs(x) + tol / 3.0, tol2 = 2.0 * tol1;
        if (it >= MCP_LMM_MAXIT || fabs(x - xm) <= tol2 - 0.5 * (b - a)) {
            phase = 2;
            u = 0.0;
            continue;
        }
        bool golden = true;
        if (fabs(e) > tol1) {
            double r = (x - w) * (fx - fv), q = (x - v) * (fx - fw), pp = (x - v) * q - (x - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) pp = -pp;
            else q = -q;
            r = e;
            e = d;
            if (fabs(pp) < fabs(0.5 * q * r) && pp > q * (a - x) && pp < q * (b - x)) {
                d = pp / q;
                const double uu = x + d;
                if (uu - a < tol2 || b - uu < tol2) d = xm - x >= 0.0 ? tol1 : -tol1;
                golden = false;
            }
        }
        if (golden) {
            e = x >= xm ? a - x : b - x;
            d = cgold * e;
        }
        u = fabs(d) >= tol1 ? x + d : (d >= 0.0 ? x + tol1 : x - tol1);
    }
    fmin_out = fx;
    return x;
}

template <int PT>
static __device__ __noinline__ double lmm_fmin(LmmCtx<PT>& c, double a, double b, double tol, double& fmin_out, double& f0_out) {
    return lmm_fmin_over([&](double th) { return lmm_dev<PT>(c, th); }, a, b, tol, fmin_out, f0_out);
}
// full model, small P: the search runs on registers (see LmmRegStats)
template <int PT>
static __device__ __noinline__ double lmm_fmin_full(const LmmCtx<PT>& c, double a, double b, double tol, double& fmin_out, double& f0_out) {
    LmmRegStats<PT> rs;
    rs.load(c);
    return lmm_fmin_over([&](double t) { return rs.dev(t); }, a, b, tol, fmin_out, f0_out);
}

// rows in cluster k among the first n rows (generation.rs cluster_of_row)
__device__ __forceinline__ uint32_t lmm_cluster_rows(const DevicePlan& pl, int k, int n) {
    const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
    if (pl.sizing_fixed_size) {
        const int rem = n - k * cs;
        return (uint32_t)(rem <= 0 ? 0 : (rem < cs ? rem : cs));
    }
    return (uint32_t)(n / cs + (k < n % cs ? 1 : 0));
}

// Fold the cluster sums into (at most two) size groups and minimise the profiled deviance.
//   csum element (k, j) = csum[(k * (PT+1) + j) * stride]; kept == nullptr → full model.
template <int PT>
static __device__ __noinline__ void lmm_fit(LmmCtx<PT>& c, const double* __restrict__ csum, size_t stride, int n, int n_clusters,
                                            const int8_t* kept, int& converged, int& boundary, double* theta_out = nullptr,
                                            int cluster_doubles = PT + 1) {
    const DevicePlan& pl = c_plan;
    const int p = c.p, nt = p * (p + 1) / 2;
    c.ng = 0;
    for (int g = 0; g < 2; g++) {
        c.grp_n[g] = 0.0;
        c.grp_m[g] = 0.0;
        c.w[g] = 0.0;
        for (int e = 0; e < nt; e++) c.M[g][e] = 0.0;
        for (int j = 0; j < p; j++) c.v[g][j] = 0.0;
    }
    double sx[PT + 1];
    for (int k = 0; k < n_clusters; k++) {
        const uint32_t cnk = lmm_cluster_rows(pl, k, n);
        if (cnk == 0) continue;
        int g;
        if (c.ng == 0) { g = 0; c.ng = 1; c.grp_n[0] = (double)cnk; }
        else if ((double)cnk == c.grp_n[0]) g = 0;
        else { g = 1; if (c.ng == 1) { c.ng = 2; c.grp_n[1] = (double)cnk; } }
        c.grp_m[g] += 1.0;
        const double* base = csum + (size_t)k * cluster_doubles * stride;
        for (int j = 0; j < p; j++) sx[j] = base[(size_t)(kept ? kept[j] : j) * stride];
        sx[p] = base[(size_t)PT * stride];
        int t = 0;
        for (int a = 0; a < p; a++) {
            for (int bb = 0; bb <= a; bb++) { c.M[g][t] = fma(sx[a], sx[bb], c.M[g][t]); t++; }
            c.v[g][a] = fma(sx[a], sx[p], c.v[g][a]);
        }
        c.w[g] = fma(sx[p], sx[p], c.w[g]);
    }
    double fbest, f0;
    double th;
    bool searched = false;
    if constexpr (PT <= 6) {  // register-resident search (bit-identical values); wider models keep the generic evaluator
        if (p == PT) {
            th = lmm_fmin_full<PT>(c, 0.0, MCP_LMM_THETA_MAX, MCP_LMM_TOL, fbest, f0);
            searched = true;
        }
    }
    if (!searched) th = lmm_fmin<PT>(c, 0.0, MCP_LMM_THETA_MAX, MCP_LMM_TOL, fbest, f0);
    boundary = 0;
    converged = 1;
    if (!(fbest < f0)) { th = 0.0; boundary = 1 | (1 << 2); }  // tau-hat = 0: OLS fallback (spec.rs:562-565); component 0 pinned
    else if (MCP_LMM_THETA_MAX - th < 1e-6 * MCP_LMM_THETA_MAX) { boundary = 2; converged = 0; }
    const double f = lmm_dev<PT>(c, th);  // leaves L, beta, sigma² at theta-hat
    if (isinf(f) || f != f) converged = 0;
    if (theta_out) *theta_out = th;
}

// ---- random intercept + ONE random slope (q = 2): profiled REML over theta = (l00, l10, l11), Lambda = [[l00, 0], [l10, l11]].
// Replaces glmm::fit_on + BOBYQA for this shape (batch.rs:1214-1561, mixed_workspace.rs:203-272); the oracle's slope_dev /
// slope_nm / lmm_slope_fit are the op-for-op twins.  Per cluster: cs1 = sums of [x | y], cs2 = sums of x_s · [x | y];
// Z'Z = [[n_c, cs1[s]], [cs1[s], cs2[s]]]; M_c = Lambda' Z'Z Lambda + I = L_c L_c'; R_c = L_c^-1 Lambda' Z_c'[X | y];
// A = X'X − Σ R_c'R_c, b, c likewise; dev = Σ log det M_c + log det A + (n − p) log sigma².  Full model only: a reduced fit
// (sparse-factor exclusion, test design) strips the slope like the reference (batch.rs:1357-1375).
#define MCP_SLOPE_MAXQ 3
#define MCP_SLOPE_MAXD 8  // q (q + 1) / 2 for the primary grouping (+ one scale per extra grouping)
#define MCP_SLOPE_NM_EVAL_PER_DIM 300
#define MCP_SLOPE_THETA_MAX 30.0
template <int PT>
struct SlopeData {
    const double* cs;  // cluster k, block a (0: sums of [x | y]; a >= 1: of x_{s_a} · [x | y]), entry j at cs[((k·q + a)(PT+1) + j)·stride]
    size_t stride;
    int n, n_clusters, q;
    int slope_col[MCP_SLOPE_MAXQ];  // slope_col[a] = design column of slope a (a >= 1)
    int n_eval;
};

// theta = the lower triangle of Lambda row by row: (l00, l10, l11[, l20, l21, l22]); the diagonal enters as |.|.  Same operations
// in the same order as the oracle's slope_dev.
// (Q is a template parameter so that the q x q blocks live in registers: the runtime-q form ran 2.7x slower.)
template <int PT, int Q>
static __device__ __noinline__ double slope_dev_q(LmmCtx<PT>& c, SlopeData<PT>& d, const double* th) {
    const DevicePlan& pl = c_plan;
    constexpr int NT = PT * (PT + 1) / 2;
    constexpr int q = Q;
    double Lm[Q][Q];
#pragma unroll
    for (int i = 0, t = 0; i < q; i++)
#pragma unroll
        for (int j = 0; j <= i; j++, t++) Lm[i][j] = i == j ? fabs(th[t]) : th[t];
    for (int e = 0; e < NT; e++) c.A[e] = c.xtx[e];
    for (int j = 0; j < PT; j++) c.b[j] = c.xty[j];
    double cc = c.yty, ldv = 0.0;
    double r[Q][PT + 1];
    for (int k = 0; k < d.n_clusters; k++) {
        const uint32_t cnk = lmm_cluster_rows(pl, k, d.n);
        if (cnk == 0) continue;
        const double* sk[Q];
#pragma unroll
        for (int a = 0; a < q; a++) sk[a] = d.cs + (size_t)(k * q + a) * (PT + 1) * d.stride;
        double S[Q][Q], T[Q][Q], M[Q][Q], C[Q][Q];
        S[0][0] = (double)cnk;
#pragma unroll
        for (int a = 1; a < q; a++) {
            S[a][0] = S[0][a] = sk[0][(size_t)d.slope_col[a] * d.stride];
#pragma unroll
            for (int b = 1; b <= a; b++) S[a][b] = S[b][a] = sk[b][(size_t)d.slope_col[a] * d.stride];
        }
#pragma unroll
        for (int i = 0; i < q; i++)
#pragma unroll
            for (int j = 0; j < q; j++) {
                double acc = 0.0;
#pragma unroll
                for (int m = j; m < q; m++) acc += Lm[m][j] * S[i][m];
                T[i][j] = acc;
            }
#pragma unroll
        for (int i = 0; i < q; i++)
#pragma unroll
            for (int j = 0; j <= i; j++) {
                double acc = 0.0;
#pragma unroll
                for (int m = j; m < q; m++) acc += T[m][i] * Lm[m][j];
                M[i][j] = i == j ? acc + 1.0 : acc;
            }
#pragma unroll
        for (int j = 0; j < q; j++) {
            double dg = M[j][j];
#pragma unroll
            for (int m = 0; m < j; m++) dg -= C[j][m] * C[j][m];
            C[j][j] = sqrt(dg);
#pragma unroll
            for (int i = j + 1; i < q; i++) {
                double acc = M[i][j];
#pragma unroll
                for (int m = 0; m < j; m++) acc -= C[i][m] * C[j][m];
                C[i][j] = acc / C[j][j];
            }
        }
        {
            double l = own_log(C[0][0]);
#pragma unroll
            for (int j = 1; j < q; j++) l += own_log(C[j][j]);
            ldv += l;
        }
        for (int j = 0; j <= PT; j++) {
#pragma unroll
            for (int i = 0; i < q; i++) {
                double t = 0.0;
#pragma unroll
                for (int m = i; m < q; m++) t += Lm[m][i] * sk[m][(size_t)j * d.stride];
#pragma unroll
                for (int m = 0; m < i; m++) t -= C[i][m] * r[m][j];
                r[i][j] = t / C[i][i];
            }
        }
        int t = 0;
        for (int a = 0; a < PT; a++) {
            for (int bb = 0; bb <= a; bb++) {
#pragma unroll
                for (int i = 0; i < q; i++) c.A[t] = fma(-r[i][a], r[i][bb], c.A[t]);
                t++;
            }
#pragma unroll
            for (int i = 0; i < q; i++) c.b[a] = fma(-r[i][a], r[i][PT], c.b[a]);
        }
#pragma unroll
        for (int i = 0; i < q; i++) cc = fma(-r[i][PT], r[i][PT], cc);
    }
    d.n_eval++;
    c.ms.factor_and_solve(c.A, c.b, PT, cc, 0.0, c.n_rows);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (!c.ms.out.converged || !(c.ms.out.rss > 0.0) || isinf(c.ms.out.rss) || isinf(ldv) || ldv != ldv) return inf;
    double lda = 0.0;
    for (int j = 0; j < PT; j++) lda += own_log(c.ms.L[tri(j, j)]);
    return 2.0 * ldv + 2.0 * lda + c.ms.out.df_resid * own_log(c.ms.out.sigma_sq);
}

template <int PT>
static __device__ __forceinline__ double slope_dev(LmmCtx<PT>& c, SlopeData<PT>& d, const double* th) {
    return d.q == 2 ? slope_dev_q<PT, 2>(c, d, th) : slope_dev_q<PT, 3>(c, d, th);
}

// Nelder–Mead in D = q (q + 1) / 2 dimensions; same comparisons in the same order as the oracle's slope_nm.
template <int PT>
static __device__ __noinline__ double slope_nm(LmmCtx<PT>& c, SlopeData<PT>& d, double* x, double step, int max_eval) {
    const int D = d.q * (d.q + 1) / 2, V = D + 1;
    double sx[MCP_SLOPE_MAXD + 1][MCP_SLOPE_MAXD], sf[MCP_SLOPE_MAXD + 1];
    for (int i = 0; i < V; i++) {
        for (int j = 0; j < D; j++) sx[i][j] = x[j] + (i == j + 1 ? step : 0.0);
        sf[i] = slope_dev<PT>(c, d, sx[i]);
    }
    for (;;) {
        int lo = 0, hi = 0, nh;
        for (int i = 1; i < V; i++) {
            if (sf[i] < sf[lo]) lo = i;
            if (sf[i] > sf[hi] || !(sf[i] == sf[i])) hi = i;
        }
        nh = lo;
        for (int i = 0; i < V; i++)
            if (i != hi && (sf[i] > sf[nh])) nh = i;
        double size = 0.0;
        for (int i = 0; i < V; i++)
            for (int j = 0; j < D; j++) {
                const double dd = fabs(sx[i][j] - sx[lo][j]);
                if (dd > size) size = dd;
            }
        // (a cancelled batch is discarded: leaving the search early only shortens the wait — these fits run for milliseconds)
        if ((fabs(sf[hi] - sf[lo]) <= 1e-11 * (fabs(sf[lo]) + 1.0) && size <= 1e-7) || d.n_eval >= max_eval || mcp_cancelled()) {
            for (int j = 0; j < D; j++) x[j] = sx[lo][j];
            return sf[lo];
        }
        double cen[MCP_SLOPE_MAXD], xr[MCP_SLOPE_MAXD], xe[MCP_SLOPE_MAXD], xc[MCP_SLOPE_MAXD];
        for (int j = 0; j < D; j++) {
            double sum = 0.0;
            for (int i = 0; i < V; i++)
                if (i != hi) sum += sx[i][j];
            cen[j] = sum / (double)D;
            xr[j] = cen[j] + (cen[j] - sx[hi][j]);
        }
        const double fr = slope_dev<PT>(c, d, xr);
        if (fr < sf[lo]) {
            for (int j = 0; j < D; j++) xe[j] = cen[j] + 2.0 * (cen[j] - sx[hi][j]);
            const double fe = slope_dev<PT>(c, d, xe);
            if (fe < fr) {
                for (int j = 0; j < D; j++) sx[hi][j] = xe[j];
                sf[hi] = fe;
            } else {
                for (int j = 0; j < D; j++) sx[hi][j] = xr[j];
                sf[hi] = fr;
            }
        } else if (fr < sf[nh]) {
            for (int j = 0; j < D; j++) sx[hi][j] = xr[j];
            sf[hi] = fr;
        } else {
            const bool outside = fr < sf[hi];
            for (int j = 0; j < D; j++) xc[j] = outside ? cen[j] + 0.5 * (xr[j] - cen[j]) : cen[j] + 0.5 * (sx[hi][j] - cen[j]);
            const double fc = slope_dev<PT>(c, d, xc);
            if (fc < (outside ? fr : sf[hi])) {
                for (int j = 0; j < D; j++) sx[hi][j] = xc[j];
                sf[hi] = fc;
            } else {
                for (int i = 0; i < V; i++) {
                    if (i == lo) continue;
                    for (int j = 0; j < D; j++) sx[i][j] = sx[lo][j] + 0.5 * (sx[i][j] - sx[lo][j]);
                    sf[i] = slope_dev<PT>(c, d, sx[i]);
                }
            }
        }
    }
}

template <int PT>
static __device__ __noinline__ void lmm_slope_fit(LmmCtx<PT>& c, SlopeData<PT>& d, int& converged, int& boundary, double* theta) {
    const int q = d.q, D = q * (q + 1) / 2, max_eval = MCP_SLOPE_NM_EVAL_PER_DIM * D;
    double x[MCP_SLOPE_MAXD];
    for (int i = 0, t = 0; i < q; i++)
        for (int j = 0; j <= i; j++, t++) x[t] = i == j ? 1.0 : 0.0;
    d.n_eval = 0;
    slope_nm<PT>(c, d, x, 0.5, max_eval / 2);
    double f = slope_nm<PT>(c, d, x, 0.05, max_eval);  // restart from the best vertex
    boundary = 0;
    for (int i = 0, t = 0; i < q; i++)
        for (int j = 0; j <= i; j++, t++)
            if (i == j) {
                x[t] = fabs(x[t]);
                if (x[t] < 1e-6) boundary |= 1 | (4 << i);  // variance component i pinned at 0 (singular fit): code 1, mask bit i
            }
    f = slope_dev<PT>(c, d, x);  // leaves A, L, beta, sigma² at theta-hat
    converged = !(isinf(f) || f != f) && d.n_eval < max_eval;
    for (int t = 0; t < D; t++)
        if (fabs(x[t]) > MCP_SLOPE_THETA_MAX) {
            boundary = 2 | (boundary & ~3);
            converged = 0;
        }
    for (int j = 0; j < D; j++) theta[j] = x[j];
}

// Variance components of Lambda-hat (the oracle's slope_components): out = {Sigma_00 .., corr(1,0)[, corr(2,0), corr(2,1)]}.
static __device__ __noinline__ void slope_components(const double* th, int q, double sigma_sq, double* var_out, double* corr_out) {
    double Lm[MCP_SLOPE_MAXQ][MCP_SLOPE_MAXQ], G[MCP_SLOPE_MAXQ][MCP_SLOPE_MAXQ];
    for (int i = 0, t = 0; i < q; i++)
        for (int j = 0; j <= i; j++, t++) Lm[i][j] = th[t];
    for (int i = 0; i < q; i++)
        for (int j = 0; j <= i; j++) {
            double acc = 0.0;
            for (int m = 0; m <= j; m++) acc += Lm[i][m] * Lm[j][m];
            G[i][j] = acc;
        }
    for (int i = 0; i < q; i++) var_out[i] = sigma_sq * G[i][i];
    int o = 0;
    for (int i = 1; i < q; i++)
        for (int j = 0; j < i; j++) corr_out[o++] = G[i][j] / sqrt(G[i][i] * G[j][j]);
}

// Joint Wald chi-square of the testable targets: beta_T' [Cov_TT]^-1 beta_T, Cov = sigma² A^-1.
template <int PT>
static __device__ __noinline__ double lmm_joint_wald(LmmCtx<PT>& c, const int* cols, int k) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (k <= 0 || k > PT) return nan;
    const int p = c.p;
    double U[PT][PT], G[PT * (PT + 1) / 2], bt[PT];
    for (int t = 0; t < k; t++) {
        c.ms.t_sq(cols[t], -1);  // leaves u = L^-1 e_col in c.ms.w[col..p)
        for (int i = 0; i < p; i++) U[t][i] = i >= cols[t] ? c.ms.w[i] : 0.0;
    }
    for (int a = 0; a < k; a++)
        for (int bb = 0; bb <= a; bb++) {
            double sacc = 0.0;
            for (int i = 0; i < p; i++) sacc = fma(U[a][i], U[bb][i], sacc);
            G[tri(a, bb)] = sacc;
        }
    for (int t = 0; t < k; t++) bt[t] = c.ms.beta[cols[t]];
    const double sigma_sq = c.ms.out.sigma_sq;
    MemSolver<PT> gs;
    gs.factor_and_solve(G, bt, k, 0.0, 0.0, 2u * (uint32_t)k + 2u);
    if (!gs.out.converged) return nan;
    double q = 0.0;
    for (int t = 0; t < k; t++) q = fma(gs.w[t], gs.w[t], q);
    return q / sigma_sq;
}

// Scratch of one wave (stride = studies per wave, padded to whole CTAs; everything [..][study] so that the studies of a warp
// read and write consecutive doubles):
//   running cluster sums  csum[k][j]          k < nc_cap, j <= PT (x_0..x_{P-1}, y)
//   snapshots             snap[g][k][j]       the cluster sums after the first ss[g] rows
//   statistics            stat[g][e]          X'X packed lower | X'y | y'y after the first ss[g] rows
template <int PT>
struct LmmLayout {
    static constexpr int NACC = PT * (PT + 1) / 2;
    static constexpr int E = NACC + PT + 1;
    // per cluster: the sums of [x | y] and — with a random slope — behind them the sums of x_s · [x | y]
    static __host__ __device__ size_t cluster_doubles(int n_slopes) { return (size_t)(PT + 1) * (size_t)(1 + n_slopes); }
    static __host__ __device__ size_t csum_doubles(int nc_cap, int n_slopes = 0) { return (size_t)nc_cap * cluster_doubles(n_slopes); }
    static __host__ __device__ size_t per_study_doubles(int nc_cap, int n_ss, int n_slopes = 0) {
        return csum_doubles(nc_cap, n_slopes) * (size_t)(1 + n_ss) + (size_t)n_ss * E;
    }
};

// Stage 1 — one study per thread: generate, accumulate X'X / X'y / y'y in registers and the cluster sums in HBM; at every
// grid point leave the statistics and a snapshot of the cluster sums for the fit kernel.
template <int PT>
__global__ void __launch_bounds__(MCP_BLOCK) lmm_gen_kernel(uint64_t base_seed, uint64_t sim_begin, uint32_t n_sims,
                                                            double* __restrict__ scratch) {
    using LY = LmmLayout<PT>;
    constexpr int NACC = LY::NACC;
    const DevicePlan& pl = c_plan;
    extern __shared__ __align__(16) float smem[];
    const Slab<MCP_BLOCK> slab{smem + threadIdx.x};
    const uint32_t sim = blockIdx.x * MCP_BLOCK + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * MCP_BLOCK;
    if (sim >= n_sims || mcp_cancelled()) return;
    const uint64_t key = pcg_mix64(base_seed, sim_begin + sim);
    const int max_n = pl.max_n;
    const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
    const int nc_cap = pl.sizing_fixed_size ? (max_n + cs - 1) / cs : cs;
    const int ns = pl.n_slopes, CD = (int)LY::cluster_doubles(ns);  // doubles per cluster: [x | y] sums (+ the x_s-weighted ones)
    double* const csum = scratch + sim;
    double* const snap = csum + LY::csum_doubles(nc_cap, ns) * stride;
    double* const stat = snap + LY::csum_doubles(nc_cap, ns) * (size_t)pl.n_ss * stride;
    for (int e = 0; e < nc_cap * CD; e++) csum[(size_t)e * stride] = 0.0;

    double acc[NACC], xty[PT];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < PT; i++) xty[i] = 0.0;
    double yty = 0.0;
    ScenDesign<PT> sdv;
    const bool scen = !pl.optimistic;
    if (scen) {
        scen_populate<PT>(sdv, base_seed, sim_begin + sim);
        if (!sdv.ok) atomicExch(pl.error_flag, 2);
    }
    SimCtx cx;
    init_sim_ctx<PT>(cx, key, sim, scen ? &sdv : nullptr);
    int g = 0;
    uint32_t n_next = pl.ss[0];
    for (int b = 0; b < (max_n + 3) >> 2; b++) {
        if (scen) gen_block<PT, true>(b, slab, cx, &sdv);
        else gen_block<PT, false>(b, slab, cx, nullptr);
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
            const int row = 4 * b + r;
            if (row >= max_n) break;
            double xr[PT];
            const double yd = scen ? row_xy<PT, true>(b, r, slab, cx, xr, &sdv) : row_xy<PT, false>(b, r, slab, cx, xr, nullptr);
#pragma unroll
            for (int a = 0; a < PT; a++) {
#pragma unroll
                for (int c = 0; c <= a; c++) acc[tri(a, c)] = fma(xr[a], xr[c], acc[tri(a, c)]);
                xty[a] = fma(xr[a], yd, xty[a]);
            }
            yty = fma(yd, yd, yty);
            {  // cluster sums (plain adds in row order, like the host-side restatement)
                const int kc = pl.sizing_fixed_size ? row / cs : row % cs;
                double* base = csum + (size_t)kc * CD * stride;
#pragma unroll
                for (int j = 0; j < PT; j++) base[(size_t)j * stride] += xr[j];
                base[(size_t)PT * stride] += yd;
                for (int a = 0; a < ns; a++) {  // random slopes: sums of x_s · [x | y], one fma per term in row order (the oracle's chains)
                    double* b2 = base + (size_t)(a + 1) * (PT + 1) * stride;
                    double xs = 0.0;
#pragma unroll
                    for (int j = 0; j < PT; j++)
                        if (j == pl.slope_col[a]) xs = xr[j];
#pragma unroll
                    for (int j = 0; j < PT; j++) b2[(size_t)j * stride] = fma(xs, xr[j], b2[(size_t)j * stride]);
                    b2[(size_t)PT * stride] = fma(xs, yd, b2[(size_t)PT * stride]);
                }
            }
            if ((uint32_t)(row + 1) == n_next) {
                double* const so = stat + (size_t)g * LY::E * stride;
#pragma unroll
                for (int i = 0; i < NACC; i++) so[(size_t)i * stride] = acc[i];
#pragma unroll
                for (int i = 0; i < PT; i++) so[(size_t)(NACC + i) * stride] = xty[i];
                so[(size_t)(NACC + PT) * stride] = yty;
                const int n_clusters = pl.sizing_fixed_size ? (row + 1 + cs - 1) / cs : cs;
                double* const sn = snap + (size_t)g * LY::csum_doubles(nc_cap, ns) * stride;
                for (int e = 0; e < n_clusters * CD; e++) sn[(size_t)e * stride] = csum[(size_t)e * stride];
                g++;
                n_next = g < pl.n_ss ? pl.ss[g] : 0xffffffffu;
            }
        }
    }
}

// introspect::fit_provided_data twin, MLE arm (introspect.rs:448-530): ONE supplied dataset through the same statistics chains
// (lmm_gen_kernel) and the same profiled-REML fit (lmm_fit) as the power kernels.  Cluster ids follow the spec's sizing
// (generation.rs cluster_of_row; the host checks supplied ids against it).  plane = [row][column] f32, columns x_1 .. x_{P-1}, y;
// csum = zeroed scratch of n_clusters x (P + 1) doubles.  out: beta[P] | Var(beta)[P] | z²[P] | converged, boundary_hit, 0,
// sigma²-hat, tau²-hat = theta² sigma², profiled REML criterion at theta-hat, NaN.
template <int PT>
__global__ void lmm_fit_data_kernel(const float* __restrict__ plane, int n, double* __restrict__ csum, double* __restrict__ out) {
    constexpr int NACC = PT * (PT + 1) / 2;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const DevicePlan& pl = c_plan;
    const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
    double acc[NACC], xty[PT];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < PT; i++) xty[i] = 0.0;
    double yty = 0.0;
    for (int row = 0; row < n; row++) {
        double xr[PT];
        xr[0] = 1.0;
#pragma unroll
        for (int j = 1; j < PT; j++) xr[j] = (double)plane[(size_t)row * PT + (j - 1)];
        const double yd = (double)plane[(size_t)row * PT + (PT - 1)];
#pragma unroll
        for (int a = 0; a < PT; a++) {
#pragma unroll
            for (int c = 0; c <= a; c++) acc[tri(a, c)] = fma(xr[a], xr[c], acc[tri(a, c)]);
            xty[a] = fma(xr[a], yd, xty[a]);
        }
        yty = fma(yd, yd, yty);
        const int kc = pl.sizing_fixed_size ? row / cs : row % cs;
        double* base = csum + (size_t)kc * (PT + 1) * (1 + pl.n_slopes);
#pragma unroll
        for (int j = 0; j < PT; j++) base[j] += xr[j];
        base[PT] += yd;
        for (int a = 0; a < pl.n_slopes; a++) {
            double* b2 = base + (a + 1) * (PT + 1);
            double xs = 0.0;
#pragma unroll
            for (int j = 0; j < PT; j++)
                if (j == pl.slope_col[a]) xs = xr[j];
#pragma unroll
            for (int j = 0; j < PT; j++) b2[j] = fma(xs, xr[j], b2[j]);
            b2[PT] = fma(xs, yd, b2[PT]);
        }
    }
    LmmCtx<PT> c;
    c.p = PT;
    c.n_rows = (uint32_t)n;
    c.yty = yty;
#pragma unroll
    for (int i = 0; i < NACC; i++) c.xtx[i] = acc[i];
#pragma unroll
    for (int i = 0; i < PT; i++) c.xty[i] = xty[i];
    const int n_clusters = pl.sizing_fixed_size ? (n + cs - 1) / cs : cs;
    int conv = 0, bnd = 0;
    double theta = 0.0, th3[MCP_SLOPE_MAXD] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (pl.n_slopes > 0) {
        SlopeData<PT> sd{csum, 1, n, n_clusters, 1 + pl.n_slopes, {0, pl.slope_col[0], pl.slope_col[1]}, 0};
        lmm_slope_fit<PT>(c, sd, conv, bnd, th3);
        theta = th3[0];
    } else {
        lmm_fit<PT>(c, csum, 1, n, n_clusters, nullptr, conv, bnd, &theta);
    }
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int t = 0; t < PT; t++) {
        out[t] = c.ms.beta[t];
        out[PT + t] = conv ? c.ms.var_of(t) : nan;
        out[2 * PT + t] = conv ? c.ms.t_sq(t, -1) : nan;
    }
    out[3 * PT] = (double)conv;
    out[3 * PT + 1] = (double)bnd;
    out[3 * PT + 2] = 0.0;
    out[3 * PT + 3] = c.ms.out.sigma_sq;
    out[3 * PT + 4] = theta * theta * c.ms.out.sigma_sq;
    out[3 * PT + 5] = nan;
    out[3 * PT + 6] = nan;
    // random slopes: variances of the slope effects and the correlations (1,0) [(2,0) (2,1)] of Sigma-hat = sigma² Lambda Lambda'
    for (int i = 7; i <= 13; i++) out[3 * PT + i] = nan;
    if (pl.n_slopes > 0) {
        double vv[MCP_SLOPE_MAXQ], cr[3];
        slope_components(th3, 1 + pl.n_slopes, c.ms.out.sigma_sq, vv, cr);
        out[3 * PT + 4] = vv[0];
        for (int a = 0; a < pl.n_slopes; a++) out[3 * PT + 7 + a] = vv[1 + a];
        for (int a = 0; a < (pl.n_slopes == 2 ? 3 : 1); a++) out[3 * PT + 9 + a] = cr[a];
    }
}

// Stage 2 — one (study, sample size) fit per thread, grid = (blocks of studies, n_ss): profiled REML by Brent over the cluster
// sufficient statistics of that prefix, Wald tests, joint Wald, boundary flag, counters / byte tensors.  The per-level counts of
// the sparse-factor exclusion are the dummy column sums X'X[dummy][0] (a reference-coded dummy is exactly 0 / 1).
template <int PT>
__global__ void __launch_bounds__(MCP_BLOCK, MCP_LMM_FIT_MINB(PT)) lmm_fit_kernel(uint32_t n_sims, const double* __restrict__ scratch, size_t stride) {
    using LY = LmmLayout<PT>;
    constexpr int NACC = LY::NACC;
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
    const int g = pl.n_ss - 1 - (int)blockIdx.y;  // longest fits first
    if (sim < n_sims) {
        SimOut so{cnt_s, (uint64_t)sim};
        const int max_n = pl.max_n;
        const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
        const int nc_cap = pl.sizing_fixed_size ? (max_n + cs - 1) / cs : cs;
        const int ns = pl.n_slopes;
        const double* const snap = scratch + sim + LY::csum_doubles(nc_cap, ns) * stride + (size_t)g * LY::csum_doubles(nc_cap, ns) * stride;
        const double* const st = scratch + sim + LY::csum_doubles(nc_cap, ns) * (size_t)(1 + pl.n_ss) * stride + (size_t)g * LY::E * stride;
        const int n = (int)pl.ss[g];
        uint8_t excl[MCP_MAX_FACTORS];
        for (int f = 0; f < MCP_MAX_FACTORS; f++) excl[f] = 0;
        const bool track_excl = pl.n_factors > 0 && pl.min_level_count > 0;
        int variant = pl.base_variant;
        if (track_excl) {
            variant = 0;
            for (int f = 0; f < pl.n_factors; f++) {
                uint32_t rest = (uint32_t)n, mn = 0xffffffffu;
                for (int l = 1; l < pl.f_levels[f]; l++) {
                    const uint32_t cnt = (uint32_t)st[(size_t)tri(pl.f_col[f] + l - 1, 0) * stride];
                    mn = min(mn, cnt);
                    rest -= cnt;
                }
                mn = min(mn, rest);
                excl[f] = mn < (uint32_t)pl.min_level_count;
                variant |= (int)excl[f] << f;
            }
        }
        const McpVariant& vr = pl.variants[variant];
        const bool reduced = vr.p_red != PT;
        const int n_clusters = pl.sizing_fixed_size ? (n + cs - 1) / cs : cs;
        LmmCtx<PT> c;
        c.p = vr.p_red;
        c.n_rows = (uint32_t)n;
        c.yty = st[(size_t)(NACC + PT) * stride];
        for (int ri = 0; ri < vr.p_red; ri++) {
            c.xty[ri] = st[(size_t)(NACC + vr.kept[ri]) * stride];
            for (int rj = 0; rj <= ri; rj++) c.xtx[tri(ri, rj)] = st[(size_t)tri(vr.kept[ri], vr.kept[rj]) * stride];
        }
        int conv = 0, bnd = 0;
        if (ns > 0 && !reduced) {  // random intercept + slope: 3-parameter REML (the reduced fit strips the slope, batch.rs:1357-1375)
            SlopeData<PT> sd{snap, stride, n, n_clusters, 1 + ns, {0, pl.slope_col[0], pl.slope_col[1]}, 0};
            double theta[MCP_SLOPE_MAXD];
            lmm_slope_fit<PT>(c, sd, conv, bnd, theta);
        } else {
            lmm_fit<PT>(c, snap, stride, n, n_clusters, reduced ? vr.kept : nullptr, conv, bnd, nullptr, (int)LY::cluster_doubles(ns));
        }
        double all_t[MCP_MAX_T];
        int cols[MCP_MAX_T];
        int k_red = 0;
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int t = 0; t < pl.n_marg; t++) {
            const int col = vr.remap[pl.target[t]];
            all_t[t] = nan;
            if (col >= 0) cols[k_red++] = col;
        }
        int jsig = 0;
        if (conv) {
            for (int t = 0; t < pl.n_marg; t++) {
                const int col = vr.remap[pl.target[t]];
                if (col >= 0) all_t[t] = c.ms.t_sq(col, -1);
            }
            const double jc = pl.crit_joint[(size_t)variant * pl.n_ss + g];
            const double jt = lmm_joint_wald<PT>(c, cols, k_red);
            jsig = !(isinf(jt) || jt != jt) && jt > jc;
        }
        FitOut fo = c.ms.out;
        fo.converged = conv;
        emit_fit(so, g, variant, fo, vr.p_red, all_t, excl, /*ov_given=*/0, jsig, bnd);
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

// =====================================================================================================================
// General structure: primary grouping (random intercept + up to two slopes) plus up to two intercept-only extra groupings
// (crossed / nested).  Replaces the glmm::fit_on arm of run_one_sim (batch.rs:1214-1561, mixed_workspace.rs).  The oracle
// (its gen_dev) states the penalised-least-squares criterion on dense matrices; here the primary block stays
// block diagonal (q x q per level) and only the extras' levels are dense — every entry is the oracle's fma chain with the
// structurally-zero terms left out, so the deviances, the simplex path and the fit agree bit for bit.
//   stage 1  lmm_plane_kernel   one study per thread: the f32 data plane [row][column][study] (as the GLM arm's)
//   stage 2  lmm_general_kernel one (study, sample size) fit per thread: Z'Z / Z'[X | y] / X'X from the plane rows (one fma per
//            term, rows in order), then the search; matrices in a per-fit scratch region [entry][fit] (coalesced over a warp).
struct GenLayout {
    int K, q0, Qe, tq;
    size_t o_spp, o_sep, o_see, o_zxy, o_cp, o_l21, o_l22, o_r, total;
    __host__ __device__ GenLayout(int K_, int q0_, int Qe_, int PT) : K(K_), q0(q0_), Qe(Qe_), tq(q0_ * (q0_ + 1) / 2) {
        const size_t spp = (size_t)K * tq, sep = (size_t)Qe * K * q0, see = (size_t)Qe * (Qe + 1) / 2,
                     zxy = ((size_t)K * q0 + Qe) * (PT + 1);
        o_spp = 0;
        o_sep = o_spp + spp;
        o_see = o_sep + sep;
        o_zxy = o_see + see;
        o_cp = o_zxy + zxy;   // work: Cholesky factors of the primary blocks
        o_l21 = o_cp + spp;   //       L21 (extras x primary)
        o_l22 = o_l21 + sep;  //       L22 (extras x extras, lower)
        o_r = o_l22 + see;    //       R = L^-1 Lambda' Z'[X | y]
        total = o_r + zxy;
    }
};

template <int PT>
struct GenFit {
    double* sc;  // this fit's scratch: entry e at sc[e · stride]
    size_t stride;
    GenLayout ly;
    int q;             // random-effect components of the primary grouping in THIS fit (the reduced fit strips the slopes)
    const int* kept;   // design columns of this fit (nullptr: all)
    int n_extra, E0;
    int n_eval;
    __device__ __forceinline__ double& at(size_t e) const { return sc[e * stride]; }
};

template <int PT>
static __device__ __noinline__ double gen_dev(LmmCtx<PT>& c, GenFit<PT>& d, const double* th) {
    const GenLayout& ly = d.ly;
    const int p = c.p, q = d.q, q0 = ly.q0, K = ly.K, Qe = ly.Qe, tq = ly.tq;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double Lm[MCP_SLOPE_MAXQ][MCP_SLOPE_MAXQ], te[2];
    int t = 0;
    for (int i = 0; i < q; i++)
        for (int j = 0; j <= i; j++, t++) Lm[i][j] = i == j ? fabs(th[t]) : th[t];
    for (int g = 0; g < d.n_extra; g++) te[g] = fabs(th[t++]);
    d.n_eval++;
    double ldv = 0.0;
    // ---- primary blocks: M_k = Lambda' S_k Lambda + I, C_k = chol(M_k)
    for (int k = 0; k < K; k++) {
        double S[MCP_SLOPE_MAXQ][MCP_SLOPE_MAXQ], T[MCP_SLOPE_MAXQ][MCP_SLOPE_MAXQ], M[MCP_SLOPE_MAXQ][MCP_SLOPE_MAXQ], C[MCP_SLOPE_MAXQ][MCP_SLOPE_MAXQ];
        for (int a = 0; a < q; a++)
            for (int b = 0; b <= a; b++) S[a][b] = S[b][a] = d.at(ly.o_spp + (size_t)k * tq + tri(a, b));
        for (int r = 0; r < q; r++)
            for (int j = 0; j <= r; j++) {
                double acc = 0.0;
                for (int bb = j; bb < q; bb++) acc += Lm[bb][j] * S[r][bb];
                T[r][j] = acc;
            }
        for (int i = 0; i < q; i++)
            for (int j = 0; j <= i; j++) {
                double acc = 0.0;
                for (int aa = i; aa < q; aa++) acc += Lm[aa][i] * T[aa][j];
                M[i][j] = i == j ? acc + 1.0 : acc;
            }
        for (int j = 0; j < q; j++) {
            double dg = M[j][j];
            for (int m = 0; m < j; m++) dg = fma(-C[j][m], C[j][m], dg);
            if (!(dg > 0.0)) return inf;
            const double cjj = sqrt(dg);
            C[j][j] = cjj;
            ldv += own_log(cjj);
            for (int i = j + 1; i < q; i++) {
                double sacc = M[i][j];
                for (int m = 0; m < j; m++) sacc = fma(-C[i][m], C[j][m], sacc);
                C[i][j] = sacc / cjj;
            }
        }
        for (int a = 0; a < q; a++)
            for (int b = 0; b <= a; b++) d.at(ly.o_cp + (size_t)k * tq + tri(a, b)) = C[a][b];
        // L21 rows of this block: for every extra level e, L21[e][(k, a)] = (M[e][(k, a)] - sum_{m < a} L21[e][(k, m)] C[a][m]) / C[a][a]
        for (int e = 0; e < Qe; e++) {
            const double tg = te[e >= d.E0 ? 1 : 0];
            double w[MCP_SLOPE_MAXQ], l21[MCP_SLOPE_MAXQ];
            for (int a = 0; a < q; a++) w[a] = d.at(ly.o_sep + ((size_t)e * K + k) * q0 + a);
            for (int a = 0; a < q; a++) {
                double acc = 0.0;
                for (int bb = a; bb < q; bb++) acc += Lm[bb][a] * w[bb];
                double sacc = 0.0;
                sacc += tg * acc;
                for (int m = 0; m < a; m++) sacc = fma(-l21[m], C[a][m], sacc);
                l21[a] = sacc / C[a][a];
                d.at(ly.o_l21 + ((size_t)e * K + k) * q0 + a) = l21[a];
            }
        }
    }
    // ---- extras: the oracle's dense left-looking Cholesky continued over the extra columns (primary terms first, then extras)
    for (int j = 0; j < Qe; j++) {
        const double tj = te[j >= d.E0 ? 1 : 0];
        double dg;
        {
            double acc = 0.0;
            acc += tj * d.at(ly.o_see + tri(j, j));
            double m = 0.0;
            m += tj * acc;
            dg = m + 1.0;
        }
        for (int k = 0; k < K; k++)
            for (int a = 0; a < q; a++) {
                const double l = d.at(ly.o_l21 + ((size_t)j * K + k) * q0 + a);
                dg = fma(-l, l, dg);
            }
        for (int m = 0; m < j; m++) {
            const double l = d.at(ly.o_l22 + tri(j, m));
            dg = fma(-l, l, dg);
        }
        if (!(dg > 0.0)) return inf;
        const double ljj = sqrt(dg);
        d.at(ly.o_l22 + tri(j, j)) = ljj;
        ldv += own_log(ljj);
        for (int i = j + 1; i < Qe; i++) {
            const double ti = te[i >= d.E0 ? 1 : 0];
            double sacc;
            {
                double acc = 0.0;
                acc += tj * d.at(ly.o_see + tri(i, j));
                double m = 0.0;
                m += ti * acc;
                sacc = m;
            }
            for (int k = 0; k < K; k++)
                for (int a = 0; a < q; a++)
                    sacc = fma(-d.at(ly.o_l21 + ((size_t)i * K + k) * q0 + a), d.at(ly.o_l21 + ((size_t)j * K + k) * q0 + a), sacc);
            for (int m = 0; m < j; m++) sacc = fma(-d.at(ly.o_l22 + tri(i, m)), d.at(ly.o_l22 + tri(j, m)), sacc);
            d.at(ly.o_l22 + tri(i, j)) = sacc / ljj;
        }
    }
    // ---- R = L^-1 Lambda' Z'[X | y], column by column
    for (int cc = 0; cc <= p; cc++) {
        const int col = cc == p ? PT : (d.kept ? d.kept[cc] : cc);
        for (int k = 0; k < K; k++) {
            double rr[MCP_SLOPE_MAXQ];
            for (int a = 0; a < q; a++) {
                double u = 0.0;
                for (int aa = a; aa < q; aa++) u += Lm[aa][a] * d.at(ly.o_zxy + ((size_t)k * q0 + aa) * (PT + 1) + col);
                for (int m = 0; m < a; m++) u = fma(-d.at(ly.o_cp + (size_t)k * tq + tri(a, m)), rr[m], u);
                rr[a] = u / d.at(ly.o_cp + (size_t)k * tq + tri(a, a));
                d.at(ly.o_r + ((size_t)k * q0 + a) * (PT + 1) + cc) = rr[a];
            }
        }
        for (int e = 0; e < Qe; e++) {
            double u = 0.0;
            u += te[e >= d.E0 ? 1 : 0] * d.at(ly.o_zxy + ((size_t)K * q0 + e) * (PT + 1) + col);
            for (int k = 0; k < K; k++)
                for (int a = 0; a < q; a++)
                    u = fma(-d.at(ly.o_l21 + ((size_t)e * K + k) * q0 + a), d.at(ly.o_r + ((size_t)k * q0 + a) * (PT + 1) + cc), u);
            for (int m = 0; m < e; m++) u = fma(-d.at(ly.o_l22 + tri(e, m)), d.at(ly.o_r + ((size_t)K * q0 + m) * (PT + 1) + cc), u);
            d.at(ly.o_r + ((size_t)K * q0 + e) * (PT + 1) + cc) = u / d.at(ly.o_l22 + tri(e, e));
        }
    }
    // ---- A = X'X - R_x' R_x, b = X'y - R_x' R_y, c = y'y - |R_y|²: every random-effect column in order
    const int nt = p * (p + 1) / 2;
    for (int e = 0; e < nt; e++) c.A[e] = c.xtx[e];
    for (int j = 0; j < p; j++) c.b[j] = c.xty[j];
    double cc0 = c.yty;
    for (int i = 0; i < K * q + Qe; i++) {
        const size_t row = i < K * q ? (size_t)(i / q) * q0 + (size_t)(i % q) : (size_t)K * q0 + (size_t)(i - K * q);
        double ri[PT + 1];
        for (int j = 0; j <= p; j++) ri[j] = d.at(ly.o_r + row * (PT + 1) + j);
        int e = 0;
        for (int a = 0; a < p; a++) {
            for (int bb = 0; bb <= a; bb++, e++) c.A[e] = fma(-ri[a], ri[bb], c.A[e]);
            c.b[a] = fma(-ri[a], ri[p], c.b[a]);
        }
        cc0 = fma(-ri[p], ri[p], cc0);
    }
    c.ms.factor_and_solve(c.A, c.b, p, cc0, 0.0, c.n_rows);
    if (!c.ms.out.converged || !(c.ms.out.rss > 0.0) || isinf(c.ms.out.rss) || isinf(ldv) || ldv != ldv) return inf;
    double lda = 0.0;
    for (int j = 0; j < p; j++) lda += own_log(c.ms.L[tri(j, j)]);
    return 2.0 * ldv + 2.0 * lda + c.ms.out.df_resid * own_log(c.ms.out.sigma_sq);
}

// The oracle's nm_minimise over gen_dev (same comparisons in the same order as slope_nm above).
template <int PT>
static __device__ __noinline__ double gen_nm(LmmCtx<PT>& c, GenFit<PT>& d, int D, double* x, double step, int max_eval) {
    const int V = D + 1;
    double sx[MCP_SLOPE_MAXD + 1][MCP_SLOPE_MAXD], sf[MCP_SLOPE_MAXD + 1];
    for (int i = 0; i < V; i++) {
        for (int j = 0; j < D; j++) sx[i][j] = x[j] + (i == j + 1 ? step : 0.0);
        sf[i] = gen_dev<PT>(c, d, sx[i]);
    }
    for (;;) {
        int lo = 0, hi = 0, nh;
        for (int i = 1; i < V; i++) {
            if (sf[i] < sf[lo]) lo = i;
            if (sf[i] > sf[hi] || !(sf[i] == sf[i])) hi = i;
        }
        nh = lo;
        for (int i = 0; i < V; i++)
            if (i != hi && (sf[i] > sf[nh])) nh = i;
        double size = 0.0;
        for (int i = 0; i < V; i++)
            for (int j = 0; j < D; j++) {
                const double dd = fabs(sx[i][j] - sx[lo][j]);
                if (dd > size) size = dd;
            }
        // (a cancelled batch is discarded: leaving the search early only shortens the wait — these fits run for milliseconds)
        if ((fabs(sf[hi] - sf[lo]) <= 1e-11 * (fabs(sf[lo]) + 1.0) && size <= 1e-7) || d.n_eval >= max_eval || mcp_cancelled()) {
            for (int j = 0; j < D; j++) x[j] = sx[lo][j];
            return sf[lo];
        }
        double cen[MCP_SLOPE_MAXD], xr[MCP_SLOPE_MAXD], xe[MCP_SLOPE_MAXD], xc[MCP_SLOPE_MAXD];
        for (int j = 0; j < D; j++) {
            double sum = 0.0;
            for (int i = 0; i < V; i++)
                if (i != hi) sum += sx[i][j];
            cen[j] = sum / (double)D;
            xr[j] = cen[j] + (cen[j] - sx[hi][j]);
        }
        const double fr = gen_dev<PT>(c, d, xr);
        if (fr < sf[lo]) {
            for (int j = 0; j < D; j++) xe[j] = cen[j] + 2.0 * (cen[j] - sx[hi][j]);
            const double fe = gen_dev<PT>(c, d, xe);
            if (fe < fr) {
                for (int j = 0; j < D; j++) sx[hi][j] = xe[j];
                sf[hi] = fe;
            } else {
                for (int j = 0; j < D; j++) sx[hi][j] = xr[j];
                sf[hi] = fr;
            }
        } else if (fr < sf[nh]) {
            for (int j = 0; j < D; j++) sx[hi][j] = xr[j];
            sf[hi] = fr;
        } else {
            const bool outside = fr < sf[hi];
            for (int j = 0; j < D; j++) xc[j] = outside ? cen[j] + 0.5 * (xr[j] - cen[j]) : cen[j] + 0.5 * (sx[hi][j] - cen[j]);
            const double fc = gen_dev<PT>(c, d, xc);
            if (fc < (outside ? fr : sf[hi])) {
                for (int j = 0; j < D; j++) sx[hi][j] = xc[j];
                sf[hi] = fc;
            } else {
                for (int i = 0; i < V; i++) {
                    if (i == lo) continue;
                    for (int j = 0; j < D; j++) sx[i][j] = sx[lo][j] + 0.5 * (sx[i][j] - sx[lo][j]);
                    sf[i] = gen_dev<PT>(c, d, sx[i]);
                }
            }
        }
    }
}

template <int PT>
static __device__ __noinline__ void lmm_general_fit(LmmCtx<PT>& c, GenFit<PT>& d, int& converged, int& boundary, double* theta_out = nullptr) {
    const int q = d.q, D = q * (q + 1) / 2 + d.n_extra, max_eval = MCP_SLOPE_NM_EVAL_PER_DIM * D;
    double x[MCP_SLOPE_MAXD];
    int t = 0;
    for (int i = 0; i < q; i++)
        for (int j = 0; j <= i; j++, t++) x[t] = i == j ? 1.0 : 0.0;
    for (int g = 0; g < d.n_extra; g++) x[t++] = 1.0;
    d.n_eval = 0;
    gen_nm<PT>(c, d, D, x, 0.5, max_eval / 2);
    double f = gen_nm<PT>(c, d, D, x, 0.05, max_eval);
    boundary = 0;
    t = 0;
    for (int i = 0; i < q; i++)
        for (int j = 0; j <= i; j++, t++)
            if (i == j) {
                x[t] = fabs(x[t]);
                if (x[t] < 1e-6) boundary |= 1 | (4 << i);
            }
    for (int g = 0; g < d.n_extra; g++, t++) {
        x[t] = fabs(x[t]);
        if (x[t] < 1e-6) boundary |= 1 | (4 << (q + g));  // components of the fitted model: primary diagonal, then the extras
    }
    f = gen_dev<PT>(c, d, x);
    converged = !(isinf(f) || f != f) && d.n_eval < max_eval;
    for (int j = 0; j < D; j++)
        if (fabs(x[j]) > MCP_SLOPE_THETA_MAX) {
            boundary = 2 | (boundary & ~3);
            converged = 0;
        }
    if (theta_out)
        for (int j = 0; j < D; j++) theta_out[j] = x[j];
}

// Z'Z / Z'[X | y] (scratch, zeroed here) and X'X / X'y / y'y (registers) of rows [0, n) of one study's plane: one fma per term, rows
// in order — the oracle's gen_add_row / suff_add_rows.
template <int PT>
__device__ __forceinline__ void gen_accumulate(const float* __restrict__ pp, size_t pstride, int n, double* __restrict__ sc, size_t stride,
                                               const GenLayout& ly, int E0, double (&acc)[PT * (PT + 1) / 2], double (&xty)[PT], double& yty) {
    constexpr int NACC = PT * (PT + 1) / 2;
    const DevicePlan& pl = c_plan;
    const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
    const int K = ly.K, q0 = ly.q0;
    for (size_t e = 0; e < ly.o_cp; e++) sc[e * stride] = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < PT; i++) xty[i] = 0.0;
    yty = 0.0;
    for (int row = 0; row < n; row++) {
        double xy[PT + 1];
        xy[0] = 1.0;
#pragma unroll
        for (int j = 1; j <= PT; j++) xy[j] = (double)pp[((size_t)row * PT + (j - 1)) * pstride];
#pragma unroll
        for (int a = 0; a < PT; a++) {
#pragma unroll
            for (int b = 0; b <= a; b++) acc[tri(a, b)] = fma(xy[a], xy[b], acc[tri(a, b)]);
            xty[a] = fma(xy[a], xy[PT], xty[a]);
        }
        yty = fma(xy[PT], xy[PT], yty);
        const int k = pl.sizing_fixed_size ? row / cs : row % cs;
        double z[MCP_SLOPE_MAXQ];
        z[0] = 1.0;
        for (int a = 1; a < q0; a++) {
            double xs = 0.0;
#pragma unroll
            for (int j = 0; j < PT; j++)
                if (j == pl.slope_col[a - 1]) xs = xy[j];
            z[a] = xs;
        }
        int ecol[2];
        for (int e = 0; e < pl.n_extra; e++) ecol[e] = (e ? E0 : 0) + extra_level_of_row(pl, e, row);
        // the row's non-zero random-effect columns in order: primary components, then one column per extra grouping
        for (int a = 0; a < q0; a++) {
            for (int b = 0; b <= a; b++) {
                double& v = sc[(ly.o_spp + (size_t)k * ly.tq + tri(a, b)) * stride];
                v = fma(z[a], z[b], v);
            }
            double* zr = sc + (ly.o_zxy + ((size_t)k * q0 + a) * (PT + 1)) * stride;
#pragma unroll
            for (int j = 0; j <= PT; j++) zr[(size_t)j * stride] = fma(z[a], xy[j], zr[(size_t)j * stride]);
        }
        for (int e = 0; e < pl.n_extra; e++) {
            for (int a = 0; a < q0; a++) {
                double& v = sc[(ly.o_sep + ((size_t)ecol[e] * K + k) * q0 + a) * stride];
                v = fma(1.0, z[a], v);
            }
            for (int e2 = 0; e2 <= e; e2++) {
                double& v = sc[(ly.o_see + tri(ecol[e], ecol[e2])) * stride];
                v = fma(1.0, 1.0, v);
            }
            double* zr = sc + (ly.o_zxy + ((size_t)K * q0 + ecol[e]) * (PT + 1)) * stride;
#pragma unroll
            for (int j = 0; j <= PT; j++) zr[(size_t)j * stride] = fma(1.0, xy[j], zr[(size_t)j * stride]);
        }
    }
}

// fit_provided_data twin for a spec with extra groupings (introspect.rs:174-562): plane = [row][x_1 .. x_{P-1}, y] f32, scratch =
// GenLayout::total doubles.  out as lmm_fit_data_kernel's, plus the extra groupings' variances at [3P + 12], [3P + 13].
template <int PT>
__global__ void lmm_general_fit_data_kernel(const float* __restrict__ plane, int n, double* __restrict__ scratch, double* __restrict__ out) {
    constexpr int NACC = PT * (PT + 1) / 2;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const DevicePlan& pl = c_plan;
    const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
    const int K = pl.sizing_fixed_size ? (pl.max_n + cs - 1) / cs : cs;
    const int q0 = 1 + pl.n_slopes, E0 = pl.extra_levels[0], Qe = pl.extra_levels[0] + pl.extra_levels[1];
    const GenLayout ly(K, q0, Qe, PT);
    LmmCtx<PT> c;
    double acc[NACC], xty[PT], yty;
    gen_accumulate<PT>(plane, 1, n, scratch, 1, ly, E0, acc, xty, yty);
    c.p = PT;
    c.n_rows = (uint32_t)n;
    c.yty = yty;
#pragma unroll
    for (int i = 0; i < NACC; i++) c.xtx[i] = acc[i];
#pragma unroll
    for (int i = 0; i < PT; i++) c.xty[i] = xty[i];
    GenFit<PT> gd{scratch, 1, ly, q0, nullptr, pl.n_extra, E0, 0};
    int conv = 0, bnd = 0;
    double th[MCP_SLOPE_MAXD] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    lmm_general_fit<PT>(c, gd, conv, bnd, th);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int t = 0; t < PT; t++) {
        out[t] = c.ms.beta[t];
        out[PT + t] = conv ? c.ms.var_of(t) : nan;
        out[2 * PT + t] = conv ? c.ms.t_sq(t, -1) : nan;
    }
    out[3 * PT] = (double)conv;
    out[3 * PT + 1] = (double)bnd;
    out[3 * PT + 2] = 0.0;
    out[3 * PT + 3] = c.ms.out.sigma_sq;
    out[3 * PT + 5] = nan;
    out[3 * PT + 6] = nan;
    for (int i = 7; i <= 13; i++) out[3 * PT + i] = nan;
    double vv[MCP_SLOPE_MAXQ], cr[3];
    slope_components(th, q0, c.ms.out.sigma_sq, vv, cr);
    out[3 * PT + 4] = vv[0];
    for (int a = 0; a < pl.n_slopes; a++) out[3 * PT + 7 + a] = vv[1 + a];
    for (int a = 0; a < (pl.n_slopes == 2 ? 3 : (pl.n_slopes == 1 ? 1 : 0)); a++) out[3 * PT + 9 + a] = cr[a];
    const int D0 = q0 * (q0 + 1) / 2;
    for (int g = 0; g < pl.n_extra; g++) out[3 * PT + 12 + g] = c.ms.out.sigma_sq * (th[D0 + g] * th[D0 + g]);
}

// Stage 1 — one study per thread: generate the rows and park them as f32, plane[(row · PT + column) · stride + study]
// (columns x_1 .. x_{P-1}, y), like the GLM arm's plane.
template <int PT>
__global__ void __launch_bounds__(MCP_BLOCK) lmm_plane_kernel(uint64_t base_seed, uint64_t sim_begin, uint32_t n_sims, float* __restrict__ plane) {
    const DevicePlan& pl = c_plan;
    extern __shared__ __align__(16) float smem[];
    const Slab<MCP_BLOCK> slab{smem + threadIdx.x};
    const uint32_t sim = blockIdx.x * MCP_BLOCK + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * MCP_BLOCK;
    if (sim >= n_sims || mcp_cancelled()) return;
    const uint64_t key = pcg_mix64(base_seed, sim_begin + sim);
    ScenDesign<PT> sdv;
    const bool scen = !pl.optimistic;
    if (scen) {
        scen_populate<PT>(sdv, base_seed, sim_begin + sim);
        if (!sdv.ok) atomicExch(pl.error_flag, 2);
    }
    SimCtx cx;
    init_sim_ctx<PT>(cx, key, sim, scen ? &sdv : nullptr);
    for (int b = 0; b < (pl.max_n + 3) >> 2; b++) {
        if (scen) gen_block<PT, true>(b, slab, cx, &sdv);
        else gen_block<PT, false>(b, slab, cx, nullptr);
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
            const int row = 4 * b + r;
            if (row >= pl.max_n) break;
            double xr[PT];
            const double yd = scen ? row_xy<PT, true>(b, r, slab, cx, xr, &sdv) : row_xy<PT, false>(b, r, slab, cx, xr, nullptr);
            float* const rowp = plane + (size_t)row * PT * stride + sim;
#pragma unroll
            for (int j = 1; j < PT; j++) rowp[(size_t)(j - 1) * stride] = (float)xr[j];
            rowp[(size_t)(PT - 1) * stride] = (float)yd;
        }
    }
}

// Stage 2 — one (study, sample size) fit per thread.
template <int PT>
__global__ void __launch_bounds__(MCP_BLOCK) lmm_general_kernel(uint32_t n_sims, const float* __restrict__ plane, double* __restrict__ fit_scratch,
                                                                size_t stride) {
    constexpr int NACC = PT * (PT + 1) / 2;
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
    const int g = pl.n_ss - 1 - (int)blockIdx.y;  // longest fits first
    if (sim < n_sims) {
        SimOut so{cnt_s, (uint64_t)sim};
        const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
        const int K = pl.sizing_fixed_size ? (pl.max_n + cs - 1) / cs : cs;  // capacity layout: levels without a row stay zero columns
        const int q0 = 1 + pl.n_slopes, E0 = pl.extra_levels[0], Qe = pl.extra_levels[0] + pl.extra_levels[1];
        const GenLayout ly(K, q0, Qe, PT);
        double* const sc = fit_scratch + (size_t)g * ly.total * stride + sim;
        const int n = (int)pl.ss[g];
        LmmCtx<PT> c;
        double acc[NACC], xty[PT], yty;
        gen_accumulate<PT>(plane + sim, stride, n, sc, stride, ly, E0, acc, xty, yty);
        uint8_t excl[MCP_MAX_FACTORS];
        for (int f = 0; f < MCP_MAX_FACTORS; f++) excl[f] = 0;
        const bool track_excl = pl.n_factors > 0 && pl.min_level_count > 0;
        int variant = pl.base_variant;
        if (track_excl) {  // a reference-coded dummy is exactly 0 / 1: its column sum X'X[dummy][0] is the level count
            variant = 0;
            for (int f = 0; f < pl.n_factors; f++) {
                uint32_t rest = (uint32_t)n, mn = 0xffffffffu;
                for (int l = 1; l < pl.f_levels[f]; l++) {
                    uint32_t cnt = 0;
#pragma unroll
                    for (int j = 0; j < PT; j++)
                        if (j == pl.f_col[f] + l - 1) cnt = (uint32_t)acc[tri(j, 0)];
                    mn = min(mn, cnt);
                    rest -= cnt;
                }
                mn = min(mn, rest);
                excl[f] = mn < (uint32_t)pl.min_level_count;
                variant |= (int)excl[f] << f;
            }
        }
        const McpVariant& vr = pl.variants[variant];
        const bool reduced = vr.p_red != PT;
        c.p = vr.p_red;
        c.n_rows = (uint32_t)n;
        c.yty = yty;
        {
            double full[NACC], fxty[PT];
#pragma unroll
            for (int i = 0; i < NACC; i++) full[i] = acc[i];
#pragma unroll
            for (int i = 0; i < PT; i++) fxty[i] = xty[i];
            for (int ri = 0; ri < vr.p_red; ri++) {
                c.xty[ri] = fxty[vr.kept[ri]];
                for (int rj = 0; rj <= ri; rj++) c.xtx[tri(ri, rj)] = full[tri(vr.kept[ri], vr.kept[rj])];
            }
        }
        int kept_i[PT];
        for (int ri = 0; ri < vr.p_red; ri++) kept_i[ri] = vr.kept[ri];
        GenFit<PT> gd{sc, stride, ly, reduced ? 1 : q0, reduced ? kept_i : nullptr, pl.n_extra, E0, 0};
        int conv = 0, bnd = 0;
        lmm_general_fit<PT>(c, gd, conv, bnd);
        double all_t[MCP_MAX_T];
        int cols[MCP_MAX_T];
        int k_red = 0;
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int t = 0; t < pl.n_marg; t++) {
            const int col = vr.remap[pl.target[t]];
            all_t[t] = nan;
            if (col >= 0) cols[k_red++] = col;
        }
        int jsig = 0;
        if (conv) {
            for (int t = 0; t < pl.n_marg; t++) {
                const int col = vr.remap[pl.target[t]];
                if (col >= 0) all_t[t] = c.ms.t_sq(col, -1);
            }
            const double jc = pl.crit_joint[(size_t)variant * pl.n_ss + g];
            const double jt = lmm_joint_wald<PT>(c, cols, k_red);
            jsig = !(isinf(jt) || jt != jt) && jt > jc;
        }
        FitOut fo = c.ms.out;
        fo.converged = conv;
        emit_fit(so, g, variant, fo, vr.p_red, all_t, excl, /*ov_given=*/0, jsig, bnd);
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

}  // namespace mcp
