// api.cu — host side of libmcpower_b200.so: spec lowering, launch, and the C ABI declared in
// include/mcpower_b200.h.  There is no CPU execution path here: a missing device or a failed
// launch is reported as MCP_STATUS_CUDA_ERROR, never papered over.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mcpower_b200.h"
#include "contract.hpp"
#include "critvals.hpp"
#include "device_math.cuh"
#include "launch.h"
#include "orchestrator.hpp"
#include "pipe_bench.cuh"
#include "spec.hpp"

namespace mcp {

static thread_local std::string g_last_error;
static std::atomic<int> g_force_ols_variant{0};  // test hook (mcp_test_force_ols_variant): 0 = the product's own choice

#define CUDA_CHECK(expr)                                                                                      \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            throw EngineError(MCP_CUDA_ERROR, std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " #expr); \
    } while (0)

// ---- pinned host memory for everything the device writes back ------------------------------------------------------
// A device→host copy into pageable memory is synchronous and staged through the driver's bounce buffers (~10 GB/s); the
// BatchResult byte tensors of a large call are gigabytes.  Result buffers therefore come from a process-wide pool of
// page-locked, portable blocks (cudaHostAlloc costs ~0.3 ms per MiB the first time, nothing when a freed block is reused);
// mcp_free_result hands them back.  At most kKeepBytes stay cached.
struct PinnedPool {
    struct Blk { void* p; size_t bytes; };
    std::mutex mu;
    std::vector<Blk> idle;
    std::map<void*, size_t> live;
    size_t cached = 0;
    static constexpr size_t kKeepBytes = (size_t)8 << 30, kGrain = (size_t)64 << 10;
    void* take(size_t need) {
        const size_t bytes = ((std::max<size_t>(need, 1) + kGrain - 1) / kGrain) * kGrain;
        {
            std::lock_guard<std::mutex> g(mu);
            size_t best = idle.size();
            for (size_t i = 0; i < idle.size(); i++)
                if (idle[i].bytes >= bytes && idle[i].bytes <= 2 * bytes && (best == idle.size() || idle[i].bytes < idle[best].bytes)) best = i;
            if (best != idle.size()) {
                Blk b = idle[best];
                idle.erase(idle.begin() + (long)best);
                cached -= b.bytes;
                live[b.p] = b.bytes;
                return b.p;
            }
        }
        void* p = nullptr;
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
        if (e != cudaSuccess) {  // out of lockable memory: drop the cache and retry once
            trim(0);
            cudaGetLastError();
            e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
        }
        if (e != cudaSuccess) throw EngineError(MCP_CUDA_ERROR, std::string("CUDA error: ") + cudaGetErrorString(e) + " at cudaHostAlloc(result buffer)");
        std::lock_guard<std::mutex> g(mu);
        live[p] = bytes;
        return p;
    }
    bool give(void* p) {  // false: not one of ours (plain malloc)
        Blk b{p, 0};
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = live.find(p);
            if (it == live.end()) return false;
            b.bytes = it->second;
            live.erase(it);
            if (cached + b.bytes <= kKeepBytes) {
                idle.push_back(b);
                cached += b.bytes;
                return true;
            }
        }
        cudaFreeHost(p);
        return true;
    }
    void trim(size_t keep) {
        std::vector<Blk> drop;
        {
            std::lock_guard<std::mutex> g(mu);
            while (cached > keep && !idle.empty()) {
                drop.push_back(idle.back());
                cached -= idle.back().bytes;
                idle.pop_back();
            }
        }
        for (auto& b : drop) cudaFreeHost(b.p);
    }
};
static PinnedPool g_pinned;

// ---- one context per CUDA device ------------------------------------------------------------------------------------
// c_plan is a __constant__ symbol, i.e. one copy per device: batches on the same device are serialised by the context's
// mutex, batches on different devices run concurrently (also from one call: n_devices > 1).  The context owns two
// non-blocking streams (kernels; device→host copies of finished waves), a control stream for the cancel flag, grow-only
// device arenas (cudaMalloc / cudaFree of gigabytes cost milliseconds and synchronise the device) and a small page-locked
// mailbox: progress word written by the kernels, staging for the counters.
struct Arena {
    void* p = nullptr;
    size_t bytes = 0;
    void* get(size_t need) {
        if (bytes < need) {
            if (p) cudaFree(p);
            p = nullptr;
            bytes = 0;
            CUDA_CHECK(cudaMalloc(&p, need));
            bytes = need;
        }
        return p;
    }
};

struct Mailbox {
    volatile unsigned int progress;  // written by the kernels (mapped), read by the polling host thread
    int one;                         // source of the cancel-flag copy
    int error_flag;                  // device error flag, copied back with the counters
    int pad;
};

struct DeviceCtx {
    int dev = -1;
    std::mutex mu;
    cudaStream_t compute = nullptr, copy = nullptr, ctl = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, done_compute = nullptr, done_copy = nullptr;
    std::vector<cudaEvent_t> wave_events;
    Arena scratch, tables, outputs, small, between, staging;
    Mailbox* mail = nullptr;            // page-locked, mapped
    unsigned long long* h_counters = nullptr;  // page-locked staging, grow-only
    size_t h_counters_cap = 0;
    unsigned int* d_prog = nullptr;     // device progress counter
    int* d_cancel = nullptr;            // device cancel flag
    int* d_error = nullptr;             // device error flag (2 = CorrelationNotPSD raised by a study)
    cudaEvent_t wave_event(size_t i) {
        while (wave_events.size() <= i) {
            cudaEvent_t e;
            CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            wave_events.push_back(e);
        }
        return wave_events[i];
    }
    unsigned long long* counters_staging(size_t n) {
        if (h_counters_cap < n) {
            if (h_counters) cudaFreeHost(h_counters);
            h_counters = nullptr;
            h_counters_cap = 0;
            CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&h_counters), std::max<size_t>(n, 4096) * sizeof(unsigned long long), cudaHostAllocPortable));
            h_counters_cap = std::max<size_t>(n, 4096);
        }
        return h_counters;
    }
};

static std::mutex g_ctx_mu;
static DeviceCtx* g_ctx[64] = {};

static int device_count_checked() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        throw EngineError(MCP_CUDA_ERROR, "no CUDA device available — the CUDA backend has no CPU fallback");
    return n;
}

// The context of device `dev` (−1: the calling thread's current device), created on first use.  Leaves `dev` current.
static DeviceCtx& ctx_for(int dev) {
    const int n = device_count_checked();
    if (dev < 0) CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= n || dev >= 64) throw EngineError(MCP_CUDA_ERROR, "CUDA device ordinal " + std::to_string(dev) + " out of range (" + std::to_string(n) + " visible)");
    CUDA_CHECK(cudaSetDevice(dev));
    std::lock_guard<std::mutex> g(g_ctx_mu);
    if (!g_ctx[dev]) {
        std::unique_ptr<DeviceCtx> cx(new DeviceCtx);
        cx->dev = dev;
        CUDA_CHECK(cudaStreamCreateWithFlags(&cx->compute, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&cx->copy, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&cx->ctl, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreate(&cx->e0));
        CUDA_CHECK(cudaEventCreate(&cx->e1));
        CUDA_CHECK(cudaEventCreateWithFlags(&cx->done_compute, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&cx->done_copy, cudaEventDisableTiming));
        CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&cx->mail), sizeof(Mailbox), cudaHostAllocPortable | cudaHostAllocMapped));
        std::memset(cx->mail, 0, sizeof(Mailbox));
        cx->mail->one = 1;
        void* d = nullptr;
        CUDA_CHECK(cudaMalloc(&d, 256));
        CUDA_CHECK(cudaMemset(d, 0, 256));
        cx->d_prog = static_cast<unsigned int*>(d);
        cx->d_cancel = reinterpret_cast<int*>(static_cast<char*>(d) + 64);
        cx->d_error = reinterpret_cast<int*>(static_cast<char*>(d) + 128);
        g_ctx[dev] = cx.release();
    }
    return *g_ctx[dev];
}

// Device pointer of the mapped progress word (UVA: identical to the host pointer on every 64-bit platform CUDA supports;
// asked for anyway).
static volatile unsigned int* mapped_progress_word(DeviceCtx& cx) {
    void* d = nullptr;
    CUDA_CHECK(cudaHostGetDevicePointer(&d, cx.mail, 0));
    return &static_cast<Mailbox*>(d)->progress;
}

// Carves sub-buffers out of one arena block (256-byte aligned).
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* b) : base(static_cast<char*>(b)) {}
    static size_t pad(size_t n) { return (std::max<size_t>(n, 1) + 255) & ~(size_t)255; }  // what take() advances by: never 0
    template <typename T>
    T* take(size_t n) {
        T* p = reinterpret_cast<T*>(base + off);
        off += pad(n * sizeof(T));
        return p;
    }
};

// Host twin of device own_exp (device_math.cuh): same IEEE operations, so constants folded on the host
// (1/exp(gamma²/2)) carry the bits the kernel would compute.
static double host_own_exp(double x) {
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.1332191019412) return 0.0;
    const double kf = std::rint(x * 1.4426950408889634);
    double r = std::fma(-kf, 6.93147180369123816490e-01, x);
    r = std::fma(-kf, 1.90821492927058770002e-10, r);
    static const double inv_fact[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
                                        1.0 / 5040.0,      1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,     1.0 / 6.0,
                                        0.5,               1.0,              1.0};
    double p = 1.0 / 6227020800.0;
    for (double c : inv_fact) p = std::fma(p, r, c);
    const int k = (int)kf;
    const int k1 = k / 2, k2 = k - k1;
    uint64_t b1 = (uint64_t)(k1 + 1023) << 52, b2 = (uint64_t)(k2 + 1023) << 52;
    double s1, s2;
    std::memcpy(&s1, &b1, 8);
    std::memcpy(&s2, &b2, 8);
    return p * s1 * s2;
}

// ---- exclusion variants (batch.rs:2502-2562) --------------------------------------------------
static McpVariant make_variant(const Spec& s, int mask) {
    McpVariant v;
    std::memset(&v, 0, sizeof v);
    const int base = 1 + s.k, nfd = s.n_fd, P = s.P();
    std::vector<char> keep_dummy(nfd, 1);
    int col = base;
    for (size_t f = 0; f < s.factor_n_levels.size(); f++) {
        const int width = s.factor_n_levels[f] - 1;
        if (mask & (1 << f))
            for (int d = 0; d < width; d++) keep_dummy[col + d - base] = 0;
        col += width;
    }
    for (int c = 0; c < MCP_MAX_P; c++) v.remap[c] = -1;
    int nk = 0;
    for (int c = 0; c < P; c++) {
        bool keep = true;
        if (c >= base && c < base + nfd) keep = keep_dummy[c - base];
        else if (c >= base + nfd)
            for (int cu : s.interactions[c - base - nfd])
                if (cu >= base && cu < base + nfd && !keep_dummy[cu - base]) keep = false;
        if (keep && !s.fit_columns.empty()) keep = std::find(s.fit_columns.begin(), s.fit_columns.end(), c) != s.fit_columns.end();
        if (keep) {
            v.remap[c] = (int8_t)nk;
            v.kept[nk++] = (int8_t)c;
        }
    }
    v.p_red = nk;
    v.valid = 1;
    return v;
}

static double tukey_k_for_col(const Spec& s, int col) {  // batch.rs:2233-2250
    const int base = 1 + s.k;
    if (col < base) return crit::kNaN;
    int cursor = base;
    for (int l : s.factor_n_levels) {
        if (col < cursor + l - 1) return (double)l;
        cursor += l - 1;
    }
    return crit::kNaN;
}

// One CritValueTable row (critvals.rs:872-960 + overall 979-1003) for sample size n and `n_predictors` fitted columns.
static void crit_row(const Spec& s, uint32_t n, int n_predictors, double* t_unc_sq, double* corr /*T*/, double* posthoc_sq,
                     double* overall) {
    const long df_signed = (long)n - (long)n_predictors;
    if (df_signed < 1) invalid("df = N - P = " + std::to_string(n) + " - " + std::to_string(n_predictors) + " < 1");
    const double df = (double)df_signed, alpha = s.alpha;
    const bool use_t = s.estimator == E_OLS;  // EstimatorSpec::uses_student_t
    auto quant = [&](double p) { return use_t ? crit::t_ppf(p, df) : crit::norm_ppf(p); };
    const int m = s.T();
    const double t_unc = quant(1.0 - alpha / 2.0);
    if (t_unc_sq) *t_unc_sq = t_unc * t_unc;
    if (corr) {
        switch (s.correction) {
            case C_NONE:
                for (int i = 0; i < m; i++) corr[i] = t_unc * t_unc;
                break;
            case C_BONFERRONI: {
                const double b = quant(1.0 - alpha / (2.0 * (double)m));
                for (int i = 0; i < m; i++) corr[i] = b * b;
                break;
            }
            case C_HOLM:
                for (int i = 0; i < m; i++) {
                    const double eff = alpha / (2.0 * (double)(m - i));
                    const double t = eff < 1e-12 ? crit::kInf : quant(1.0 - eff);
                    corr[i] = t * t;
                }
                break;
            case C_BH:
                for (int i = 0; i < m; i++) {
                    const double eff = (double)(i + 1) / (double)m * alpha / 2.0;
                    const double t = eff < 1e-12 ? crit::kInf : quant(1.0 - eff);
                    corr[i] = t * t;
                }
                break;
            case C_TUKEY: {  // the studentized-range quantile costs ~0.2 s: one evaluation per distinct k in the row
                double qk[MCP_MAX_LEVELS + 2];
                for (double& v : qk) v = -1.0;
                for (int i = 0; i < m; i++) {
                    const int nm = (int)s.target_indices.size();
                    const int col = i < nm ? s.target_indices[i] : s.contrast_pairs[i - nm].first;
                    const double kk = tukey_k_for_col(s, col);
                    double q;
                    if (kk == kk && kk >= 0.0 && kk <= MCP_MAX_LEVELS) {
                        double& slot = qk[(int)kk];
                        if (slot == -1.0) slot = crit::q_tukey_ppf(1.0 - alpha, kk, df);
                        q = slot;
                    } else {
                        q = crit::q_tukey_ppf(1.0 - alpha, kk, df);
                    }
                    const double t = q / crit::kSqrt2;
                    corr[i] = t * t;
                }
                break;
            }
        }
    }
    if (posthoc_sq) {
        const double pa = std::isnan(s.posthoc_alpha) ? alpha : s.posthoc_alpha;
        const double tp = quant(1.0 - pa / 2.0);
        *posthoc_sq = tp * tp;
    }
    if (overall) {
        const double npni = (double)std::max(n_predictors - 1, 0);
        if (npni < 1.0) *overall = crit::kInf;
        else if (s.estimator == E_OLS) *overall = crit::f_ppf(1.0 - alpha, npni, df);
        else if (s.estimator == E_GLM) *overall = crit::chi2_ppf(1.0 - alpha, npni);
        else *overall = crit::kInf;
    }
}

// Posthoc correction thresholds for sample size n and `n_predictors` fitted columns: one CritValueTable per block
// (alpha = posthoc_alpha, family = the block's C(k,2) contrasts, Tukey k = the block's level count), concatenated
// in block order (build_posthoc_correction_crit_with_levels + build_crit_tables, batch.rs:2252-2405).
static void posthoc_corr_row(const Spec& s, uint32_t n, int n_predictors, double* out /*n_ph*/) {
    const long df_signed = (long)n - (long)n_predictors;
    if (df_signed < 1) invalid("df = N - P = " + std::to_string(n) + " - " + std::to_string(n_predictors) + " < 1");
    const double df = (double)df_signed;
    const double pa = std::isnan(s.posthoc_alpha) ? s.alpha : s.posthoc_alpha;
    const bool use_t = s.estimator == E_OLS;
    auto quant = [&](double p) { return use_t ? crit::t_ppf(p, df) : crit::norm_ppf(p); };
    size_t o = 0;
    for (size_t b = 0; b < s.posthoc_block_ncontrasts.size(); b++) {
        const int m = s.posthoc_block_ncontrasts[b];
        const double t_tukey = s.correction == C_TUKEY ? crit::q_tukey_ppf(1.0 - pa, (double)s.posthoc_block_levels[b], df) / crit::kSqrt2 : 0.0;
        for (int i = 0; i < m; i++) {
            double t = 0.0;
            switch (s.correction) {
                case C_NONE: t = quant(1.0 - pa / 2.0); break;
                case C_BONFERRONI: t = quant(1.0 - pa / (2.0 * (double)m)); break;
                case C_HOLM: {
                    const double eff = pa / (2.0 * (double)(m - i));
                    t = eff < 1e-12 ? crit::kInf : quant(1.0 - eff);
                    break;
                }
                case C_BH: {
                    const double eff = (double)(i + 1) / (double)m * pa / 2.0;
                    t = eff < 1e-12 ? crit::kInf : quant(1.0 - eff);
                    break;
                }
                case C_TUKEY: t = t_tukey; break;
            }
            out[o++] = t * t;
        }
    }
}

// data_gen.rs:943-961 — the deterministic largest-remainder walk consumes no RNG, so its level
// sequence is identical for every simulation: computed once per batch on the host.
static std::vector<uint8_t> exact_walk_levels(const double* probs, int nl, int max_n) {
    std::vector<uint8_t> out(max_n);
    std::vector<uint32_t> counts(nl, 0);
    double total = 0.0;
    for (int g = 0; g < nl; g++) total += probs[g];
    for (int i = 0; i < max_n; i++) {
        uint32_t assigned = 0;
        for (int g = 0; g < nl; g++) assigned += counts[g];
        const double t = (double)(assigned + 1);
        int best = 0;
        double best_def = -crit::kInf;
        for (int g = 0; g < nl; g++) {
            const double def = (probs[g] / total) * t - (double)counts[g];
            if (def > best_def) {
                best_def = def;
                best = g;
            }
        }
        counts[best]++;
        out[i] = (uint8_t)best;
    }
    return out;
}

static void check_supported(const Spec& s, bool generation_only = false) {
    if (!generation_only && s.estimator == E_MLE) {
        if (s.outcome != O_CONTINUOUS) unsupported("non-Gaussian mixed models (GLMM)");
        if (!s.contrast_pairs.empty()) unsupported("contrast_pairs with estimator mle (the reference's MLE arm has no contrasts)");
    }
    if (!generation_only && s.estimator == E_GLM) {
        if (s.has_cluster) unsupported("clustered GLM (GLMM)");
        if (!s.contrast_pairs.empty()) unsupported("contrast_pairs with estimator glm (the reference's GLM arm has no contrasts, batch.rs:2060-2062)");
        if (s.outcome == O_CONTINUOUS) invalid("estimator Glm requires a binary or count outcome");
    }
    if (s.between_cols.size() > MCP_MAX_P) unsupported("between_var_indices > " + std::to_string(MCP_MAX_P));
    if (s.posthoc_contrasts.size() > MCP_MAX_PH) unsupported("posthoc contrasts > " + std::to_string(MCP_MAX_PH));
    if (s.has_slopes_or_groupings) unsupported("more than two random slopes / more than two extra groupings / slopes on extra groupings");
    if (s.n_extra > 0 && s.P() > MCP_MAX_TEMPLATE_P) unsupported("extra groupings with n_predictors > " + std::to_string(MCP_MAX_TEMPLATE_P));
    if (s.n_extra > 0 && s.estimator != E_MLE) unsupported("extra groupings outside the mixed-model estimator");
    if (s.n_slopes > 0 && s.P() > MCP_MAX_TEMPLATE_P) unsupported("random slopes with n_predictors > " + std::to_string(MCP_MAX_TEMPLATE_P));
    if (s.new_dists.size() > MCP_MAX_MENU || s.resid_dists.size() > MCP_MAX_MENU) unsupported("scenario menus > " + std::to_string(MCP_MAX_MENU));
    for (int v : s.new_dists)
        if (v >= D_UPLOADED_FACTOR) unsupported("uploaded_* distributions in the scenario menu");
    if (s.k > MCP_MAX_K) unsupported("n_non_factor > " + std::to_string(MCP_MAX_K));
    if (s.P() > MCP_MAX_P) unsupported("n_predictors > " + std::to_string(MCP_MAX_P));
    if (s.T() > MCP_MAX_T) unsupported("targets > " + std::to_string(MCP_MAX_T));
    if (s.factor_n_levels.size() > MCP_MAX_FACTORS) unsupported("factors > " + std::to_string(MCP_MAX_FACTORS));
    if (s.factor_proportions.size() > MCP_MAX_LEVELS) unsupported("Σ factor levels > " + std::to_string(MCP_MAX_LEVELS));
    if (s.interactions.size() > MCP_MAX_INTER) unsupported("interactions > " + std::to_string(MCP_MAX_INTER));
    size_t comps = 0;
    for (auto& c : s.interactions) comps += c.size();
    if (comps > MCP_MAX_INTER_COMP) unsupported("interaction components > " + std::to_string(MCP_MAX_INTER_COMP));
}

// The lowered call, host side: the scalar part of the DevicePlan and ONE blob holding every device-side table (256-byte
// aligned sections).  Built once per call — the critical-value rows can cost milliseconds (Tukey quantiles) — and bound to
// each device the call drives with a single asynchronous upload.
struct Lowered {
    DevicePlan plan;
    int P = 0, T = 0, F = 0, n_ss = 0, cw = 0;
    static constexpr size_t NPOS = ~(size_t)0;
    std::vector<uint8_t> blob;
    size_t o_ss = NPOS, o_variants = NPOS, o_unc = NPOS, o_corr = NPOS, o_ov = NPOS, o_joint = NPOS, o_phu = NPOS, o_phc = NPOS,
           o_levels = NPOS, o_ec = NPOS, o_t3 = NPOS, o_un = NPOS, o_ud = NPOS;
    template <typename T_>
    size_t add(const std::vector<T_>& v) {
        const size_t off = (blob.size() + 255) & ~(size_t)255;
        blob.resize(off + std::max<size_t>(v.size(), 1) * sizeof(T_));
        if (!v.empty()) std::memcpy(blob.data() + off, v.data(), v.size() * sizeof(T_));
        return off;
    }
};

// Point the plan's table pointers at a device copy of the blob.
static void bind_tables(const Lowered& lw, DevicePlan& pl, const void* dev_blob) {
    const char* base = static_cast<const char*>(dev_blob);
    auto at = [&](size_t off) -> const void* { return off == Lowered::NPOS ? nullptr : base + off; };
    pl.ss = static_cast<const uint32_t*>(at(lw.o_ss));
    pl.variants = static_cast<const McpVariant*>(at(lw.o_variants));
    pl.crit_unc = static_cast<const double*>(at(lw.o_unc));
    pl.crit_corr = static_cast<const double*>(at(lw.o_corr));
    pl.crit_overall = static_cast<const double*>(at(lw.o_ov));
    pl.crit_joint = static_cast<const double*>(at(lw.o_joint));
    pl.crit_ph_unc = static_cast<const double*>(at(lw.o_phu));
    pl.crit_ph_corr = static_cast<const double*>(at(lw.o_phc));
    pl.levels = static_cast<const uint8_t*>(at(lw.o_levels));
    pl.exact_counts = static_cast<const uint32_t*>(at(lw.o_ec));
    pl.t3_table = static_cast<const double*>(at(lw.o_t3));
    pl.upload_normal = static_cast<const double*>(at(lw.o_un));
    pl.upload_data = static_cast<const double*>(at(lw.o_ud));
}

// SimulationSpec → DevicePlan scalars + the table blob (host only; no CUDA call).  Output pointers are filled by the caller.
static void lower(const Spec& s, const uint32_t* ss, uint32_t n_ss, Lowered& lw, bool thresholds = true) {
    DevicePlan& pl = lw.plan;
    std::memset(&pl, 0, sizeof pl);
    const int P = s.P(), T = s.T(), F = (int)s.factor_n_levels.size();
    // lw.P = the kernels' template width: P itself up to 16, else one of the padded wide widths (plan.h, mcp_template_p); the plan
    // carries the real P, and on a padded width every fit variant below keeps only the real columns
    lw.P = mcp_template_p(P); lw.T = T; lw.F = F; lw.n_ss = (int)n_ss;
    pl.k = s.k; pl.n_fd = s.n_fd; pl.n_factors = F; pl.n_inter = (int)s.interactions.size();
    pl.P = P; pl.T = T; pl.n_marg = (int)s.target_indices.size(); pl.n_contrasts = (int)s.contrast_pairs.size();
    pl.n_ss = (int)n_ss; pl.max_n = (int)ss[n_ss - 1];
    pl.estimator = s.estimator; pl.outcome = s.outcome; pl.link_probit = s.link_probit; pl.correction = s.correction;
    pl.report_overall = s.report_overall; pl.min_level_count = s.min_level_count;
    // design
    std::vector<double> L;
    if (s.k > 0 && !correlation_llt(s.correlation, s.k, L))
        throw EngineError(MCP_CORRELATION_NOT_PSD, "correlation matrix is not positive semi-definite");
    pl.l_identity = 1;
    for (int j = 0; j < s.k; j++)
        for (int m = 0; m < s.k; m++) {
            const double v = L[(size_t)j * s.k + m];
            pl.L[j * MCP_MAX_K + m] = v;
            pl.corr[m * MCP_MAX_K + j] = s.correlation[(size_t)m * s.k + j];
            if (v != (j == m ? 1.0 : 0.0)) pl.l_identity = 0;
        }
    pl.all_normal = 1;
    for (int j = 0; j < s.k; j++) {
        pl.var_type[j] = s.var_types[j];
        pl.var_pinned[j] = j < (int)s.var_pinned.size() ? s.var_pinned[j] : 0;
        pl.var_param[j] = j < (int)s.var_params.size() ? s.var_params[j] : 0.5;  // data_gen.rs:347
        if (s.var_types[j] != D_NORMAL) pl.all_normal = 0;
    }
    int off = 0, col = 1 + s.k;
    pl.exact_walk = 1;
    for (int j = 0; j < MCP_MAX_P; j++) pl.col_factor[j] = -1;
    for (int f = 0; f < F; f++) {
        pl.f_levels[f] = s.factor_n_levels[f];
        pl.f_off[f] = off;
        pl.f_col[f] = col;
        const int fs = f < (int)s.factor_sampled.size() ? s.factor_sampled[f] : -1;
        pl.f_sampled[f] = fs < 0 ? (s.sampled_factor_proportions ? 1 : 0) : fs;
        if (pl.f_sampled[f]) {
            pl.exact_walk = 0;
            pl.n_sampled_factors++;
        }
        for (int g = 0; g < s.factor_n_levels[f]; g++) pl.f_props[off + g] = s.factor_proportions[off + g];
        for (int d = 0; d < s.factor_n_levels[f] - 1 && col + d < MCP_MAX_P; d++) {
            pl.col_factor[col + d] = (int8_t)f;
            pl.col_level[col + d] = (int8_t)(d + 1);
        }
        off += s.factor_n_levels[f];
        col += s.factor_n_levels[f] - 1;
    }
    // uploaded-data arms: NORTA table / strict-bootstrap frame + column map
    pl.bootstrap = s.frame_map.empty() ? 0 : 1;
    pl.seq_per_row = pl.n_sampled_factors + pl.bootstrap;
    pl.un_rows = s.un_rows; pl.un_cols = s.un_cols; pl.ud_rows = s.ud_rows; pl.ud_cols = s.ud_cols;
    for (int c = 0; c < MCP_MAX_K + MCP_MAX_FACTORS; c++) pl.frame_map[c] = -1;
    for (int c = 0; c < (int)s.frame_map.size(); c++) pl.frame_map[c < s.k ? c : MCP_MAX_K + (c - s.k)] = s.frame_map[c];
    for (int j = 0; j < s.k; j++) pl.var_param_bs[j] = j < (int)s.var_params.size() ? s.var_params[j] : 0.0;
    if (!s.upload_normal.empty()) lw.o_un = lw.add(s.upload_normal);
    if (!s.upload_data.empty() && pl.bootstrap) lw.o_ud = lw.add(s.upload_data);
    if (pl.bootstrap) pl.exact_walk = 0;  // copied factor levels differ per study: per-study level counts, every exclusion mask
    int q = 0;
    for (size_t j = 0; j < s.interactions.size(); j++) {
        pl.inter_ptr[j] = q;
        for (int c : s.interactions[j]) pl.inter_cols[q++] = c;
    }
    pl.inter_ptr[s.interactions.size()] = q;
    for (int j = 0; j < P; j++) {
        pl.beta[j] = s.effect_sizes[j];
        if (s.effect_sizes[j] != 0.0) {
            pl.beta_nonzero_mask |= 1u << j;
            pl.nz_col[pl.n_nz] = (int8_t)j;
            pl.nz_beta[pl.n_nz++] = s.effect_sizes[j];
        }
    }
    pl.n_ph = (int)s.posthoc_contrasts.size();
    for (int c = 0; c < pl.n_ph; c++) {
        pl.ph_contrast[2 * c] = s.posthoc_contrasts[c].first;
        pl.ph_contrast[2 * c + 1] = s.posthoc_contrasts[c].second;
    }
    for (int t = 0; t < pl.n_marg; t++) pl.target[t] = s.target_indices[t];
    for (int t = 0; t < pl.n_contrasts; t++) {
        pl.contrast[2 * t] = s.contrast_pairs[t].first;
        pl.contrast[2 * t + 1] = s.contrast_pairs[t].second;
    }
    pl.residual_dist = s.residual_dist; pl.residual_pinned = s.residual_pinned; pl.residual_df = s.residual_df;
    pl.has_cluster = s.has_cluster; pl.sizing_fixed_size = s.sizing_fixed_size; pl.sizing_value = s.sizing_value;
    pl.tau_squared = s.tau_squared;
    if (s.has_cluster) {
        const int sv = std::max(s.sizing_value, 1);
        pl.n_clusters = s.sizing_fixed_size ? pl.max_n / sv : sv;  // generation.rs n_clusters_at(max_n)
    }
    // scenario.lme: random-effect distribution through draw_residual (data_gen.rs:146-161, 1289-1320), per-study ICC jitter
    pl.has_lme = s.has_lme;
    pl.re_dist = s.has_lme ? s.re_dist : R_NORMAL;
    pl.re_df = s.re_df;
    pl.icc_noise_sd = s.icc_noise_sd;
    {
        const double df = s.re_df > 3.0 ? s.re_df : 3.0;
        const long long dfi = std::max<long long>(std::llround(df), 1);
        pl.re_words = pl.re_dist == R_HIGH_KURTOSIS ? (int)(1 + dfi) : ((pl.re_dist == R_RIGHT_SKEWED || pl.re_dist == R_LEFT_SKEWED) ? (int)dfi : 1);
    }
    pl.n_slopes = s.has_cluster ? s.n_slopes : 0;
    for (int a = 0; a < 2; a++) {
        pl.slope_col[a] = s.slope_col[a];
        pl.slope_variance[a] = s.slope_variance[a];
        pl.slope_corr[a] = s.slope_corr[a];
    }
    pl.slope_corr12 = s.slope_corr12;
    pl.n_extra = s.has_cluster ? s.n_extra : 0;
    for (int g = 0; g < 2; g++) {
        pl.extra_nested[g] = s.extra_nested[g];
        pl.extra_count[g] = std::max(s.extra_count[g], 1);
        pl.extra_tau_squared[g] = s.extra_tau_squared[g];
        const int sv = std::max(s.sizing_value, 1);
        pl.extra_levels[g] = g >= pl.n_extra ? 0
                             : (!s.extra_nested[g] ? pl.extra_count[g]
                                                   : (s.sizing_fixed_size ? (pl.max_n + sv - 1) / sv : sv) * pl.extra_count[g]);  // generation.rs:301-320
    }
    // the fit keeps the extras' levels dense (Schur complement + Cholesky per REML evaluation, one thread per fit): bounded
    if (pl.extra_levels[0] + pl.extra_levels[1] > MCP_MAX_EXTRA_LEVELS)
        unsupported("extra groupings with more than " + std::to_string(MCP_MAX_EXTRA_LEVELS) + " levels in total at the largest sample size");
    // cluster-level predictors: only with a cluster (data_gen.rs:619-620).  A broadcast factor no longer follows the exact walk
    // row by row, so its per-level counts are read off the statistics like a sampled factor's.
    pl.n_between = s.has_cluster ? (int)s.between_cols.size() : 0;
    for (int b = 0; b < pl.n_between; b++) {
        pl.between_cols[b] = (int8_t)s.between_cols[b];
        if (s.between_cols[b] >= 1 + s.k) pl.exact_walk = 0;
    }
    pl.optimistic = s.is_optimistic();
    // scenario knobs (spec.rs:117-176)
    pl.heterogeneity = s.heterogeneity;
    pl.corr_noise_sd = s.corr_noise_sd;
    pl.dist_change_prob = s.dist_change_prob;
#ifndef MCP_COOP_FORCE
#define MCP_COOP_FORCE -1  // (tuning builds: 0 / 1 pins the choice)
#endif
    // the warp-cooperative marginal pass pays when a good share of the (study, column) blocks stays Normal — measured on the config-3
    // sweep: +8 % at a swap probability of 0.5, -5 % at 0.8 (profiles/r02_config3.md)
    pl.coop_marginals = MCP_COOP_FORCE >= 0 ? MCP_COOP_FORCE : (s.dist_change_prob > 0.0 && s.dist_change_prob < MCP_COOP_MAX_PROB ? 1 : 0);
    pl.resid_change_prob = s.resid_change_prob;
    pl.n_new_dists = (int)s.new_dists.size();
    for (size_t i = 0; i < s.new_dists.size(); i++) pl.new_dists[i] = s.new_dists[i];
    pl.n_resid_dists = (int)s.resid_dists.size();
    for (size_t i = 0; i < s.resid_dists.size(); i++) pl.resid_dists[i] = s.resid_dists[i];
    {
        // heteroskedasticity: gamma = ln(max(lambda,1))/4, driver moments from the BASE spec (spec.rs:436-515,
        // data_gen.rs:753-776); 1/exp(gamma²/2) uses the same owned exp as the kernel
        const double lambda = s.hsk_ratio > 1.0 ? s.hsk_ratio : 1.0;
        const double gamma = std::log(lambda) / 4.0;
        const int pf = 1 + s.k + s.n_fd;
        std::vector<double> mu(pf, 0.0), sd(pf, 0.0);
        mu[0] = 1.0;
        for (int j = 0; j < s.k; j++) {
            if (s.var_types[j] == D_BINARY) {
                const double pp = pl.var_param[j];
                mu[1 + j] = pp;
                sd[1 + j] = std::sqrt(std::max(pp * (1.0 - pp), 0.0));
            } else {
                sd[1 + j] = 1.0;
            }
        }
        int c = 1 + s.k, po = 0;
        for (int nl : s.factor_n_levels) {
            for (int d = 0; d + 1 < nl; d++) {
                const double pp = s.factor_proportions[po + d + 1];
                mu[c + d] = pp;
                sd[c + d] = std::sqrt(std::max(pp * (1.0 - pp), 0.0));
            }
            c += nl - 1;
            po += nl;
        }
        double lp_mean = 0.0, lp_var = 0.0;
        for (int j = 0; j < pf; j++) lp_mean += s.effect_sizes[j] * mu[j];
        for (int j = 0; j < pf; j++) lp_var += s.effect_sizes[j] * s.effect_sizes[j] * sd[j] * sd[j];
        for (int j = 0; j < s.k; j++)
            for (int m = j + 1; m < s.k; m++)
                lp_var += 2.0 * s.effect_sizes[1 + j] * s.effect_sizes[1 + m] * s.correlation[(size_t)m * s.k + j] * sd[1 + j] * sd[1 + m];
        const double lp_std = std::sqrt(std::max(lp_var, 0.0));
        double center = 0.0, stdv = 0.0;
        if (s.hsk_driver < 0) {
            center = lp_mean;
            stdv = lp_std;
        } else if (s.hsk_driver < pf) {
            center = mu[s.hsk_driver];
            stdv = sd[s.hsk_driver];
        }
        pl.hsk_gamma = gamma;
        pl.hsk_driver = s.hsk_driver;
        pl.apply_hsk = s.outcome == O_CONTINUOUS && gamma > 1e-30 && stdv > 1e-30;
        pl.hsk_center = center;
        pl.hsk_inv_std = pl.apply_hsk ? 1.0 / stdv : 0.0;
        pl.hsk_inv_norm = host_own_exp(-gamma * gamma / 2.0);
    }

    // variants + thresholds
    const bool track = thresholds && F > 0 && s.min_level_count > 0;
    if (!track) pl.min_level_count = 0;
    const int nv = track ? (1 << F) : 1;
    pl.n_variants = nv;
    pl.base_variant = 0;
    std::vector<McpVariant> variants(nv);
    for (int m = 0; m < nv; m++) variants[m] = make_variant(s, m);
    // which (variant, n_idx) pairs can occur?  Exact walks: exactly one mask per N (host-simulated).
    std::vector<std::vector<uint8_t>> levels(F);
    std::vector<uint8_t> levels_flat;
    bool any_exact = false;
    for (int f = 0; f < F; f++)
        if (!pl.f_sampled[f]) {
            levels[f] = exact_walk_levels(s.factor_proportions.data() + pl.f_off[f], s.factor_n_levels[f], pl.max_n);
            any_exact = true;
        }
    if (any_exact) {
        levels_flat.assign((size_t)F * pl.max_n, 0);
        for (int f = 0; f < F; f++)
            if (!levels[f].empty()) std::copy(levels[f].begin(), levels[f].end(), levels_flat.begin() + (size_t)f * pl.max_n);
    }
    const bool all_masks = track && (!pl.exact_walk || s.estimator == E_GLM);  // GLM: the separation fallback can drop any factor
    std::vector<std::vector<char>> needed(nv, std::vector<char>(n_ss, all_masks ? 1 : 0));
    if (!track) {
        for (uint32_t g = 0; g < n_ss; g++) needed[0][g] = 1;
    } else if (!all_masks) {
        for (uint32_t g = 0; g < n_ss; g++) {
            int mask = 0;
            for (int f = 0; f < F; f++) {
                std::vector<uint32_t> cnt(s.factor_n_levels[f], 0);
                for (uint32_t i = 0; i < ss[g]; i++) cnt[levels[f][i]]++;
                if (*std::min_element(cnt.begin(), cnt.end()) < (uint32_t)s.min_level_count) mask |= 1 << f;
            }
            needed[mask][g] = 1;
        }
    }
    const double nan = crit::kNaN;
    std::vector<double> c_unc((size_t)nv * n_ss, nan), c_corr((size_t)nv * n_ss * std::max(T, 1), nan), c_ov((size_t)nv * n_ss, nan);
    std::vector<double> c_joint((size_t)nv * n_ss, nan);
    if (s.estimator == E_MLE) {
        // joint Wald chi-square threshold: chi2(1-alpha, #testable targets) (critvals.rs:962-975, batch.rs:1455-1468)
        for (int m = 0; m < nv; m++) {
            int k_red = 0;
            for (int ti : s.target_indices) k_red += variants[m].remap[ti] >= 0;
            const bool reduced = variants[m].p_red != P;
            const double jc = !reduced ? (T > 0 ? crit::chi2_ppf(1.0 - s.alpha, (double)T) : nan)
                                       : (k_red == 0 ? crit::kInf : crit::chi2_ppf(1.0 - s.alpha, (double)k_red));
            for (uint32_t g = 0; g < n_ss; g++) c_joint[(size_t)m * n_ss + g] = jc;
        }
    }
    // the full-model table is always validated first, like build_crit_tables at batch entry (batch.rs:349-350)
    for (uint32_t g = 0; thresholds && g < n_ss; g++) {
        double dummy;
        crit_row(s, ss[g], P, &dummy, nullptr, nullptr, nullptr);
    }
    // thresholds depend on (N, p_red) only — memoise like reduced_crit_entry (batch.rs:2418-2437)
    const int NPH = (int)s.posthoc_contrasts.size();
    std::vector<double> c_phu((size_t)nv * n_ss, nan), c_phc((size_t)nv * n_ss * std::max(NPH, 1), nan);
    struct Memo { uint32_t n; int p; double unc, ov, phu; std::vector<double> corr, phc; };
    std::vector<Memo> memo;
    for (int m = 0; m < nv; m++)
        for (uint32_t g = 0; g < n_ss; g++) {
            if (!thresholds || !needed[m][g]) continue;
            const int pr = variants[m].p_red;
            if ((long)ss[g] - pr < 1) continue;
            const Memo* hit = nullptr;
            for (auto& e : memo)
                if (e.n == ss[g] && e.p == pr) hit = &e;
            if (!hit) {
                Memo e{ss[g], pr, 0, 0, 0, std::vector<double>(std::max(T, 1)), std::vector<double>(std::max(NPH, 1))};
                crit_row(s, ss[g], pr, &e.unc, e.corr.data(), &e.phu, &e.ov);
                if (NPH > 0) posthoc_corr_row(s, ss[g], pr, e.phc.data());
                memo.push_back(e);
                hit = &memo.back();
            }
            const size_t r = (size_t)m * n_ss + g;
            c_unc[r] = hit->unc;
            c_ov[r] = hit->ov;
            for (int t = 0; t < T; t++) c_corr[r * T + t] = hit->corr[t];
            c_phu[r] = hit->phu;
            for (int c = 0; c < NPH; c++) c_phc[r * NPH + c] = hit->phc[c];
        }
    std::vector<uint32_t> ssv(ss, ss + n_ss);
    lw.o_ss = lw.add(ssv);
    lw.o_variants = lw.add(variants);
    lw.o_unc = lw.add(c_unc);
    lw.o_corr = lw.add(c_corr);
    lw.o_ov = lw.add(c_ov);
    lw.o_joint = lw.add(c_joint);
    if (NPH > 0) {
        lw.o_phu = lw.add(c_phu);
        lw.o_phc = lw.add(c_phc);
    }
    if (any_exact) lw.o_levels = lw.add(levels_flat);
    if (F > 0 && pl.exact_walk) {  // exact walks are sim-invariant: the per-level prefix counts at every grid point too
        std::vector<uint32_t> ec((size_t)n_ss * MCP_MAX_LEVELS, 0);
        for (int f = 0; f < F; f++) {  // one pass over the walk per factor: running counts snapshotted at the grid points
            std::vector<uint32_t> run(s.factor_n_levels[f], 0);
            uint32_t g = 0;
            for (uint32_t i = 0; i < ss[n_ss - 1]; i++) {
                run[levels[f][i]]++;
                while (g < n_ss && ss[g] == i + 1) {
                    for (int l = 0; l < s.factor_n_levels[f]; l++) ec[(size_t)g * MCP_MAX_LEVELS + pl.f_off[f] + l] = run[l];
                    g++;
                }
            }
        }
        lw.o_ec = lw.add(ec);
    }
    bool need_t3 = false;
    for (int v : s.var_types) need_t3 |= v == D_HIGH_KURTOSIS;
    for (int v : s.new_dists) need_t3 |= v == D_HIGH_KURTOSIS;  // a scenario swap can introduce the t(3) marginal
    if (need_t3) lw.o_t3 = lw.add(crit::build_t3_table());
    pl.counter_width = mcp_counter_width(T, F, NPH);
    pl.off_hist_cor = 2 * T + 3 + (T + NPH);
    pl.off_fx = pl.off_hist_cor + (T + NPH) + 1;
    pl.off_joint = pl.off_fx + F;
    pl.off_bh = pl.off_joint + 1;
    pl.off_ph_unc = pl.off_bh + 1;
    pl.off_ph_cor = pl.off_ph_unc + NPH;
    pl.off_sing = pl.off_ph_cor + NPH;
    pl.off_fsep = pl.off_sing + 1;
    pl.off_pin = pl.off_fsep + F;
    lw.cw = pl.counter_width;
    pl.smem_counters = (size_t)n_ss * pl.counter_width * sizeof(unsigned) <= 16 * 1024;
}

// ---- wave transposition of the per-fit outputs ---------------------------------------------------------------------------
// The fit kernels run one (study, grid point) fit per thread with the studies of a warp at the SAME grid point: in the
// sim-major BatchResult layout their bytes lie n_ss·W apart (every store its own sector: the 50-point sweep wrote its 1.1 GB
// at ~100 GB/s and took 15 % longer than in COUNTS mode).  Waves of a multi-point BYTES / CAPTURE call therefore write a
// grid-point-major staging block ([g][sim][W], neighbouring studies = neighbouring bytes) and this kernel moves the wave into
// the sim-major tensors: coalesced reads along the studies, coalesced writes along the grid points, through shared memory.
struct TransposeDesc {
    const uint8_t* in;   // [n_ss][cnt][w]
    uint8_t* out;        // [cnt][n_ss][w] (already offset to the wave's first sim)
    uint32_t w;          // bytes per cell
};
struct TransposeArgs {
    TransposeDesc d[10];
};
constexpr int kTransposeSims = 32, kTransposeThreads = 256, kTransposeTileBytes = 32 * 1024;

__global__ void __launch_bounds__(kTransposeThreads) transpose_outputs_kernel(TransposeArgs a, uint32_t cnt, uint32_t n_ss) {
    extern __shared__ uint8_t tile[];  // [grid point in chunk][sim in tile][w]: a flat copy of the source rows
    const TransposeDesc d = a.d[blockIdx.z];
    const uint32_t w = d.w;
    const uint32_t gc = max(1u, min(n_ss, (uint32_t)kTransposeTileBytes / (kTransposeSims * w)));  // grid points per chunk
    const uint32_t s0 = blockIdx.x * kTransposeSims, ns = min((uint32_t)kTransposeSims, cnt - s0);
    const uint32_t g0 = blockIdx.y * gc;
    if (g0 >= n_ss) return;
    const uint32_t ng = min(gc, n_ss - g0);
    const uint32_t row = ns * w;  // contiguous bytes of one grid point of this tile
    for (uint32_t gl = 0; gl < ng; gl++) {
        const uint8_t* src = d.in + ((size_t)(g0 + gl) * cnt + s0) * w;
        for (uint32_t i = threadIdx.x; i < row; i += kTransposeThreads) tile[gl * row + i] = src[i];
    }
    __syncthreads();
    // destination: for sim s the ng·w bytes of its grid points g0 .. g0 + ng − 1 are contiguous; one warp per sim
    const uint32_t out_row = ng * w;
    for (uint32_t sl = threadIdx.x >> 5; sl < ns; sl += kTransposeThreads / 32) {
        uint8_t* dst = d.out + ((size_t)(s0 + sl) * n_ss + g0) * w;
        uint32_t i = threadIdx.x & 31u;
        uint32_t gl = i / w, b = i - gl * w;  // i = gl·w + b, advanced by 32 per step without dividing again
        const uint32_t dg = 32u / w, db = 32u - dg * w;
        for (; i < out_row; i += 32u) {
            dst[i] = tile[gl * row + sl * w + b];
            gl += dg;
            b += db;
            if (b >= w) {
                b -= w;
                gl++;
            }
        }
    }
}

// ---- kernel dispatch: ask each kernel translation unit in turn (launch.h) -------------------------
// bytes of one per-study scenario design at template width P (an upper bound within the width class)
static size_t scen_design_bytes(int P) {
    return P <= MCP_MAX_TEMPLATE_P ? sizeof(ScenDesign<MCP_MAX_TEMPLATE_P>) : (P <= MCP_WIDE_P_A ? sizeof(ScenDesign<MCP_WIDE_P_A>) : sizeof(ScenDesign<MCP_WIDE_P_B>));
}

static void dispatch_power(int P, int kind, const DevicePlan& plan, const KernelArgs& a) {
    cudaError_t err = cudaSuccess;
    bool handled = false;
#define X(id) if (!handled) handled = tu##id##_launch_power(P, kind, !plan.optimistic, plan, a, &err);
    MCP_TU_LIST(X)
#undef X
    if (!handled) unsupported("n_predictors = " + std::to_string(P));
    CUDA_CHECK(err);
}

template <typename F>
static int guarded(F&& body) {
    try {
        body();
        return MCP_OK;
    } catch (const EngineError& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_last_error = std::string("internal error: ") + e.what();
        return MCP_CUDA_ERROR;
    }
}

// The devices one call drives (McpOptions.device / n_devices), ascending — also the lock order.
static std::vector<int> resolve_devices(const McpOptions& opt) {
    const int visible = device_count_checked();
    int first = opt.device;
    if (first < 0) CUDA_CHECK(cudaGetDevice(&first));
    int n = opt.n_devices == 0 ? 1 : opt.n_devices;
    if (n < 0) n = visible - first;
    if (first >= visible || n < 1 || first + n > visible)
        throw EngineError(MCP_CUDA_ERROR, "devices " + std::to_string(first) + ".." + std::to_string(first + n - 1) + " requested, " +
                                              std::to_string(visible) + " visible");
    std::vector<int> out(n);
    for (int i = 0; i < n; i++) out[i] = first + i;
    return out;
}

// One device's share of a call: a contiguous range of the call's sims.
struct Shard {
    DeviceCtx* cx = nullptr;
    uint64_t begin = 0;  // first sim of the shard, relative to the call
    uint32_t count = 0;
    DevicePlan pl;
    unsigned launches = 0, waves = 0;
    uint64_t h2d = 0, d2h = 0;
    uint32_t units_per_sim = 1;  // progress units the kernels count per sim (OLS: 1 study; GLM / MLE: n_ss fits)
    bool copies = false, done = false;
    float ms = 0.f;
};

// Host views of the call-wide result tensors (element sizes per cell follow BatchResult, spec.rs:552-600).
struct HostOut {
    uint8_t *unc = nullptr, *cor = nullptr, *conv = nullptr, *overall = nullptr, *fx = nullptr, *joint = nullptr, *bh = nullptr, *ph_unc = nullptr,
            *ph_cor = nullptr;
    double* tsq = nullptr;
};

// Everything one device has to do for its shard, enqueued without a single host-side wait: table upload, waves of kernel
// launches on the compute stream, device→host copies of each finished wave on the copy stream (BYTES / CAPTURE: wave w travels
// while wave w + 1 computes), counters into the page-locked staging buffer.
static void enqueue_shard(Shard& sh, const Lowered& lw, const Spec& s, uint32_t n_ss, uint64_t base_seed, uint64_t call_sim_begin,
                          const McpOptions& opt, const HostOut& ho, bool sink) {
    DeviceCtx& cx = *sh.cx;
    CUDA_CHECK(cudaSetDevice(cx.dev));
    const cudaStream_t st = cx.compute;
    DevicePlan& pl = sh.pl;
    pl = lw.plan;
    const int T = lw.T, F = lw.F;
    const uint32_t n_sims = sh.count;
    const bool bytes = opt.mode & MCP_MODE_BYTES, capture = opt.mode & MCP_MODE_CAPTURE;
    const bool glm = s.estimator == E_GLM, mle = s.estimator == E_MLE;
    const int kind = glm ? KIND_GLM : (mle ? KIND_LMM : KIND_OLS);
    sh.copies = bytes || capture;
    // ---- tables: one blob, one upload
    void* dblob = cx.tables.get(lw.blob.size());
    CUDA_CHECK(cudaMemcpyAsync(dblob, lw.blob.data(), lw.blob.size(), cudaMemcpyHostToDevice, st));
    sh.h2d += lw.blob.size();
    bind_tables(lw, pl, dblob);
    // ---- per-call device state: counters, progress counter / cancel flag / error flag (one 256-byte block)
    const size_t ncnt = (size_t)n_ss * lw.cw;
    pl.counters = static_cast<unsigned long long*>(cx.small.get(Carver::pad(ncnt * sizeof(unsigned long long))));
    CUDA_CHECK(cudaMemsetAsync(pl.counters, 0, ncnt * sizeof(unsigned long long), st));
    CUDA_CHECK(cudaMemsetAsync(cx.d_prog, 0, 256, st));
    pl.error_flag = cx.d_error;
    sh.units_per_sim = kind == KIND_OLS ? 1u : n_ss;
    if (sink) {
        const uint64_t units = (uint64_t)n_sims * sh.units_per_sim;
        if (units > 0xffffffffull) unsupported("progress reporting for more than 2^32 fits per device and call");
        pl.prog_counter = cx.d_prog;
        pl.prog_host = mapped_progress_word(cx);
        pl.prog_step = (uint32_t)std::max<uint64_t>(units / 50, 1);  // batch.rs:449-450
        pl.cancel_flag = cx.d_cancel;
        cx.mail->progress = 0;
    }
    // ---- device-side result tensors of the shard
    const size_t cells = (size_t)n_sims * n_ss;
    {
        size_t total = 0;
        if (bytes) {
            total += 2 * Carver::pad(cells * T) + Carver::pad(cells);
            if (s.report_overall) total += Carver::pad(cells);
            if (F > 0) total += Carver::pad(cells * F);
            if (pl.n_ph > 0) total += 2 * Carver::pad(cells * pl.n_ph);
            if (mle) total += 2 * Carver::pad(cells);
        }
        if (capture) total += Carver::pad(cells * T * sizeof(double));
        Carver oc(total ? cx.outputs.get(total) : nullptr);
        if (bytes) {
            pl.out_unc = oc.take<uint8_t>(cells * T);
            pl.out_cor = oc.take<uint8_t>(cells * T);
            pl.out_conv = oc.take<uint8_t>(cells);
            if (s.report_overall) pl.out_overall = oc.take<uint8_t>(cells);
            if (F > 0) pl.out_fx = oc.take<uint8_t>(cells * F);
            if (pl.n_ph > 0) {
                pl.out_ph_unc = oc.take<uint8_t>(cells * pl.n_ph);
                pl.out_ph_cor = oc.take<uint8_t>(cells * pl.n_ph);
            }
            if (mle) {
                pl.out_joint = oc.take<uint8_t>(cells);
                pl.out_bh = oc.take<uint8_t>(cells);
            }
        }
        if (capture) pl.out_tsq = oc.take<double>(cells * T);
    }
    // ---- wave size.  The batch is never cut for progress reporting (the kernels publish it); waves exist only where the scratch
    // arena bounds the studies in flight, and — BYTES / CAPTURE — so that the copy of one wave overlaps the next wave's kernels.
    uint32_t wave = opt.wave_sims ? opt.wave_sims : n_sims;
    wave = std::max<uint32_t>(wave, 1);
    if (!opt.wave_sims && sh.copies) {
        size_t per_sim = 0;
        if (bytes) per_sim += (size_t)n_ss * (2 * T + 1 + (s.report_overall ? 1 : 0) + F + 2 * pl.n_ph + (mle ? 2 : 0));
        if (capture) per_sim += (size_t)n_ss * T * sizeof(double);
        // pieces of >= 64 MiB and >= 64 Ki studies, at most 8: every wave boundary drains the device once (a 62 Ki-study wave is 3.3
        // rounds of resident CTAs: sixteen of them cost the 50-point sweep 12 % of its kernel time), while one wave's copy hides behind
        // the next wave's kernels — on a shard of a multi-device call (8 GPUs x 144 MB of config 2) two waves hide half of the host
        // ingest that otherwise follows the kernels
        const size_t out_bytes = per_sim * n_sims;
        const uint32_t pieces = (uint32_t)std::min<size_t>(std::max<size_t>(out_bytes >> 26, 1), 8);
        if (pieces > 1) wave = std::max<uint32_t>((((n_sims + pieces - 1) / pieces + 127u) / 128u) * 128u, 1u << 16);
    }
    float* glm_scratch = nullptr;
    if (glm) {
        // IRLS makes several passes over each study's data: the f32 plane of a wave is kept in HBM as
        // [row][column][sim] (coalesced across the sims of a warp); the wave is capped at ~6 GiB of scratch
        // (+ the per-study scenario designs behind the plane on the scenario path, sized for the widest template)
        const size_t per_sim = (size_t)pl.max_n * lw.P * sizeof(float) + (pl.optimistic ? 0 : scen_design_bytes(lw.P));
        const uint32_t cap = (uint32_t)std::max<size_t>(((size_t)6 << 30) / std::max<size_t>(per_sim, 1), 128);
        wave = std::min(wave, cap);
        wave = std::max<uint32_t>((wave / 128u) * 128u, 128u);
        glm_scratch = static_cast<float*>(cx.scratch.get((size_t)wave * per_sim));
    }
    double* lmm_scratch = nullptr;
    bool lmm_quad = false;  // mixed-model generator on the quad kernel (cluster sums fit in shared memory; kernel_tu.cu decides the same way)
    if (mle) {
        const int P_ = lw.P, STUDY_ = mcp_stage_geom((P_ + 2) / 2).study;
        const int cs_ = std::max(s.sizing_value, 1);
        const size_t nc_ = s.sizing_fixed_size ? (size_t)(pl.max_n + cs_ - 1) / cs_ : (size_t)cs_;
        const size_t smem_q = (size_t)MCP_BLOCK * 4 * (P_ | 1) * sizeof(float) + (size_t)(MCP_BLOCK / 4) * STUDY_ * sizeof(double) +
                              (size_t)(MCP_BLOCK / 4) * (((nc_ * (P_ + 1)) | 1) + (nc_ | 1)) * sizeof(double);
        lmm_quad = smem_q <= 56 * 1024 && pl.n_slopes == 0 && pl.n_extra == 0;  // slope sums / extra groupings: one-thread-per-study generators
        // per-cluster sums [cluster][column][sim] in HBM; wave capped at ~4 GiB of scratch
        // + a snapshot of them and the X'X | X'y | y'y statistics per grid point (LmmLayout in kernels_lmm.cuh)
        size_t per_sim = (nc_ * (lw.P + 1) * (1 + pl.n_slopes) * (1 + (size_t)n_ss) + (size_t)n_ss * ((size_t)lw.P * (lw.P + 1) / 2 + lw.P + 1)) * sizeof(double) +
                         (pl.optimistic ? 0 : scen_design_bytes(lw.P));  // + scenario designs behind the layout
        if (pl.n_extra > 0) {  // extra groupings: the f32 plane + one dense scratch region per (study, sample size) fit (GenLayout)
            const size_t K = nc_, q0 = 1 + (size_t)pl.n_slopes, Qe = (size_t)pl.extra_levels[0] + (size_t)pl.extra_levels[1];
            const size_t fit_doubles = 2 * (K * q0 * (q0 + 1) / 2 + Qe * K * q0 + Qe * (Qe + 1) / 2 + (K * q0 + Qe) * (size_t)(lw.P + 1));
            per_sim = (size_t)pl.max_n * lw.P * sizeof(float) + 8 + (size_t)n_ss * fit_doubles * sizeof(double);
            lmm_quad = false;
        }
        const uint32_t cap = (uint32_t)std::max<size_t>(((size_t)4 << 30) / std::max<size_t>(per_sim, 1), 128);
        wave = std::min(wave, cap);
        wave = std::max<uint32_t>(((wave + 127u) / 128u) * 128u, 128u);  // the kernels' stride: whole 128-study CTAs
        if (wave > cap) wave = std::max<uint32_t>((cap / 128u) * 128u, 128u);
        lmm_scratch = static_cast<double*>(cx.scratch.get((size_t)wave * per_sim));
    }
    // OLS kernel choice (results are bit-identical whichever runs; tests/test_gpu_parity.py checks that):
    //   5 lean path — plain linear models (Normal marginals/residual, exact-walk factors, no cluster/scenario): lane-quad
    //     generate+accumulate kernel leaves X'X|X'y|y'y per fit in HBM, a thread-per-fit kernel solves;
    //   6 split path — everything else on the OLS arm: the general lane-quad kernel in generate+accumulate form
    //     (scenario designs from their own one-thread-per-study kernel), same statistics buffer, same solve kernel;
    //   4 general quad kernel solving by call, 1 thread-per-study kernel — kept for cross-checks (tests select them through
    //     the process-wide test hook mcp_test_force_ols_variant; the product path never reads the environment).
    const bool wide = lw.P != pl.P;  // padded width: the split path's kernels only
    const bool lean_ok = !wide && kind == KIND_OLS && pl.optimistic && pl.outcome == O_CONTINUOUS && pl.all_normal &&
                         (pl.n_factors == 0 || pl.exact_walk) && pl.residual_dist == R_NORMAL && !pl.has_cluster && !pl.bootstrap;
    int ols_variant = lean_ok ? 5 : 6;
    switch (g_force_ols_variant.load(std::memory_order_relaxed)) {
        case 1: ols_variant = 1; break;
        case 4: ols_variant = 4; break;
        case 5: ols_variant = lean_ok ? 5 : 4; break;
        case 6: ols_variant = 6; break;
        default: break;
    }
    if (wide) ols_variant = 6;
    if (kind == KIND_OLS && (ols_variant == 5 || ols_variant == 6)) {
        // statistics buffer [grid point][entry][sim] in HBM (448 B per fit at P = 9); wave capped at ~4 GiB of scratch, ~12 GiB
        // on the scenario path: its one-thread-per-study design kernel is latency-bound (Jacobi repair in local memory)
        // and wants as many studies in flight as the 180 GB part allows
        const size_t E = (size_t)lw.P * (lw.P + 1) / 2 + lw.P + 1;
        const size_t per_sim = (size_t)n_ss * E * sizeof(double);
        const size_t cap_bytes = (ols_variant == 6 && !pl.optimistic) ? (size_t)12 << 30 : (size_t)4 << 30;
        const uint32_t cap = (uint32_t)std::max<size_t>(cap_bytes / per_sim, 128);
        wave = std::min(wave, cap);
        if (wave < n_sims) wave = std::max<uint32_t>((wave / 128u) * 128u, 128u);
        // variant 6 on the scenario path keeps the per-study scenario designs behind the statistics (sized for the widest template)
        const size_t design_bytes = ols_variant == 6 && !pl.optimistic ? scen_design_bytes(lw.P) : 0;
        lmm_scratch = static_cast<double*>(cx.scratch.get((size_t)wave * (n_ss * E * sizeof(double) + design_bytes)));
    }
    // grid-point-major staging of one wave's per-fit outputs (see transpose_outputs_kernel); a single grid point needs none
    const bool staged = sh.copies && n_ss > 1;
    DevicePlan stg;  // the out_* pointers of the staging block
    std::memset(&stg, 0, sizeof stg);
    if (staged) {
        const size_t wc = (size_t)std::min<uint32_t>(wave, n_sims) * n_ss;
        size_t total = 0;
        if (bytes) {
            total += 2 * Carver::pad(wc * T) + Carver::pad(wc);
            if (pl.out_overall) total += Carver::pad(wc);
            if (pl.out_fx) total += Carver::pad(wc * F);
            if (pl.out_ph_unc) total += 2 * Carver::pad(wc * pl.n_ph);
            if (pl.out_joint) total += 2 * Carver::pad(wc);
        }
        if (capture) total += Carver::pad(wc * T * sizeof(double));
        Carver sc(cx.staging.get(total));
        if (bytes) {
            stg.out_unc = sc.take<uint8_t>(wc * T);
            stg.out_cor = sc.take<uint8_t>(wc * T);
            stg.out_conv = sc.take<uint8_t>(wc);
            if (pl.out_overall) stg.out_overall = sc.take<uint8_t>(wc);
            if (pl.out_fx) stg.out_fx = sc.take<uint8_t>(wc * F);
            if (pl.out_ph_unc) {
                stg.out_ph_unc = sc.take<uint8_t>(wc * pl.n_ph);
                stg.out_ph_cor = sc.take<uint8_t>(wc * pl.n_ph);
            }
            if (pl.out_joint) {
                stg.out_joint = sc.take<uint8_t>(wc);
                stg.out_bh = sc.take<uint8_t>(wc);
            }
        }
        if (capture) stg.out_tsq = sc.take<double>(wc * T);
    }
    if (pl.n_between > 0) {  // representative-row values of the cluster-level predictors, [cluster][column][study in wave]
        const size_t cs = (size_t)std::max(s.sizing_value, 1);
        const size_t n_rep = s.sizing_fixed_size ? ((size_t)pl.max_n + cs - 1) / cs : std::min<size_t>(cs, (size_t)pl.max_n);
        pl.between_stride = wave;
        pl.between_tab = static_cast<float*>(cx.between.get(n_rep * pl.n_between * (size_t)wave * sizeof(float)));
    }
    // accumulate / generate + solve / fit per wave, + the scenario-design kernel on the scenario path (the thread-per-study
    // cross-check kernels, variants 1 and 4, are single launches)
    const bool two_stage = kind != KIND_OLS || ols_variant == 5 || ols_variant == 6;
    const unsigned launches_per_wave = two_stage ? ((pl.optimistic || (mle && !lmm_quad)) ? 2 : 3) : 1;
    CUDA_CHECK(cudaEventRecord(cx.e0, st));
    for (uint64_t done = 0; done < n_sims;) {
        const uint32_t cnt = (uint32_t)std::min<uint64_t>(wave, n_sims - done);
        // the byte tensors are indexed by the sim's position in the shard, so waves shift the base pointers
        DevicePlan shifted = pl;
        if (staged) {  // this wave writes the staging block, grid-point-major with stride = its own study count
            shifted.out_stride = cnt;
            shifted.out_unc = stg.out_unc; shifted.out_cor = stg.out_cor; shifted.out_conv = stg.out_conv;
            shifted.out_overall = stg.out_overall; shifted.out_fx = stg.out_fx; shifted.out_joint = stg.out_joint;
            shifted.out_bh = stg.out_bh; shifted.out_ph_unc = stg.out_ph_unc; shifted.out_ph_cor = stg.out_ph_cor;
            shifted.out_tsq = stg.out_tsq;
        } else if (done > 0 && sh.copies) {
            if (bytes) {
                shifted.out_unc = pl.out_unc + done * n_ss * T;
                shifted.out_cor = pl.out_cor + done * n_ss * T;
                shifted.out_conv = pl.out_conv + done * n_ss;
                if (pl.out_overall) shifted.out_overall = pl.out_overall + done * n_ss;
                if (pl.out_fx) shifted.out_fx = pl.out_fx + done * n_ss * F;
                if (pl.out_joint) shifted.out_joint = pl.out_joint + done * n_ss;
                if (pl.out_bh) shifted.out_bh = pl.out_bh + done * n_ss;
                if (pl.out_ph_unc) {
                    shifted.out_ph_unc = pl.out_ph_unc + done * n_ss * pl.n_ph;
                    shifted.out_ph_cor = pl.out_ph_cor + done * n_ss * pl.n_ph;
                }
            }
            if (capture) shifted.out_tsq = pl.out_tsq + done * n_ss * T;
        }
        KernelArgs ka{base_seed, call_sim_begin + sh.begin + done, cnt, st, glm_scratch, lmm_scratch, ols_variant};
        dispatch_power(lw.P, kind, shifted, ka);
        sh.h2d += sizeof(DevicePlan);
        sh.launches += launches_per_wave;
        if (staged) {  // staging → the sim-major device tensors, all of the wave's tensors in one launch
            TransposeArgs ta;
            std::memset(&ta, 0, sizeof ta);
            unsigned nt = 0, wmax = 1;
            auto add = [&](const void* in, void* out, size_t w) {
                if (!in || !out || w == 0) return;
                ta.d[nt].in = static_cast<const uint8_t*>(in);
                ta.d[nt].out = static_cast<uint8_t*>(out) + done * n_ss * w;
                ta.d[nt].w = (uint32_t)w;
                wmax = std::max<unsigned>(wmax, (unsigned)w);
                nt++;
            };
            add(stg.out_unc, pl.out_unc, (size_t)T);
            add(stg.out_cor, pl.out_cor, (size_t)T);
            add(stg.out_conv, pl.out_conv, 1);
            add(stg.out_overall, pl.out_overall, 1);
            add(stg.out_fx, pl.out_fx, (size_t)F);
            add(stg.out_joint, pl.out_joint, 1);
            add(stg.out_bh, pl.out_bh, 1);
            add(stg.out_ph_unc, pl.out_ph_unc, (size_t)pl.n_ph);
            add(stg.out_ph_cor, pl.out_ph_cor, (size_t)pl.n_ph);
            add(stg.out_tsq, pl.out_tsq, (size_t)T * sizeof(double));
            if (nt > 0) {
                // grid.y covers the tensor with the narrowest cells (most grid points per chunk is per tensor: blocks past a
                // tensor's last chunk return at once)
                const unsigned gc_min = std::max(1u, std::min<unsigned>(n_ss, (unsigned)kTransposeTileBytes / (kTransposeSims * wmax)));
                const dim3 grid((cnt + kTransposeSims - 1) / kTransposeSims, (n_ss + gc_min - 1) / gc_min, nt);
                static bool attr_set[64] = {};
                if (!attr_set[cx.dev]) {
                    CUDA_CHECK(cudaFuncSetAttribute(transpose_outputs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTransposeTileBytes + 8192));
                    attr_set[cx.dev] = true;
                }
                transpose_outputs_kernel<<<grid, kTransposeThreads, kTransposeTileBytes + 8192, st>>>(ta, cnt, n_ss);
                CUDA_CHECK(cudaGetLastError());
                sh.launches += 1;
            }
        }
        if (sh.copies) {  // this wave's slices travel on the copy stream while the next wave computes
            cudaEvent_t ev = cx.wave_event(sh.waves);
            CUDA_CHECK(cudaEventRecord(ev, st));
            CUDA_CHECK(cudaStreamWaitEvent(cx.copy, ev, 0));
            const uint64_t at = sh.begin + done;  // position in the call-wide host tensors
            auto fetch = [&](auto* host, const auto* dev, size_t row) {
                if (!host || !dev || row == 0) return;
                using Elem = std::remove_cv_t<std::remove_pointer_t<decltype(dev)>>;
                const size_t nb = (size_t)cnt * row * sizeof(Elem);
                CUDA_CHECK(cudaMemcpyAsync(host + at * row, dev + done * row, nb, cudaMemcpyDeviceToHost, cx.copy));
                sh.d2h += nb;
            };
            fetch(ho.unc, pl.out_unc, (size_t)n_ss * T);
            fetch(ho.cor, pl.out_cor, (size_t)n_ss * T);
            fetch(ho.conv, pl.out_conv, (size_t)n_ss);
            fetch(ho.overall, pl.out_overall, (size_t)n_ss);
            fetch(ho.fx, pl.out_fx, (size_t)n_ss * F);
            fetch(ho.joint, pl.out_joint, (size_t)n_ss);
            fetch(ho.bh, pl.out_bh, (size_t)n_ss);
            fetch(ho.ph_unc, pl.out_ph_unc, (size_t)n_ss * pl.n_ph);
            fetch(ho.ph_cor, pl.out_ph_cor, (size_t)n_ss * pl.n_ph);
            fetch(ho.tsq, pl.out_tsq, (size_t)n_ss * T);
        }
        sh.waves++;
        done += cnt;
    }
    CUDA_CHECK(cudaEventRecord(cx.e1, st));
    unsigned long long* hc = cx.counters_staging(ncnt);
    CUDA_CHECK(cudaMemcpyAsync(hc, pl.counters, ncnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(&cx.mail->error_flag, cx.d_error, sizeof(int), cudaMemcpyDeviceToHost, st));
    sh.d2h += ncnt * sizeof(unsigned long long) + sizeof(int);
    CUDA_CHECK(cudaEventRecord(cx.done_compute, st));
    if (sh.copies) CUDA_CHECK(cudaEventRecord(cx.done_copy, cx.copy));
}

static bool shard_finished(Shard& sh) {
    if (sh.done) return true;
    cudaError_t e = cudaEventQuery(sh.cx->done_compute);
    if (e == cudaErrorNotReady) return false;
    CUDA_CHECK(e);
    if (sh.copies) {
        e = cudaEventQuery(sh.cx->done_copy);
        if (e == cudaErrorNotReady) return false;
        CUDA_CHECK(e);
    }
    sh.done = true;
    return true;
}

}  // namespace mcp

using namespace mcp;

extern "C" {

const char* mcp_last_error(void) { return g_last_error.c_str(); }
const char* mcp_version(void) { return "mcpower_b200 0.2.0 (sm_100a)"; }
uint64_t mcp_splitmix64(uint64_t z) { return splitmix64(z); }

int mcp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int mcp_test_force_ols_variant(int variant) {
    return g_force_ols_variant.exchange(variant, std::memory_order_relaxed);
}

void mcp_free_result(McpBatchResult* r) {
    if (!r) return;
    void* owned[] = {r->uncorrected, r->corrected, r->converged, r->overall, r->factor_excluded, r->joint, r->boundary_hit,
                     r->t_sq, r->posthoc_unc, r->posthoc_cor, r->counters};
    for (void* p : owned)
        if (p && !g_pinned.give(p)) std::free(p);
    std::free(r);
}

// run_batch on a parsed spec: validation, lowering (once), one shard per device — everything enqueued asynchronously —, then
// the calling thread waits: it polls the devices' progress words and events, serves the ProgressSink, raises the cancel flag.
// Shared by mcp_run_batch and the orchestrator entry points (mcp_find_power / mcp_find_sample_size).
// progress_offset / progress_total: the (current, total) the sink sees are offset + done and total (a multi-scenario call
// reports one range over all scenarios, find_power.rs:337-380); total 0 = this batch alone.
static McpBatchResult* run_batch_core(const Spec& s, const uint32_t* sample_sizes, uint32_t n_ss, uint32_t n_sims, uint64_t base_seed,
                                      const McpOptions* options, McpProgressFn progress, void* user, uint64_t progress_offset = 0,
                                      uint64_t progress_total = 0) {
    McpOptions opt{-1, MCP_MODE_COUNTS, 0, 0, 0, nullptr};
    if (options) opt = *options;
    if (opt.mode == 0) opt.mode = MCP_MODE_COUNTS;
    const auto t_start = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    if (!sample_sizes) invalid("sample_sizes pointer is NULL");
    validate_spec(s, sample_sizes, n_ss);
    if (n_ss > 65535u) invalid("more than 65535 sample sizes in one batch");  // grid.y of the fit kernels
    check_supported(s);
    Lowered lw;
    lower(s, sample_sizes, n_ss, lw);  // host only: every spec-level error surfaces before a device is touched
    const double lower_ms = ms_since(t_start);
    const auto t_enqueue = std::chrono::steady_clock::now();
    const std::vector<int> devices = resolve_devices(opt);
    const int T = lw.T, F = lw.F;
    const bool bytes = opt.mode & MCP_MODE_BYTES, capture = opt.mode & MCP_MODE_CAPTURE;
    const bool mle = s.estimator == E_MLE;
    const bool sink = progress != nullptr || opt.is_cancelled != nullptr;

    // ---- shards: contiguous sim-id ranges, device g takes [g·ceil(S/G), ...); empty shards are dropped
    std::vector<Shard> shards;
    {
        const uint64_t G = devices.size(), per = (n_sims + G - 1) / G;
        for (uint64_t g = 0; g < G; g++) {
            const uint64_t b = std::min<uint64_t>(g * per, n_sims), e = std::min<uint64_t>(b + per, n_sims);
            if (e > b || (g == 0 && n_sims == 0)) {
                Shard sh;
                sh.cx = &ctx_for(devices[g]);
                sh.begin = b;
                sh.count = (uint32_t)(e - b);
                shards.push_back(sh);
            }
        }
    }
    std::vector<std::unique_lock<std::mutex>> locks;
    for (auto& sh : shards) locks.emplace_back(sh.cx->mu);  // ascending ordinals: no lock-order inversion between calls

    // ---- the call-wide result (host)
    McpBatchResult* r = static_cast<McpBatchResult*>(std::calloc(1, sizeof(McpBatchResult)));
    if (!r) throw EngineError(MCP_CUDA_ERROR, "out of host memory");
    r->n_sims = n_sims; r->n_sample_sizes = n_ss; r->n_targets = (uint32_t)T; r->n_factors = (uint32_t)F;
    r->counter_width = (uint32_t)lw.cw;
    r->n_posthoc = (uint32_t)lw.plan.n_ph;
    r->n_devices = (uint32_t)shards.size();
    bool enqueued = false;
    struct Guard {  // on any failure: let the devices drain (they may still write into the result buffers), then free
        McpBatchResult* r; std::vector<Shard>* shards; bool* enqueued; bool ok = false;
        ~Guard() {
            if (ok) return;
            if (*enqueued)
                for (auto& sh : *shards) {
                    cudaSetDevice(sh.cx->dev);
                    cudaDeviceSynchronize();
                }
            mcp_free_result(r);
        }
    } guard{r, &shards, &enqueued};
    const size_t cells = (size_t)n_sims * n_ss, ncnt = (size_t)n_ss * lw.cw;
    r->counters = static_cast<uint64_t*>(std::calloc(std::max<size_t>(ncnt, 1), sizeof(uint64_t)));
    HostOut ho;
    CUDA_CHECK(cudaSetDevice(shards[0].cx->dev));
    if (bytes) {
        auto take = [&](size_t n) { return static_cast<uint8_t*>(g_pinned.take(n)); };
        ho.unc = r->uncorrected = take(cells * T);
        ho.cor = r->corrected = take(cells * T);
        ho.conv = r->converged = take(cells);
        if (s.report_overall) ho.overall = r->overall = take(cells);
        if (F > 0) ho.fx = r->factor_excluded = take(cells * F);
        if (lw.plan.n_ph > 0) {
            ho.ph_unc = r->posthoc_unc = take(cells * lw.plan.n_ph);
            ho.ph_cor = r->posthoc_cor = take(cells * lw.plan.n_ph);
        }
        if (mle) {
            ho.joint = r->joint = take(cells);
            ho.bh = r->boundary_hit = take(cells);
        }
    }
    if (capture) ho.tsq = r->t_sq = static_cast<double*>(g_pinned.take(cells * T * sizeof(double)));

    // ---- enqueue every device's work, then wait
    enqueued = true;
    for (auto& sh : shards) enqueue_shard(sh, lw, s, n_ss, base_seed, opt.sim_begin, opt, ho, sink);
    r->lower_ms = lower_ms;
    r->enqueue_ms = ms_since(t_enqueue);
    const auto t_wait = std::chrono::steady_clock::now();
    bool cancelled = false;
    if (sink) {
        const uint64_t total = progress_total ? progress_total : n_sims;
        const uint64_t step = std::max<uint64_t>(n_sims / 50, 1);  // batch.rs:449-450
        uint64_t reported = 0;
        auto t_last = std::chrono::steady_clock::now();
        int spins = 0;
        for (;;) {
            bool all = true;
            for (auto& sh : shards) all = shard_finished(sh) && all;
            if (all) break;
            if (!cancelled) {
                uint64_t done_sims = 0;
                for (auto& sh : shards) done_sims += std::min<uint64_t>(sh.cx->mail->progress / sh.units_per_sim, sh.count);
                if (n_sims > 0) done_sims = std::min<uint64_t>(done_sims, n_sims - 1);  // the last report follows the last kernel
                bool stop = false;
                if (progress && done_sims >= reported + step) {
                    reported = done_sims - done_sims % step;
                    stop = !progress(progress_offset + reported, total, user);
                }
                if (!stop && opt.is_cancelled) {
                    const auto now = std::chrono::steady_clock::now();
                    if (now - t_last >= std::chrono::microseconds(100)) {
                        t_last = now;
                        stop = opt.is_cancelled(user) != 0;
                    }
                }
                if (stop) {  // CTAs that have not started yet see the flag and exit; running ones finish their studies
                    cancelled = true;
                    for (auto& sh : shards) {
                        CUDA_CHECK(cudaSetDevice(sh.cx->dev));
                        CUDA_CHECK(cudaMemcpyAsync(sh.cx->d_cancel, &sh.cx->mail->one, sizeof(int), cudaMemcpyHostToDevice, sh.cx->ctl));
                    }
                }
            }
            // a default-size call (1600 sims: ~50 µs of kernels) ends inside the spinning phase; a sleep costs ~100 µs of timer
            // slack, so the thread only starts sleeping once the call has outlived 300 µs
            if (++spins < 64 || std::chrono::steady_clock::now() - t_wait < std::chrono::microseconds(300)) std::this_thread::yield();
            else std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
        if (cancelled)
            for (auto& sh : shards) {
                CUDA_CHECK(cudaSetDevice(sh.cx->dev));
                CUDA_CHECK(cudaStreamSynchronize(sh.cx->ctl));
            }
    } else {
        for (auto& sh : shards) {
            CUDA_CHECK(cudaSetDevice(sh.cx->dev));
            CUDA_CHECK(cudaEventSynchronize(sh.cx->done_compute));
            if (sh.copies) CUDA_CHECK(cudaEventSynchronize(sh.cx->done_copy));
            sh.done = true;
        }
    }
    enqueued = false;  // drained
    r->wait_ms = ms_since(t_wait);
    if (cancelled) throw EngineError(MCP_CANCELLED, "cancelled by host");  // partial results discarded (batch.rs:128-129)
    bool not_psd = false;
    for (auto& sh : shards) {
        CUDA_CHECK(cudaSetDevice(sh.cx->dev));
        CUDA_CHECK(cudaEventElapsedTime(&sh.ms, sh.cx->e0, sh.cx->e1));
        r->kernel_ms = std::max<double>(r->kernel_ms, sh.ms);  // the devices run concurrently
        r->gpu_launches += sh.launches;
        r->n_waves += sh.waves;
        r->h2d_bytes += sh.h2d;
        r->d2h_bytes += sh.d2h;
        not_psd = not_psd || sh.cx->mail->error_flag == 2;
        // the path's only reduction: integer counters per (sample size, test).  The partial sums must reach the host anyway
        // (they ARE the result), one page-locked row block per device, so they are added here rather than by an NCCL
        // all-reduce that would then be followed by the same device→host copy.
        const unsigned long long* hc = sh.cx->h_counters;
        for (size_t i = 0; i < ncnt; i++) r->counters[i] += hc[i];
    }
    if (not_psd) throw EngineError(MCP_CORRELATION_NOT_PSD, "correlation matrix is not positive semi-definite");
    if (progress && !progress(progress_offset + n_sims, progress_total ? progress_total : n_sims, user))
        throw EngineError(MCP_CANCELLED, "cancelled by host");
    guard.ok = true;
    return r;
}

int mcp_run_batch(const char* spec_json, size_t spec_len, const uint32_t* sample_sizes, uint32_t n_ss, uint32_t n_sims,
                  uint64_t base_seed, const McpOptions* options, McpProgressFn progress, void* user, McpBatchResult** out) {
    return guarded([&] {
        if (!out) invalid("out pointer is NULL");
        *out = nullptr;
        Spec s = parse_spec(spec_json, spec_len);
        *out = run_batch_core(s, sample_sizes, n_ss, n_sims, base_seed, options, progress, user);
    });
}

// ------------------------------------------------------------------------------------------------------------------
// Orchestrator entry points: the callers of run_batch, for ONE lowered scenario.

// aggregate_batch (aggregation.rs:61-271) from the per-cell counters the kernels reduced (identical integers to the
// fold over the byte tensors: tests prove COUNTS == aggregate(BYTES)).
static std::vector<orch::PowerResult> power_results_from(const Spec& s, const McpBatchResult* r, const uint32_t* ss) {
    const int T = (int)r->n_targets, F = (int)r->n_factors, C = (int)r->n_posthoc, J = T + C;
    const int off_hc = 2 * T + 3 + J, off_fx = off_hc + J + 1, off_joint = off_fx + F, off_bh = off_joint + 1;
    const int off_pu = off_bh + 1, off_pc = off_pu + C, off_sing = off_pc + C, off_fsep = off_sing + 1, off_pin = off_fsep + F;
    std::vector<orch::PowerResult> out(r->n_sample_sizes);
    for (uint32_t g = 0; g < r->n_sample_sizes; g++) {
        const uint64_t* row = r->counters + (size_t)g * r->counter_width;
        orch::PowerResult& pr = out[g];
        pr.n = ss[g];
        pr.n_sims = r->n_sims;
        pr.target_indices = s.target_indices;
        pr.contrast_pairs = s.contrast_pairs;
        pr.k_unc.assign(row, row + T);
        pr.k_cor.assign(row + T, row + 2 * T);
        pr.conv_count = row[2 * T];
        pr.overall_count = row[2 * T + 1];
        pr.has_overall = s.report_overall;
        pr.hist_unc.assign(row + 2 * T + 2, row + 2 * T + 3 + J);
        pr.hist_cor.assign(row + off_hc, row + off_hc + J + 1);
        pr.excl_sparse.assign(row + off_fx, row + off_fx + F);
        pr.excl_sep.assign(row + off_fsep, row + off_fsep + F);
        pr.estimator = s.estimator == E_OLS ? 0 : (s.estimator == E_GLM ? 1 : 2);
        pr.joint_count = row[off_joint];
        pr.boundary_hits = row[off_bh];
        pr.singular_count = row[off_sing];
        pr.n_components = s.has_cluster ? 1 + s.n_slopes + s.n_extra : 0;
        for (int k = 0; k < MCP_MAX_PIN; k++) pr.pin_counts[k] = row[off_pin + k];
        size_t o = 0;
        for (size_t b = 0; b < s.posthoc_block_ncontrasts.size(); b++) {
            orch::PosthocPower ph;
            ph.n_levels = (uint64_t)s.posthoc_block_levels[b];
            const int m = s.posthoc_block_ncontrasts[b];
            ph.k_unc.assign(row + off_pu + o, row + off_pu + o + m);
            ph.k_cor.assign(row + off_pc + o, row + off_pc + o + m);
            o += (size_t)m;
            pr.posthoc.push_back(ph);
        }
        if (r->boundary_hit)
            for (uint32_t i = 0; i < r->n_sims; i++) pr.boundary_hit.push_back(r->boundary_hit[(size_t)i * r->n_sample_sizes + g] & 3);  // the code; bits 2.. are the pinned components
    }
    return out;
}

static std::vector<std::string> factor_preflight_warnings(const Spec& s, uint64_t n);

static char* dup_json(const std::string& s) {
    char* p = static_cast<char*>(std::malloc(s.size() + 1));
    std::memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

// find_power for one lowered scenario: run_batch + aggregate_batch → JSON array of PowerResult, one per sample size.
// `report` carries the target / contrast indices the results are labelled with (the design_test term space when the
// scenario came from a contract, find_power.rs:353-358; the spec's own indices otherwise).
static std::string find_power_json(const Spec& s, const Spec& report, const uint32_t* sample_sizes, uint32_t n_ss, uint32_t n_sims,
                                   uint64_t base_seed, const McpOptions* options, McpProgressFn progress, void* user,
                                   const std::vector<std::string>& warnings = {}, uint64_t progress_offset = 0, uint64_t progress_total = 0) {
    McpBatchResult* r = run_batch_core(s, sample_sizes, n_ss, n_sims, base_seed, options, progress, user, progress_offset, progress_total);
    struct Free { McpBatchResult* r; ~Free() { mcp_free_result(r); } } fr{r};
    std::vector<orch::PowerResult> per_n = power_results_from(report, r, sample_sizes);
    std::string o = "[";
    for (size_t i = 0; i < per_n.size(); i++) {
        if (i) o += ',';
        per_n[i].grid_warnings.insert(per_n[i].grid_warnings.end(), warnings.begin(), warnings.end());
        // pre-run sparse-factor advisory (find_power.rs:364-368; fixed-allocation factors only)
        for (auto& w : factor_preflight_warnings(s, sample_sizes[i])) per_n[i].grid_warnings.push_back(w);
        orch::power_result_json(o, per_n[i]);
    }
    o += ']';
    return o;
}

int mcp_find_power(const char* spec_json, size_t spec_len, const uint32_t* sample_sizes, uint32_t n_ss, uint32_t n_sims,
                   uint64_t base_seed, const McpOptions* options, McpProgressFn progress, void* user, char** json_out) {
    return guarded([&] {
        if (!json_out) invalid("json_out pointer is NULL");
        *json_out = nullptr;
        Spec s = parse_spec(spec_json, spec_len);
        *json_out = dup_json(find_power_json(s, s, sample_sizes, n_ss, n_sims, base_seed, options, progress, user));
    });
}

// resolve_cluster_grid_params (find_sample_size.rs:24-79) over ALL scenarios of a call: one shared atom (scenarios that
// disagree are an error, OrchestratorError::MixedClusterAtoms), grid floor and pre-grid warnings from the first clustered
// scenario; limits from configs/config.json:27-28.
static void cluster_grid_params(const std::vector<const Spec*>& specs, uint64_t& atom, uint64_t& hard_min, std::vector<std::string>& warnings) {
    const uint64_t min_rows = 2, min_clusters = 5;
    atom = 1;
    hard_min = 1;
    const Spec* cl = nullptr;
    for (const Spec* s : specs) {
        if (!s->has_cluster) continue;
        const uint64_t a = (uint64_t)std::max(s->sizing_value, 1);
        if (!cl) cl = s;
        else if ((uint64_t)std::max(cl->sizing_value, 1) != a)
            invalid("scenarios in this call disagree on the cluster atom (" + std::to_string(std::max(cl->sizing_value, 1)) + " vs " + std::to_string(a) +
                    "); a single shared grid cannot be a multiple of every atom");
    }
    if (!cl) return;
    const uint64_t v = (uint64_t)std::max(cl->sizing_value, 1);
    atom = v;
    if (!cl->sizing_fixed_size) {
        if (v < min_clusters) warnings.push_back("only " + std::to_string(v) + " clusters (< " + std::to_string(min_clusters) + "); τ² estimates may be unstable");
        hard_min = v * min_rows;
    } else {
        if (v < min_rows)
            invalid("FixedSize cluster regime needs cluster_size >= " + std::to_string(min_rows) + " (no within-cluster information below that); got " +
                    std::to_string(v));
        hard_min = min_clusters * v;
    }
}

// Level counts of the deterministic largest-remainder walk after n rows (fixed_allocation_counts, data_gen.rs:966-972).
static std::vector<uint32_t> fixed_allocation_counts(const double* probs, int nl, uint64_t n) {
    std::vector<uint32_t> counts(nl, 0);
    if (n == 0) return counts;
    const std::vector<uint8_t> lv = exact_walk_levels(probs, nl, (int)std::min<uint64_t>(n, 0x7fffffffu));
    for (uint8_t l : lv) counts[l]++;
    return counts;
}

static bool factor_is_sampled(const Spec& s, size_t f) {
    const int fs = f < s.factor_sampled.size() ? s.factor_sampled[f] : -1;
    return fs < 0 ? s.sampled_factor_proportions : fs != 0;
}

// factor_preflight_warnings (find_power.rs:67-116): fixed-allocation factors whose sparsest level stays under the minimum at n.
static std::vector<std::string> factor_preflight_warnings(const Spec& s, uint64_t n) {
    std::vector<std::string> out;
    const uint32_t k_min = (uint32_t)std::max(s.min_level_count, 0);
    if (k_min == 0 || !s.upload_data.empty()) return out;
    size_t off = 0;
    for (size_t f = 0; f < s.factor_n_levels.size(); f++) {
        const int l = std::max(s.factor_n_levels[f], 0);
        if (!factor_is_sampled(s, f) && l > 0) {
            const std::vector<uint32_t> counts = fixed_allocation_counts(s.factor_proportions.data() + off, l, n);
            size_t lvl = 0;
            for (size_t i = 1; i < counts.size(); i++)
                if (counts[i] < counts[lvl]) lvl = i;  // min_by_key: the first minimum
            if (counts[lvl] < k_min)
                out.push_back("factor " + std::to_string(f + 1) + ": level " + std::to_string(lvl + 1) + " receives " + std::to_string(counts[lvl]) + " of " +
                              std::to_string(n) + " observations (minimum " + std::to_string(k_min) + "); the factor is excluded from every simulation at N=" +
                              std::to_string(n) + " and its effects report power 0 — increase N or this level's proportion");
        }
        off += (size_t)l;
    }
    return out;
}

// factor_preflight_sample_size (find_power.rs:122-170): factors sparse at the grid's lower bound, and the N at which they clear it.
static std::vector<std::string> factor_preflight_sample_size(const Spec& s, const std::vector<uint64_t>& grid) {
    std::vector<std::string> out;
    const uint32_t k_min = (uint32_t)std::max(s.min_level_count, 0);
    if (grid.empty() || k_min == 0 || !s.upload_data.empty()) return out;
    const uint64_t lo = grid.front(), hi = grid.back();
    size_t off = 0;
    for (size_t f = 0; f < s.factor_n_levels.size(); f++) {
        const int l = std::max(s.factor_n_levels[f], 0);
        if (!factor_is_sampled(s, f) && l > 0) {
            const double* probs = s.factor_proportions.data() + off;
            const std::vector<uint32_t> at_lo = fixed_allocation_counts(probs, l, lo);
            if (std::any_of(at_lo.begin(), at_lo.end(), [&](uint32_t c) { return c < k_min; })) {
                // min_inclusion_n (data_gen.rs:977-986): one forward walk up to the grid ceiling
                const std::vector<uint8_t> lv = exact_walk_levels(probs, l, (int)std::min<uint64_t>(hi, 0x7fffffffu));
                std::vector<uint32_t> run(l, 0);
                uint64_t n_inc = 0;
                for (size_t i = 0; i < lv.size() && !n_inc; i++) {
                    run[lv[i]]++;
                    if (std::all_of(run.begin(), run.end(), [&](uint32_t c) { return c >= k_min; })) n_inc = i + 1;
                }
                if (n_inc)
                    out.push_back("factor " + std::to_string(f + 1) + ": levels reach the " + std::to_string(k_min) + "-observation minimum only at N >= " +
                                  std::to_string(n_inc) + "; grid points below that exclude the factor and report power 0 for its effects");
                else
                    out.push_back("factor " + std::to_string(f + 1) + ": a level stays under the " + std::to_string(k_min) +
                                  "-observation minimum across the whole [" + std::to_string(lo) + ", " + std::to_string(hi) +
                                  "] range; the factor is excluded everywhere and its effects report power 0");
            }
        }
        off += (size_t)l;
    }
    return out;
}

// The shared N grid of a find_sample_size call (find_sample_size.rs:99-107): cluster guards over every scenario, then build_grid.
static std::vector<uint64_t> shared_grid(const std::vector<const Spec*>& specs, uint64_t from_size, uint64_t to_size, uint64_t by, int32_t by_auto,
                                         int32_t log_grid, uint64_t& atom, std::vector<std::string>& warnings) {
    uint64_t hard_min = 1;
    cluster_grid_params(specs, atom, hard_min, warnings);
    std::vector<uint64_t> grid;
    const orch::GridError ge = orch::build_grid(from_size, to_size, by, by_auto != 0, log_grid != 0, atom, hard_min, grid, warnings);
    if (ge == orch::GRID_INVALID_BOUNDS) invalid("sample-size bounds: need 1 <= from <= to");
    if (ge == orch::GRID_EMPTY) invalid("no multiple of the cluster atom lies inside the sample-size bounds");
    if (ge == orch::GRID_SINGLE_POINT) invalid("the sample-size bounds leave a single grid point after cluster snapping");
    if (ge != orch::GRID_OK) invalid("cluster atom must be positive");
    if (grid.size() > 65535) invalid("the sample-size grid has " + std::to_string(grid.size()) + " points (at most 65535 per call)");
    if (grid.back() > 0xffffffffull) invalid("sample sizes above 2^32 - 1 are not supported");
    return grid;
}

// find_sample_size for one lowered scenario (find_sample_size.rs:83-257): cluster-aware grid, one run_batch over the whole
// grid (nested row prefixes of the same studies), per-N PowerResult fold, grid-empirical first-N and model-based
// crossing fits per target / per joint count / for the overall test.  JSON object = SampleSizeResult.
static std::string find_sample_size_json(const Spec& s, const Spec& report, const std::vector<uint64_t>& grid, uint64_t atom,
                                         std::vector<std::string> warnings, double target_power, uint32_t n_sims, uint64_t base_seed,
                                         const McpOptions* options, McpProgressFn progress, void* user, uint64_t progress_offset = 0,
                                         uint64_t progress_total = 0) {
    // per-scenario preflight: the allocation mode is a scenario knob, the grid is shared (find_sample_size.rs:142-146)
    for (auto& w : factor_preflight_sample_size(s, grid)) warnings.push_back(w);
    std::vector<uint32_t> ss(grid.begin(), grid.end());
    McpBatchResult* r = run_batch_core(s, ss.data(), (uint32_t)ss.size(), n_sims, base_seed, options, progress, user, progress_offset,
                                       progress_total);
    struct Free { McpBatchResult* r; ~Free() { mcp_free_result(r); } } fr{r};
    const std::vector<orch::PowerResult> per_n = power_results_from(report, r, ss.data());
    const orch::SampleSizeDerivations d = orch::derive_sample_size_outputs(per_n, grid, target_power, atom);
    std::string o = "{\"grid_or_trace\":[";
    for (size_t i = 0; i < per_n.size(); i++) {
        if (i) o += ',';
        orch::power_result_json(o, per_n[i]);
    }
    o += "],\"first_achieved\":";
    orch::jarr(o, d.first_achieved, [](std::string& t, int64_t v) { orch::jopt(t, v); });
    o += ",\"first_joint_achieved\":";
    orch::jarr(o, d.first_joint_achieved, [](std::string& t, int64_t v) { orch::jopt(t, v); });
    o += ",\"fitted\":";
    orch::jarr(o, d.fitted, [](std::string& t, const orch::CrossingFit& f) { orch::crossing_fit_json(t, f); });
    o += ",\"fitted_joint\":";
    orch::jarr(o, d.fitted_joint, [](std::string& t, const orch::CrossingFit& f) { orch::crossing_fit_json(t, f); });
    o += ",\"first_overall_achieved\":";
    orch::jopt(o, d.has_overall ? d.first_overall_achieved : -1);
    o += ",\"fitted_overall\":";
    if (d.has_overall) orch::crossing_fit_json(o, d.fitted_overall);
    else o += "null";
    o += ",\"cluster_atom\":" + std::to_string(atom) + ",\"target_power\":";
    orch::jnum(o, target_power);
    o += ",\"grid_warnings\":";
    orch::jarr(o, warnings, [](std::string& t, const std::string& w) { orch::jstr(t, w); });
    o += ",\"kernel_ms\":";
    orch::jnum(o, r->kernel_ms);
    o += '}';
    return o;
}

int mcp_find_sample_size(const char* spec_json, size_t spec_len, uint64_t from_size, uint64_t to_size, uint64_t by, int32_t by_auto,
                         int32_t log_grid, double target_power, uint32_t n_sims, uint64_t base_seed, const McpOptions* options,
                         McpProgressFn progress, void* user, char** json_out) {
    return guarded([&] {
        if (!json_out) invalid("json_out pointer is NULL");
        *json_out = nullptr;
        Spec s = parse_spec(spec_json, spec_len);
        uint64_t atom = 1;
        std::vector<std::string> warnings;
        const std::vector<uint64_t> grid = shared_grid({&s}, from_size, to_size, by, by_auto, log_grid, atom, warnings);
        *json_out = dup_json(find_sample_size_json(s, s, grid, atom, warnings, target_power, n_sims, base_seed, options, progress, user));
    });
}

// ------------------------------------------------------------------------------------------------------------------
// Contract entry points: the hosts' own payload — the `Vec<SimulationContract>` blob
// engine-py / engine-r forward verbatim — lowered by contract.hpp and run through the same paths as above.

struct LoweredScenario {
    std::string label;
    std::string spec_json;
    Spec spec, report;  // report = spec with the host-report (design_test term) target / contrast indices
    contract::Lowering low;
};

// lower_contracts (find_power.rs:251-288): validate + lower every contract, force factor_min_level_count to the
// configured limit (configs/config.json:31), keep the report-space twins of the target indices.
static std::vector<LoweredScenario> lower_contracts(const uint8_t* bytes, size_t len, int32_t format) {
    if (!bytes) invalid("contracts pointer is NULL");
    const std::vector<JsonValue> raw = contract::decode_contracts(bytes, len, format);
    std::vector<LoweredScenario> out;
    for (size_t idx = 0; idx < raw.size(); idx++) {
        LoweredScenario sc;
        try {
            const contract::Contract c = contract::parse_contract(raw[idx]);
            if (const JsonValue* nm = c.scenario.get("name")) sc.label = nm->str;
            sc.low = contract::lower_contract(c, /*factor_min_level_count: limits.factor_min_level_count*/ 5);
        } catch (const EngineError& e) {
            throw EngineError(e.code, "invalid scenarios input: contract " + std::to_string(idx) + " (\"" + sc.label + "\"): " + e.what());
        }
        sc.spec_json = sc.low.spec_json;
        sc.spec = parse_spec(sc.spec_json.c_str(), sc.spec_json.size());
        sc.report = sc.spec;
        sc.report.target_indices.assign(sc.low.report_targets.begin(), sc.low.report_targets.end());
        sc.report.contrast_pairs.clear();
        for (auto& p : sc.low.report_contrasts) sc.report.contrast_pairs.push_back({(int)p.first, (int)p.second});
        out.push_back(std::move(sc));
    }
    return out;
}

int mcp_lower_contracts(const uint8_t* contracts, size_t len, int32_t format, char** json_out) {
    return guarded([&] {
        if (!json_out) invalid("json_out pointer is NULL");
        *json_out = nullptr;
        const std::vector<LoweredScenario> scs = lower_contracts(contracts, len, format);
        std::string o = "[";
        for (size_t i = 0; i < scs.size(); i++) {
            if (i) o += ',';
            o += "{\"label\":";
            orch::jstr(o, scs[i].label);
            o += ",\"spec\":" + scs[i].spec_json + ",\"report_target_indices\":";
            orch::jarr(o, scs[i].low.report_targets, [](std::string& t, uint32_t v) { t += std::to_string(v); });
            o += ",\"report_contrast_pairs\":";
            orch::jarr(o, scs[i].low.report_contrasts,
                       [](std::string& t, const std::pair<uint32_t, uint32_t>& p) { t += "[" + std::to_string(p.first) + "," + std::to_string(p.second) + "]"; });
            o += '}';
        }
        o += ']';
        *json_out = dup_json(o);
    });
}

// find_power over contracts (find_power.rs:301-392): N snapped to the cluster atom (snap_single_n, find_power.rs:25-58),
// every scenario run with the same call seed splitmix64(user_seed) — scenarios of one call are paired runs — and the
// results labelled in report space.  JSON = ScenarioResult<PowerResult>: {"scenarios": [[label, PowerResult], ...]}.
int mcp_find_power_contracts(const uint8_t* contracts, size_t len, int32_t format, uint64_t sample_size, uint32_t n_sims, uint64_t user_seed,
                             const McpOptions* options, McpProgressFn progress, void* user, char** json_out) {
    return guarded([&] {
        if (!json_out) invalid("json_out pointer is NULL");
        *json_out = nullptr;
        const std::vector<LoweredScenario> scs = lower_contracts(contracts, len, format);
        uint64_t n = sample_size;
        std::vector<std::string> warnings;
        for (auto& sc : scs)
            if (sc.spec.has_cluster) {
                const uint64_t a = (uint64_t)std::max(sc.spec.sizing_value, 1);
                if (a > 1) {
                    uint64_t snapped = (n / a) * a;
                    if (snapped == 0) snapped = a;
                    if (sc.spec.sizing_fixed_size && snapped / a < 5)
                        invalid("sample size " + std::to_string(snapped) + " gives " + std::to_string(snapped / a) + " clusters of size " + std::to_string(a) +
                                " (minimum 5)");
                    if (snapped != n)
                        warnings.push_back("sample_size " + std::to_string(n) + " not a multiple of the cluster atom " + std::to_string(a) + "; using " +
                                           std::to_string(snapped));
                    n = snapped;
                }
                break;
            }
        if (n == 0 || n > 0xffffffffull) invalid("sample_size out of range");
        const uint32_t n32 = (uint32_t)n;
        const uint64_t call_seed = mcp_splitmix64(user_seed);
        std::string o = "{\"scenarios\":[";
        for (size_t i = 0; i < scs.size(); i++) {
            if (i) o += ',';
            o += '[';
            orch::jstr(o, scs[i].label);
            // the sink sees one range over the whole call: scenario i starts at i * n_sims of n_sims * scenarios (find_power.rs:311-337)
            const std::string arr = find_power_json(scs[i].spec, scs[i].report, &n32, 1, n_sims, call_seed, options, progress, user, warnings,
                                                    (uint64_t)i * n_sims, (uint64_t)n_sims * scs.size());
            o += ',' + arr.substr(1, arr.size() - 2) + ']';
        }
        o += "]}";
        *json_out = dup_json(o);
    });
}

// find_sample_size over contracts (find_sample_size.rs:83-257): {"scenarios": [[label, SampleSizeResult], ...]}.
int mcp_find_sample_size_contracts(const uint8_t* contracts, size_t len, int32_t format, uint64_t from_size, uint64_t to_size, uint64_t by,
                                   int32_t by_auto, int32_t log_grid, double target_power, uint32_t n_sims, uint64_t user_seed,
                                   const McpOptions* options, McpProgressFn progress, void* user, char** json_out) {
    return guarded([&] {
        if (!json_out) invalid("json_out pointer is NULL");
        *json_out = nullptr;
        const std::vector<LoweredScenario> scs = lower_contracts(contracts, len, format);
        const uint64_t call_seed = mcp_splitmix64(user_seed);
        // ONE grid for the call: atom, floor and cluster warnings are resolved over every scenario (find_sample_size.rs:24-107)
        std::vector<const Spec*> specs;
        for (auto& sc : scs) specs.push_back(&sc.spec);
        uint64_t atom = 1;
        std::vector<std::string> warnings;
        const std::vector<uint64_t> grid = shared_grid(specs, from_size, to_size, by, by_auto, log_grid, atom, warnings);
        std::string o = "{\"scenarios\":[";
        for (size_t i = 0; i < scs.size(); i++) {
            if (i) o += ',';
            o += '[';
            orch::jstr(o, scs[i].label);
            // progress in draws: scenario i starts at i * n_sims of n_sims * scenarios (the reference counts fits = draws x grid
            // points, find_sample_size.rs:109-113; the ratio current / total is the same)
            o += ',' + find_sample_size_json(scs[i].spec, scs[i].report, grid, atom, warnings, target_power, n_sims, call_seed, options, progress,
                                             user, (uint64_t)i * n_sims, (uint64_t)n_sims * scs.size()) + ']';
        }
        o += "]}";
        *json_out = dup_json(o);
    });
}

void mcp_free_string(char* s) { std::free(s); }

// Pure helpers, exported so the reference's known-answer tests (fit.rs:262-491, grid.rs:393-638, aggregation.rs)
// run against this implementation.
int mcp_fit_crossing(const uint64_t* grid, const uint64_t* counts, uint32_t n_points, uint64_t n_sims, double target_power,
                     uint64_t atom, McpCrossingFit* out) {
    return guarded([&] {
        if (!out) invalid("out pointer is NULL");
        const orch::CrossingFit f = orch::fit_crossing(std::vector<uint64_t>(grid, grid + n_points),
                                                       std::vector<uint64_t>(counts, counts + n_points), n_sims, target_power, atom);
        out->status = f.status;
        out->n_star = f.n_star;
        out->n_achievable = f.n_achievable;
        out->has_ci_lo = f.has_ci_lo;
        out->has_ci_hi = f.has_ci_hi;
        out->ci_lo = f.ci_lo;
        out->ci_hi = f.ci_hi;
        out->n_min = f.n_min;
        out->has_n_approx = f.has_n_approx;
        out->n_approx = f.n_approx;
        out->max_violation = f.max_violation;
    });
}

int mcp_build_grid(uint64_t from_size, uint64_t to_size, uint64_t by, int32_t by_auto, int32_t log_grid, uint64_t atom,
                   uint64_t hard_min, uint64_t* grid_out, uint32_t capacity, uint32_t* n_out, uint32_t* n_warnings) {
    return guarded([&] {
        std::vector<uint64_t> grid;
        std::vector<std::string> warnings;
        const orch::GridError ge = orch::build_grid(from_size, to_size, by, by_auto != 0, log_grid != 0, atom, hard_min, grid, warnings);
        if (ge != orch::GRID_OK) invalid("build_grid: error " + std::to_string((int)ge));
        if (grid.size() > capacity) invalid("build_grid: capacity too small");
        std::copy(grid.begin(), grid.end(), grid_out);
        if (n_out) *n_out = (uint32_t)grid.size();
        if (n_warnings) *n_warnings = (uint32_t)warnings.size();
    });
}

void mcp_wilson_ci(uint64_t k, uint64_t n, double* lo, double* hi) {
    const orch::Ci c = orch::wilson_ci(k, n);
    *lo = c.lo;
    *hi = c.hi;
}

int mcp_generate_sim(const char* spec_json, size_t spec_len, uint32_t max_n, uint64_t base_seed, uint64_t sim_id, int32_t device,
                     float* x_colmajor, float* y) {
    return guarded([&] {
        Spec s = parse_spec(spec_json, spec_len);
        validate_spec(s, nullptr, 0);
        check_supported(s, /*generation_only=*/true);
        if (max_n == 0) invalid("max_n must be > 0");
        DeviceCtx& cx = ctx_for(device);
        std::lock_guard<std::mutex> lock(cx.mu);
        const cudaStream_t st = cx.compute;
        Lowered lw;
        lower(s, &max_n, 1, lw, /*thresholds=*/false);
        const int P = lw.P;
        void* dblob = cx.tables.get(lw.blob.size());
        CUDA_CHECK(cudaMemcpyAsync(dblob, lw.blob.data(), lw.blob.size(), cudaMemcpyHostToDevice, st));
        bind_tables(lw, lw.plan, dblob);
        Carver oc(cx.outputs.get(Carver::pad((size_t)max_n * P * sizeof(float)) + Carver::pad((size_t)max_n * sizeof(float))));
        float* dx = oc.take<float>((size_t)max_n * P);
        float* dy = oc.take<float>(max_n);
        CUDA_CHECK(cudaMemsetAsync(cx.d_prog, 0, 256, st));
        lw.plan.error_flag = cx.d_error;
        if (lw.plan.n_between > 0) {
            lw.plan.between_stride = 1;
            lw.plan.between_tab = static_cast<float*>(cx.between.get((size_t)max_n * lw.plan.n_between * sizeof(float)));
        }
        {
            cudaError_t err = cudaSuccess;
            bool handled = false;
#define X(id) if (!handled) handled = tu##id##_dump(P, !lw.plan.optimistic, lw.plan, base_seed, sim_id, dx, dy, st, &err);
            MCP_TU_LIST(X)
#undef X
            if (!handled) unsupported("n_predictors = " + std::to_string(P));
            CUDA_CHECK(err);
        }
        // the caller's buffers are pageable: synchronous copies, ordered behind the kernel by the stream synchronisation
        CUDA_CHECK(cudaStreamSynchronize(st));
        CUDA_CHECK(cudaMemcpy(x_colmajor, dx, (size_t)max_n * lw.plan.P * sizeof(float), cudaMemcpyDeviceToHost));  // the real columns
        CUDA_CHECK(cudaMemcpy(y, dy, (size_t)max_n * sizeof(float), cudaMemcpyDeviceToHost));
        int err_flag = 0;
        CUDA_CHECK(cudaMemcpy(&err_flag, cx.d_error, sizeof(int), cudaMemcpyDeviceToHost));
        if (err_flag == 2) throw EngineError(MCP_CORRELATION_NOT_PSD, "correlation matrix is not positive semi-definite");
    });
}

int mcp_fit_data(const double* x_colmajor, const double* y, uint32_t n, uint32_t p, int32_t device, double* out) {
    return guarded([&] {
        if (p < 1 || p > MCP_MAX_TEMPLATE_P) unsupported("p outside [1, " + std::to_string(MCP_MAX_TEMPLATE_P) + "]");
        if (n <= p) invalid("df = N - P < 1");
        DeviceCtx& cx = ctx_for(device);
        std::lock_guard<std::mutex> lock(cx.mu);
        const cudaStream_t st = cx.compute;
        Carver oc(cx.outputs.get(Carver::pad((size_t)n * p * sizeof(double)) + Carver::pad((size_t)n * sizeof(double)) + Carver::pad((3 * p + 4) * sizeof(double))));
        double* dx = oc.take<double>((size_t)n * p);
        double* dy = oc.take<double>(n);
        double* dout = oc.take<double>(3 * p + 4);
        CUDA_CHECK(cudaMemcpyAsync(dx, x_colmajor, (size_t)n * p * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(dy, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
        {
            cudaError_t err = cudaSuccess;
            bool handled = false;
#define X(id) if (!handled) handled = tu##id##_fit_data((int)p, dx, dy, (int)n, dout, st, &err);
            MCP_TU_LIST(X)
#undef X
            if (!handled) unsupported("p = " + std::to_string(p));
            CUDA_CHECK(err);
        }
        CUDA_CHECK(cudaStreamSynchronize(st));
        CUDA_CHECK(cudaMemcpy(out, dout, (3 * p + 4) * sizeof(double), cudaMemcpyDeviceToHost));
    });
}

int mcp_fit_provided_data(const char* spec_json, size_t spec_len, const double* design_colmajor, uint32_t nrow, uint32_t ncol,
                          const double* outcome, const uint32_t* cluster_ids, int32_t device, double* betas, double* se, double* statistic,
                          McpFitInfo* info) {
    return guarded([&] {
        if (!design_colmajor || !outcome || !betas || !se || !statistic || !info) invalid("NULL pointer argument");
        Spec s = parse_spec(spec_json, spec_len);
        validate_spec(s, nullptr, 0);
        check_supported(s);
        const int P = s.P();
        if ((int)ncol != P) invalid("design has " + std::to_string(ncol) + " columns, the spec " + std::to_string(P) + " predictors");
        if (nrow <= ncol) invalid("df = N - P < 1");
        const int nt = (int)s.target_indices.size();
        std::memset(info, 0, sizeof *info);
        info->n_targets = nt;
        info->df_resid = (double)((long)nrow - (long)P);
        {  // CritValueTable::build with the pipeline's arguments (introspect.rs:535-548)
            double unc = 0.0;
            crit_row(s, nrow, P, &unc, nullptr, nullptr, nullptr);
            info->crit_sq_uncorrected = unc;
        }
        const double nan = crit::kNaN;
        info->sigma_sq_hat = info->tau_sq_hat = info->deviance = info->deviance_null = nan;
        for (int a = 0; a < 2; a++) info->slope_var_hat[a] = nan;
        for (int a = 0; a < 3; a++) info->slope_corr_hat[a] = nan;
        for (int a = 0; a < 2; a++) info->extra_var_hat[a] = nan;
        DeviceCtx& cx = ctx_for(device);
        std::lock_guard<std::mutex> lock(cx.mu);
        const cudaStream_t st = cx.compute;
        std::vector<double> all((size_t)3 * P + 8, 0.0);
        if (s.estimator == E_OLS) {
            Carver oc(cx.outputs.get(Carver::pad((size_t)nrow * P * sizeof(double)) + Carver::pad((size_t)nrow * sizeof(double)) +
                                     Carver::pad((3 * P + 4) * sizeof(double))));
            double* dx = oc.take<double>((size_t)nrow * P);
            double* dy = oc.take<double>(nrow);
            double* dout = oc.take<double>(3 * P + 4);
            CUDA_CHECK(cudaMemcpyAsync(dx, design_colmajor, (size_t)nrow * P * sizeof(double), cudaMemcpyHostToDevice, st));
            CUDA_CHECK(cudaMemcpyAsync(dy, outcome, (size_t)nrow * sizeof(double), cudaMemcpyHostToDevice, st));
            cudaError_t err = cudaSuccess;
            bool handled = false;
#define X(id) if (!handled) handled = tu##id##_fit_data(P, dx, dy, (int)nrow, dout, st, &err);
            MCP_TU_LIST(X)
#undef X
            if (!handled) unsupported("p = " + std::to_string(P));
            CUDA_CHECK(err);
            CUDA_CHECK(cudaStreamSynchronize(st));
            std::vector<double> o(3 * P + 4);
            CUDA_CHECK(cudaMemcpy(o.data(), dout, o.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (int j = 0; j < P; j++) betas[j] = o[j];
            for (int t = 0; t < nt; t++) {
                se[t] = o[P + s.target_indices[t]];
                statistic[t] = std::sqrt(o[2 * P + s.target_indices[t]]);
            }
            info->sigma_sq_hat = o[3 * P];
            info->converged = o[3 * P + 3] != 0.0;
            return;
        }
        if (P > MCP_MAX_TEMPLATE_P) unsupported("fit_provided_data for n_predictors > " + std::to_string(MCP_MAX_TEMPLATE_P));
        // GLM / MLE: the f32 data plane [row][x_1 .. x_{P-1}, y]; the kernels take x_0 = 1 like the generated designs
        std::vector<float> plane((size_t)nrow * P);
        for (uint32_t i = 0; i < nrow; i++) {
            if ((float)design_colmajor[i] != 1.0f) unsupported("a design whose first column is not the intercept (all ones)");
            for (int j = 1; j < P; j++) plane[(size_t)i * P + (j - 1)] = (float)design_colmajor[(size_t)j * nrow + i];
            plane[(size_t)i * P + (P - 1)] = (float)outcome[i];
        }
        const bool mle = s.estimator == E_MLE;
        size_t n_clusters = 0;
        if (mle) {
            const uint32_t cs = (uint32_t)std::max(s.sizing_value, 1);
            n_clusters = s.sizing_fixed_size ? ((size_t)nrow + cs - 1) / cs : cs;
            if (cluster_ids)
                for (uint32_t i = 0; i < nrow; i++)
                    if (cluster_ids[i] != (s.sizing_fixed_size ? i / cs : i % cs))
                        unsupported("cluster ids that do not follow the spec's cluster sizing (row " + std::to_string(i) + ")");
        }
        Lowered lw;
        lower(s, &nrow, 1, lw, /*thresholds=*/false);
        void* dblob = cx.tables.get(lw.blob.size());
        CUDA_CHECK(cudaMemcpyAsync(dblob, lw.blob.data(), lw.blob.size(), cudaMemcpyHostToDevice, st));
        bind_tables(lw, lw.plan, dblob);
        CUDA_CHECK(cudaMemsetAsync(cx.d_prog, 0, 256, st));
        lw.plan.error_flag = cx.d_error;
        size_t n_csum = n_clusters * (size_t)(P + 1) * (size_t)(1 + lw.plan.n_slopes);
        if (mle && lw.plan.n_extra > 0) {  // GenLayout (kernels_lmm.cuh): statistics + work
            const size_t K = n_clusters, q0 = 1 + (size_t)lw.plan.n_slopes, Qe = (size_t)lw.plan.extra_levels[0] + (size_t)lw.plan.extra_levels[1];
            n_csum = 2 * (K * q0 * (q0 + 1) / 2 + Qe * K * q0 + Qe * (Qe + 1) / 2 + (K * q0 + Qe) * (size_t)(P + 1));
        }
        const size_t n_out = (size_t)3 * P + 14;
        Carver oc(cx.outputs.get(Carver::pad(plane.size() * sizeof(float)) + Carver::pad(n_csum * sizeof(double)) + Carver::pad(n_out * sizeof(double))));
        float* dplane = oc.take<float>(plane.size());
        double* dcsum = oc.take<double>(n_csum);
        double* dout = oc.take<double>(n_out);
        CUDA_CHECK(cudaMemcpyAsync(dplane, plane.data(), plane.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        if (n_csum) CUDA_CHECK(cudaMemsetAsync(dcsum, 0, n_csum * sizeof(double), st));
        {
            cudaError_t err = cudaSuccess;
            bool handled = false;
            const int kind = mle ? KIND_LMM : KIND_GLM;
#define X(id) if (!handled) handled = tu##id##_fit_spec(P, kind, lw.plan, dplane, (int)nrow, dcsum, dout, st, &err);
            MCP_TU_LIST(X)
#undef X
            if (!handled) unsupported("p = " + std::to_string(P));
            CUDA_CHECK(err);
        }
        CUDA_CHECK(cudaStreamSynchronize(st));
        std::vector<double> o(n_out);
        CUDA_CHECK(cudaMemcpy(o.data(), dout, n_out * sizeof(double), cudaMemcpyDeviceToHost));
        for (int j = 0; j < P; j++) betas[j] = o[j];
        for (int t = 0; t < nt; t++) {
            se[t] = std::sqrt(o[P + s.target_indices[t]]);
            statistic[t] = std::sqrt(o[2 * P + s.target_indices[t]]);
        }
        info->converged = o[3 * P] != 0.0;
        info->boundary_hit = (int32_t)o[3 * P + 1] & 3;
        info->pinned_components = (int32_t)o[3 * P + 1] >> 2;
        info->iterations = (int32_t)o[3 * P + 2];
        info->sigma_sq_hat = o[3 * P + 3];
        info->tau_sq_hat = o[3 * P + 4];
        info->deviance = o[3 * P + 5];
        info->deviance_null = o[3 * P + 6];
        for (int a = 0; a < 2; a++) info->slope_var_hat[a] = o[3 * P + 7 + a];
        for (int a = 0; a < 3; a++) info->slope_corr_hat[a] = o[3 * P + 9 + a];
        for (int a = 0; a < 2; a++) info->extra_var_hat[a] = o[3 * P + 12 + a];
    });
}

int mcp_crit_table(const char* spec_json, size_t spec_len, const uint32_t* sample_sizes, uint32_t n_ss, uint32_t n_predictors,
                   double* t_unc_sq, double* corr_sq, double* posthoc_sq, double* overall) {
    return guarded([&] {
        Spec s = parse_spec(spec_json, spec_len);
        const int T = s.T();
        for (uint32_t g = 0; g < n_ss; g++)
            crit_row(s, sample_sizes[g], (int)n_predictors, t_unc_sq ? t_unc_sq + g : nullptr, corr_sq ? corr_sq + (size_t)g * T : nullptr,
                     posthoc_sq ? posthoc_sq + g : nullptr, overall ? overall + g : nullptr);
    });
}

int mcp_pipe_bench(int32_t kind, int32_t device, double* thread_ops_per_s) {
    return guarded([&] {
        DeviceCtx& cx = ctx_for(device);
        std::lock_guard<std::mutex> lock(cx.mu);
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, cx.dev));
        double* sink = static_cast<double*>(cx.small.get(256));
        const int iters = 1 << 14, blocks = prop.multiProcessorCount * 8, threads = 256;
        cudaEvent_t e0, e1;
        CUDA_CHECK(cudaEventCreate(&e0));
        CUDA_CHECK(cudaEventCreate(&e1));
        double ops_per_iter = 8;
        auto run = [&](int it) {
            switch (kind) {
                case 0: pipe_bench_kernel<0><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;
                case 1: pipe_bench_kernel<1><<<blocks, threads>>>(it, sink); ops_per_iter = 16; break;
                case 2: pipe_bench_kernel<2><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;
                case 3: pipe_bench_kernel<3><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;
                case 4: pipe_bench_kernel<4><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;
                case 5: pipe_bench_kernel<5><<<blocks, threads>>>(it, sink); ops_per_iter = 16; break;  // packed instructions
                case 6: pipe_bench_kernel<6><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;   // IMAD.WIDE (+ IADD)
                case 7: pipe_bench_kernel<7><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;   // LOP3
                case 8: pipe_bench_kernel<8><<<blocks, threads>>>(it, sink); ops_per_iter = 16; break;  // packed (+ 16 LOP3)
                case 10: pipe_bench_kernel<10><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;  // DFMA (+ LOP3)
                case 11: pipe_bench_kernel<11><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;  // DFMA (+ packed FP32)
                case 12: pipe_bench_kernel<12><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;  // DFMA (+ IMAD.WIDE, LOP3)
                case 9: pipe_bench_kernel<9><<<blocks, threads>>>(it, sink); ops_per_iter = 4; break;   // normal draws
                case 13: pipe_bench_kernel<13><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;  // mul.hi.u32
                case 14: pipe_bench_kernel<14><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;  // mul.lo.u32
                case 15: pipe_bench_kernel<15><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;  // mul.hi + mul.lo pairs (counted per pair)
                case 16: pipe_bench_kernel<16><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;  // FFMA, register addend
                case 17: pipe_bench_kernel<17><<<blocks, threads>>>(it, sink); ops_per_iter = 8; break;  // FFMA, immediate addend
                case 18: pipe_bench_kernel<18><<<blocks, threads>>>(it, sink); ops_per_iter = 4; break;  // integer widen + DADD
                case 19: pipe_bench_kernel<19><<<blocks, threads, 2048 * sizeof(double2)>>>(it, sink); ops_per_iter = 8; break;  // LDS.128
                default: invalid("unknown pipe_bench kind");
            }
        };
        run(256);
        CUDA_CHECK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            CUDA_CHECK(cudaEventRecord(e0));
            run(iters);
            CUDA_CHECK(cudaEventRecord(e1));
            CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
        }
        CUDA_CHECK(cudaGetLastError());
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *thread_ops_per_s = ops_per_iter * (double)iters * (double)blocks * threads / (best * 1e-3);
    });
}

}  // extern "C"
