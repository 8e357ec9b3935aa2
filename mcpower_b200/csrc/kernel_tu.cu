// kernel_tu.cu — one kernel translation unit: instantiates the power kernels for P in [MCP_P_LO, MCP_P_HI].
// Compiled several times in parallel (mcpower_b200/build.py).  MCP_TU_GLM selects the GLM kernels.
#include "kernels.cuh"
#if !defined(MCP_TU_GLM) && !defined(MCP_TU_LMM)
#include "kernels_q4.cuh"
#include "kernels_lean.cuh"
#endif
#ifdef MCP_TU_GLM
#include "kernels_glm.cuh"
#endif
#ifdef MCP_TU_LMM
#include "kernels_lmm.cuh"
#endif
#include "launch.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#ifndef MCP_TU_ID
#error "MCP_TU_ID must be defined"
#endif
#define MCP_CAT2(a, b, c) a##b##c
#define MCP_CAT(a, b, c) MCP_CAT2(a, b, c)
#define MCP_TU_FN(name) MCP_CAT(tu, MCP_TU_ID, _##name)

namespace mcp {

static size_t power_smem(const DevicePlan& pl, int P) {
    const int SP = P | 1;
    size_t smem = (size_t)MCP_BLOCK * 4 * SP * sizeof(float);
    if (pl.smem_counters) smem += (size_t)pl.n_ss * pl.counter_width * sizeof(unsigned);
    return smem;
}

#if defined(MCP_TU_GLM)
#define MCP_TU_KIND KIND_GLM
#elif defined(MCP_TU_LMM)
#define MCP_TU_KIND KIND_LMM
#else
#define MCP_TU_KIND KIND_OLS
#define MCP_TU_IS_OLS 1
#endif
// NOTE: KIND_* are enum constants, invisible to the preprocessor — the former `#if MCP_TU_KIND != KIND_OLS` was always false, so the
// GLM / LMM translation units also instantiated the dump / fit kernels.  The duplicate (weak) host stubs were merged by
// the linker and registered by several modules: a launch could run another module's copy of the kernel, reading THAT
// module's __constant__ plan (stale).  Guard with a real macro.

// dynamic shared memory opt-in: 96 KiB covers every exact width; the padded wide widths ask for what they need
template <typename K>
static void allow_smem(K kernel, size_t bytes) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes > 96 * 1024 ? bytes : 96 * 1024));
}

template <int PT>
static cudaError_t launch_one(int kind, bool scen, const DevicePlan& pl, const KernelArgs& a) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_plan, &pl, sizeof(DevicePlan), 0, cudaMemcpyHostToDevice, a.stream);
    if (e != cudaSuccess) return e;
    const size_t smem = power_smem(pl, PT);
    const uint32_t blocks = (a.n_sims + MCP_BLOCK - 1) / MCP_BLOCK;
    if (kind != MCP_TU_KIND) return cudaErrorInvalidValue;
#if defined(MCP_TU_GLM)
    {
        {
            // generator on the lane-quad kernel (PLANE mode): four rows of a study in flight per quad; on the scenario path the
            // per-study designs come from scen_design_kernel and sit behind the plane (api.cu sizes the scratch for both)
            using G = Q4Geom<PT>;
            const size_t smem_q = G::smem_bytes(0, 0, false);
            const uint32_t blocks_q = (a.n_sims + G::STUDIES - 1) / G::STUDIES;
            const size_t stride = (size_t)blocks * MCP_BLOCK;
            if (scen) {
                ScenDesign<PT>* designs = reinterpret_cast<ScenDesign<PT>*>(a.scratch + stride * (size_t)pl.max_n * PT);
                scen_design_kernel<PT, MCP_TU_KIND><<<blocks, MCP_BLOCK, 0, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims, designs);
                e = cudaGetLastError();
                if (e != cudaSuccess) return e;
                allow_smem(ols_power_kernel_q4<PT, true, true, false, true>, smem_q);
                ols_power_kernel_q4<PT, true, true, false, true><<<blocks_q, MCP_BLOCK, smem_q, a.stream>>>(
                    a.base_seed, a.sim_begin, a.n_sims, nullptr, designs, stride, nullptr, 0, a.scratch);
            } else {
                allow_smem(ols_power_kernel_q4<PT, false, true, false, true>, smem_q);
                ols_power_kernel_q4<PT, false, true, false, true><<<blocks_q, MCP_BLOCK, smem_q, a.stream>>>(
                    a.base_seed, a.sim_begin, a.n_sims, nullptr, nullptr, stride, nullptr, 0, a.scratch);
            }
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const size_t smem_fit = pl.smem_counters ? (size_t)pl.n_ss * pl.counter_width * sizeof(unsigned) : 0;
        const dim3 grid_fit(blocks, (unsigned)pl.n_ss);
        // resident CTAs per SM the fit kernel is compiled for: the IRLS accumulators (P(P+1)/2 + P doubles) and the two rows
        // in flight set the register floor; measured at P = 5 (config 4, two-row loop): 5 → 1.86e7, 4 → 2.30e7, 3 → 2.17e7 fits/s
#ifdef MCP_GLM_MINB
        constexpr int MINB = MCP_GLM_MINB;  // (tuning builds, tools/build_variant.sh)
#else
        constexpr int MINB = PT <= 6 ? 4 : (PT <= 8 ? 3 : 2);
#endif
        glm_fit_kernel<PT, MINB><<<grid_fit, MCP_BLOCK, smem_fit, a.stream>>>(a.n_sims, a.scratch);
    }
#elif defined(MCP_TU_LMM)
    if constexpr (PT <= MCP_MAX_TEMPLATE_P) if (pl.n_extra > 0) {
        // extra groupings: the f32 data plane, then one dense penalised-least-squares fit per (study, sample size)
        const size_t stride = (size_t)blocks * MCP_BLOCK;
        float* const plane = reinterpret_cast<float*>(a.scratch_d);
        double* const fit_scratch = a.scratch_d + ((size_t)pl.max_n * PT * stride + 1) / 2;
        const size_t smem_gen = (size_t)MCP_BLOCK * 4 * (PT | 1) * sizeof(float);
        allow_smem(lmm_plane_kernel<PT>, smem_gen);
        lmm_plane_kernel<PT><<<blocks, MCP_BLOCK, smem_gen, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims, plane);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const size_t smem_fit = pl.smem_counters ? (size_t)pl.n_ss * pl.counter_width * sizeof(unsigned) : 0;
        lmm_general_kernel<PT><<<dim3(blocks, (unsigned)pl.n_ss), MCP_BLOCK, smem_fit, a.stream>>>(a.n_sims, plane, fit_scratch, stride);
        return cudaGetLastError();
    }
    {
        using LY = LmmLayout<PT>;
        using G = Q4Geom<PT>;
        const size_t stride = (size_t)blocks * MCP_BLOCK;
        const int cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
        const int nc_cap = pl.sizing_fixed_size ? (pl.max_n + cs - 1) / cs : cs;
        const int clus_doubles = (int)LY::csum_doubles(nc_cap);
        const size_t smem_q = G::smem_bytes(0, 0, false) + (size_t)G::STUDIES * (size_t)((clus_doubles | 1) + (nc_cap | 1)) * sizeof(double);
        if (smem_q <= 56 * 1024 && pl.n_slopes == 0) {  // (the random-slope sums live in HBM: one-thread-per-study generator)
            // generator on the lane-quad kernel: cluster sums in shared memory (4 CTAs per SM fit), no read-modify-write of HBM per
            // row; on the scenario path the per-study designs come from scen_design_kernel and sit behind the layout
            double* snap = a.scratch_d + LY::csum_doubles(nc_cap) * stride;
            double* stat = snap + LY::csum_doubles(nc_cap) * (size_t)pl.n_ss * stride;
            const uint32_t blocks_q = (a.n_sims + G::STUDIES - 1) / G::STUDIES;
            if (scen) {
                ScenDesign<PT>* designs = reinterpret_cast<ScenDesign<PT>*>(a.scratch_d + LY::per_study_doubles(nc_cap, pl.n_ss) * stride);
                scen_design_kernel<PT, MCP_TU_KIND><<<blocks, MCP_BLOCK, 0, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims, designs);
                e = cudaGetLastError();
                if (e != cudaSuccess) return e;
                cudaFuncSetAttribute(ols_power_kernel_q4<PT, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
                ols_power_kernel_q4<PT, true, true, true><<<blocks_q, MCP_BLOCK, smem_q, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims, stat, designs,
                                                                                                    stride, snap, clus_doubles);
            } else {
                cudaFuncSetAttribute(ols_power_kernel_q4<PT, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
                ols_power_kernel_q4<PT, false, true, true><<<blocks_q, MCP_BLOCK, smem_q, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims, stat, nullptr,
                                                                                                     stride, snap, clus_doubles);
            }
        } else {
            const size_t smem_gen = (size_t)MCP_BLOCK * 4 * (PT | 1) * sizeof(float);
            allow_smem(lmm_gen_kernel<PT>, smem_gen);
            lmm_gen_kernel<PT><<<blocks, MCP_BLOCK, smem_gen, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims, a.scratch_d);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const size_t smem_fit = pl.smem_counters ? (size_t)pl.n_ss * pl.counter_width * sizeof(unsigned) : 0;
        lmm_fit_kernel<PT><<<dim3(blocks, (unsigned)pl.n_ss), MCP_BLOCK, smem_fit, a.stream>>>(a.n_sims, a.scratch_d, stride);
    }
#else
    const int variant = PT > MCP_MAX_TEMPLATE_P ? 6 : a.ols_variant;  // padded wide widths: the split path only
    if constexpr (PT <= MCP_MAX_TEMPLATE_P) if (variant == 5) {  // plain-linear-model family: lane-quad generate + accumulate kernel, then one thread per fit
        using G = LeanGeom<PT>;
        const size_t smem5 = G::smem_bytes();
        const uint32_t blocks5 = (a.n_sims + G::STUDIES - 1) / G::STUDIES;
#define MCP_LAUNCH_LEAN(...)                                                                                              \
    do {                                                                                                                  \
        cudaFuncSetAttribute(ols_lean_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5);      \
        ols_lean_kernel<__VA_ARGS__><<<blocks5, MCP_BLOCK, smem5, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims,        \
                                                                              a.scratch_d);                               \
    } while (0)
#if MCP_LEAN_FACT_ONLY
        if (pl.n_inter > 0) MCP_LAUNCH_LEAN(PT, true, true);
        else if (pl.n_factors > 0) MCP_LAUNCH_LEAN(PT, true, false, MCP_LEAN_UNROLL, true);  // dummies, no interaction: x is not written back to the slab
#else
        if (pl.n_inter > 0 || pl.n_factors > 0) MCP_LAUNCH_LEAN(PT, true, true);
#endif
        else if (!pl.l_identity) MCP_LAUNCH_LEAN(PT, true, false);
        else {
#ifdef MCP_LEAN_TUNE  // tuning build only: MCPOWER_B200_LEAN_TUNE=u<unroll> picks an instantiation at P = 9
            const char* tune = getenv("MCPOWER_B200_LEAN_TUNE");
            if (PT == 9 && tune && tune[0] == 'u' && tune[1] == '1') MCP_LAUNCH_LEAN(PT, false, false, 1);
            else if (PT == 9 && tune && tune[0] == 'u' && tune[1] == '2') MCP_LAUNCH_LEAN(PT, false, false, 2);
            else if (PT == 9 && tune && tune[0] == 'u' && tune[1] == '4') MCP_LAUNCH_LEAN(PT, false, false, 4);
            else
#endif
                MCP_LAUNCH_LEAN(PT, false, false);
        }
#undef MCP_LAUNCH_LEAN
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const size_t smem_s = pl.smem_counters ? (size_t)pl.n_ss * pl.counter_width * sizeof(unsigned) : 0;
        const dim3 grid_s((a.n_sims + MCP_BLOCK - 1) / MCP_BLOCK, (unsigned)pl.n_ss);
        ols_solve_kernel<PT><<<grid_s, MCP_BLOCK, smem_s, a.stream>>>(a.n_sims, a.scratch_d);
    }
    if (variant == 6) {  // general quad generate + accumulate kernel, then one thread per fit (kernels_q4.cuh, SPLIT)
        using G = Q4Geom<PT>;
        const size_t smem6 = G::smem_bytes(0, 0, false);
        const uint32_t blocks6 = (a.n_sims + G::STUDIES - 1) / G::STUDIES;
        if (scen) {
            // per-study scenario designs behind this wave's statistics (api.cu sizes the scratch for both)
            ScenDesign<PT>* designs = reinterpret_cast<ScenDesign<PT>*>(a.scratch_d + (size_t)pl.n_ss * (PT * (PT + 1) / 2 + PT + 1) * a.n_sims);
            scen_design_kernel<PT, MCP_TU_KIND><<<blocks, MCP_BLOCK, 0, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims, designs);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            allow_smem(ols_power_kernel_q4<PT, true, true>, smem6);
            ols_power_kernel_q4<PT, true, true><<<blocks6, MCP_BLOCK, smem6, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims, a.scratch_d, designs);
        } else {
            allow_smem(ols_power_kernel_q4<PT, false, true>, smem6);
            ols_power_kernel_q4<PT, false, true><<<blocks6, MCP_BLOCK, smem6, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims, a.scratch_d);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const size_t smem_s = pl.smem_counters ? (size_t)pl.n_ss * pl.counter_width * sizeof(unsigned) : 0;
        const dim3 grid_s((a.n_sims + MCP_BLOCK - 1) / MCP_BLOCK, (unsigned)pl.n_ss);
        ols_solve_kernel<PT><<<grid_s, MCP_BLOCK, smem_s, a.stream>>>(a.n_sims, a.scratch_d);
    }
    if constexpr (PT <= MCP_MAX_TEMPLATE_P) if (variant == 4) {  // one study per lane quad (kernels_q4.cuh)
        using G = Q4Geom<PT>;
        const size_t smem4 = G::smem_bytes(pl.n_ss, pl.counter_width, pl.smem_counters != 0);
        const uint32_t blocks4 = (a.n_sims + G::STUDIES - 1) / G::STUDIES;
        if (scen) {
            cudaFuncSetAttribute(ols_power_kernel_q4<PT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            ols_power_kernel_q4<PT, true><<<blocks4, MCP_BLOCK, smem4, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims);
        } else {
            cudaFuncSetAttribute(ols_power_kernel_q4<PT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            ols_power_kernel_q4<PT, false><<<blocks4, MCP_BLOCK, smem4, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims);
        }
    }
    if constexpr (PT <= MCP_MAX_TEMPLATE_P) if (variant != 4 && variant != 5 && variant != 6) {
        if (scen) {
            cudaFuncSetAttribute(ols_power_kernel<PT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            ols_power_kernel<PT, true><<<blocks, MCP_BLOCK, smem, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims);
        } else {
            cudaFuncSetAttribute(ols_power_kernel<PT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            ols_power_kernel<PT, false><<<blocks, MCP_BLOCK, smem, a.stream>>>(a.base_seed, a.sim_begin, a.n_sims);
        }
    }
#endif
    return cudaGetLastError();
}

template <int PT>
static cudaError_t dump_one(bool scen, const DevicePlan& pl, uint64_t base_seed, uint64_t sim_id, float* dx, float* dy, cudaStream_t st) {
#ifndef MCP_TU_IS_OLS
    return cudaErrorInvalidValue;
#else
    cudaError_t e = cudaMemcpyToSymbolAsync(c_plan, &pl, sizeof(DevicePlan), 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    if (scen) dgp_dump_kernel<PT, true><<<1, 32, 0, st>>>(base_seed, sim_id, dx, dy);
    else dgp_dump_kernel<PT, false><<<1, 32, 0, st>>>(base_seed, sim_id, dx, dy);
    return cudaGetLastError();
#endif
}

template <int PT>
static cudaError_t fit_one(const double* dx, const double* dy, int n, double* dout, cudaStream_t st) {
#ifndef MCP_TU_IS_OLS
    return cudaErrorInvalidValue;
#else
    if constexpr (PT <= MCP_MAX_TEMPLATE_P) {
        fit_data_kernel<PT><<<1, 32, 0, st>>>(dx, dy, n, dout);
        return cudaGetLastError();
    } else {
        (void)dx; (void)dy; (void)n; (void)dout; (void)st;
        return cudaErrorInvalidValue;
    }
#endif
}

// fit_provided_data twins of the GLM / mixed-model arms: one supplied dataset on the arm's own fit code.
template <int PT>
static cudaError_t fit_spec_one(const DevicePlan& pl, const float* plane, int n, double* scratch, double* dout, cudaStream_t st) {
#if defined(MCP_TU_GLM)
    cudaError_t e = cudaMemcpyToSymbolAsync(c_plan, &pl, sizeof(DevicePlan), 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    (void)scratch;
    if constexpr (PT <= MCP_MAX_TEMPLATE_P) {
        glm_fit_data_kernel<PT><<<1, 32, 0, st>>>(plane, n, dout);
        return cudaGetLastError();
    } else {
        return cudaErrorInvalidValue;
    }
#elif defined(MCP_TU_LMM)
    cudaError_t e = cudaMemcpyToSymbolAsync(c_plan, &pl, sizeof(DevicePlan), 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    if constexpr (PT <= MCP_MAX_TEMPLATE_P) {
        if (pl.n_extra > 0) {
            lmm_general_fit_data_kernel<PT><<<1, 32, 0, st>>>(plane, n, scratch, dout);
            return cudaGetLastError();
        }
    }
    lmm_fit_data_kernel<PT><<<1, 32, 0, st>>>(plane, n, scratch, dout);
    return cudaGetLastError();
#else
    (void)pl; (void)plane; (void)n; (void)scratch; (void)dout; (void)st;
    return cudaErrorInvalidValue;
#endif
}

#define MCP_CASE(p, expr) \
    case p:               \
        if (p >= MCP_P_LO && p <= MCP_P_HI) { *err = expr; return true; } \
        return false;

// The switch lists every P; only cases inside [MCP_P_LO, MCP_P_HI] instantiate templates (if constexpr-like via
// the preprocessor below).
template <int PT, bool IN = (PT >= MCP_P_LO && PT <= MCP_P_HI)>
struct Sel {
    static cudaError_t launch(int kind, bool scen, const DevicePlan& pl, const KernelArgs& a) { return launch_one<PT>(kind, scen, pl, a); }
    static cudaError_t dump(bool scen, const DevicePlan& pl, uint64_t s, uint64_t i, float* dx, float* dy, cudaStream_t st) { return dump_one<PT>(scen, pl, s, i, dx, dy, st); }
    static cudaError_t fit(const double* dx, const double* dy, int n, double* o, cudaStream_t st) { return fit_one<PT>(dx, dy, n, o, st); }
    static cudaError_t fit_spec(const DevicePlan& pl, const float* plane, int n, double* sc, double* o, cudaStream_t st) { return fit_spec_one<PT>(pl, plane, n, sc, o, st); }
};
template <int PT>
struct Sel<PT, false> {
    static cudaError_t launch(int, bool, const DevicePlan&, const KernelArgs&) { return cudaErrorInvalidValue; }
    static cudaError_t dump(bool, const DevicePlan&, uint64_t, uint64_t, float*, float*, cudaStream_t) { return cudaErrorInvalidValue; }
    static cudaError_t fit(const double*, const double*, int, double*, cudaStream_t) { return cudaErrorInvalidValue; }
    static cudaError_t fit_spec(const DevicePlan&, const float*, int, double*, double*, cudaStream_t) { return cudaErrorInvalidValue; }
};

#define MCP_ALL_P(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(24) X(32)

bool MCP_TU_FN(launch_power)(int P, int kind, bool scen, const DevicePlan& plan, const KernelArgs& a, cudaError_t* err) {
    if (P < MCP_P_LO || P > MCP_P_HI) return false;
    if (kind != MCP_TU_KIND) return false;
    switch (P) {
#define X(p) case p: *err = Sel<p>::launch(kind, scen, plan, a); return true;
        MCP_ALL_P(X)
#undef X
    }
    return false;
}

bool MCP_TU_FN(dump)(int P, bool scen, const DevicePlan& plan, uint64_t base_seed, uint64_t sim_id, float* dx, float* dy,
                     cudaStream_t st, cudaError_t* err) {
#ifndef MCP_TU_IS_OLS
    return false;
#else
    if (P < MCP_P_LO || P > MCP_P_HI) return false;
    switch (P) {
#define X(p) case p: *err = Sel<p>::dump(scen, plan, base_seed, sim_id, dx, dy, st); return true;
        MCP_ALL_P(X)
#undef X
    }
    return false;
#endif
}

bool MCP_TU_FN(fit_data)(int P, const double* dx, const double* dy, int n, double* dout, cudaStream_t st, cudaError_t* err) {
#ifndef MCP_TU_IS_OLS
    return false;
#else
    if (P < MCP_P_LO || P > MCP_P_HI) return false;
    switch (P) {
#define X(p) case p: *err = Sel<p>::fit(dx, dy, n, dout, st); return true;
        MCP_ALL_P(X)
#undef X
    }
    return false;
#endif
}

bool MCP_TU_FN(fit_spec)(int P, int kind, const DevicePlan& plan, const float* plane, int n, double* scratch, double* dout,
                         cudaStream_t st, cudaError_t* err) {
#ifdef MCP_TU_IS_OLS
    (void)P; (void)kind; (void)plan; (void)plane; (void)n; (void)scratch; (void)dout; (void)st; (void)err;
    return false;  // the OLS arm is served by fit_data
#else
    if (P < MCP_P_LO || P > MCP_P_HI || kind != MCP_TU_KIND) return false;
    switch (P) {
#define X(p) case p: *err = Sel<p>::fit_spec(plan, plane, n, scratch, dout, st); return true;
        MCP_ALL_P(X)
#undef X
    }
    return false;
#endif
}

}  // namespace mcp
