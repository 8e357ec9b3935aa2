// launch.h — interface between the host API (api.cu) and the kernel translation units (kernel_tu.cu).
// kernel_tu.cu is compiled several times with different -DMCP_TU_ID / -DMCP_P_LO / -DMCP_P_HI / -DMCP_TU_GLM so
// the template instantiations build in parallel (see mcpower_b200/build.py for the TU table).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.h"

namespace mcp {

enum { KIND_OLS = 0, KIND_GLM = 1, KIND_LMM = 2 };

struct KernelArgs {
    uint64_t base_seed, sim_begin;
    uint32_t n_sims;
    cudaStream_t stream;
    float* scratch;          // GLM: per-wave f32 data plane [row][col][sim]
    double* scratch_d;       // LMM: per-wave cluster sums [cluster][col][sim]; lean OLS: statistics [grid point][entry][sim]
    int ols_variant;         // 1 = thread per study (kernels.cuh), 4 = lane quad per study (kernels_q4.cuh), 5 = lean quad (kernels_lean.cuh)
};

// Each returns false when P is outside the TU's range; otherwise *err holds the launch status.
#define MCP_DECLARE_TU(id)                                                                                                 \
    bool tu##id##_launch_power(int P, int kind, bool scen, const DevicePlan& plan, const KernelArgs& a, cudaError_t* err); \
    bool tu##id##_dump(int P, bool scen, const DevicePlan& plan, uint64_t base_seed, uint64_t sim_id, float* dx, float* dy, \
                       cudaStream_t st, cudaError_t* err);                                                                 \
    bool tu##id##_fit_data(int P, const double* dx, const double* dy, int n, double* dout, cudaStream_t st, cudaError_t* err); \
    bool tu##id##_fit_spec(int P, int kind, const DevicePlan& plan, const float* plane, int n, double* scratch, double* dout,  \
                           cudaStream_t st, cudaError_t* err);

#define MCP_TU_LIST(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18)
MCP_TU_LIST(MCP_DECLARE_TU)

}  // namespace mcp
