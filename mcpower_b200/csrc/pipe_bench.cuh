// pipe_bench.cuh — instruction-class throughput micro-benchmarks (roofline denominators).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcp {

// ------------------------------------------------------------------------------------------
// Pipe micro-benchmarks: measured denominators for the issue/FP64 roofline (SURVEY.md §8d —
// MEASURED_PEAKS.json has no FP64/FP32/INT peak).  Each thread runs `iters` rounds of 8 independent
// dependent-chains of one instruction class; the host converts elapsed time to thread-ops/s.
template <int KIND>
__global__ void __launch_bounds__(256) pipe_bench_kernel(int iters, double* sink) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (KIND == 0) {  // DFMA
        double a0 = t, a1 = t + 1, a2 = t + 2, a3 = t + 3, a4 = t + 4, a5 = t + 5, a6 = t + 6, a7 = t + 7;
        const double m = 1.0000001, c = 1e-9;
        for (int i = 0; i < iters; i++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
        if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 0.12345) sink[0] = a0;
    } else if (KIND == 1) {  // FMUL + FADD pairs (the no-FMA Horner step)
        float a0 = t, a1 = t + 1, a2 = t + 2, a3 = t + 3, a4 = t + 4, a5 = t + 5, a6 = t + 6, a7 = t + 7;
        const float m = 0.999f;
        for (int i = 0; i < iters; i++) {
            a0 = a0 * m + 1.5f; a1 = a1 * m + 1.5f; a2 = a2 * m + 1.5f; a3 = a3 * m + 1.5f;
            a4 = a4 * m + 1.5f; a5 = a5 * m + 1.5f; a6 = a6 * m + 1.5f; a7 = a7 * m + 1.5f;
        }
        if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 0.12345f) sink[0] = a0;
    } else if (KIND == 2) {  // IMAD.WIDE.U32 + LOP3 (the Philox round)
        uint32_t c0 = t, c1 = t + 1, c2 = t + 2, c3 = t + 3, d0 = t + 4, d1 = t + 5, d2 = t + 6, d3 = t + 7;
        for (int i = 0; i < iters; i++) {
            const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
            const uint64_t q0 = (uint64_t)0xD2511F53u * d0, q1 = (uint64_t)0xCD9E8D57u * d2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ (uint32_t)i, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ (uint32_t)i;
            const uint32_t m0 = (uint32_t)(q1 >> 32) ^ d1 ^ (uint32_t)i, m2 = (uint32_t)(q0 >> 32) ^ d3 ^ (uint32_t)i;
            c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
            d1 = (uint32_t)q1; d3 = (uint32_t)q0; d0 = m0; d2 = m2;
        }
        if ((c0 ^ c1 ^ c2 ^ c3 ^ d0 ^ d1 ^ d2 ^ d3) == 0x12345u) sink[0] = c0;
    } else if (KIND == 3) {  // F2F.F64.F32 widen + DADD
        float f0 = t, f1 = t + 1, f2 = t + 2, f3 = t + 3;
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int i = 0; i < iters; i++) {
            s0 += (double)f0; s1 += (double)f1; s2 += (double)f2; s3 += (double)f3;
            f0 = __int_as_float(__float_as_int(f0) + 1); f1 = __int_as_float(__float_as_int(f1) + 1);
            f2 = __int_as_float(__float_as_int(f2) + 1); f3 = __int_as_float(__float_as_int(f3) + 1);
        }
        if (s0 + s1 + s2 + s3 == 0.12345) sink[0] = s0;
    } else if (KIND == 4) {  // DADD only (baseline for KIND 3)
        double s0 = t, s1 = t + 1, s2 = t + 2, s3 = t + 3, s4 = t + 4, s5 = t + 5, s6 = t + 6, s7 = t + 7;
        const double c = 1e-9;
        for (int i = 0; i < iters; i++) {
            s0 += c; s1 += c; s2 += c; s3 += c; s4 += c; s5 += c; s6 += c; s7 += c;
        }
        if (s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7 == 0.12345) sink[0] = s0;
    }
}

}  // namespace mcp
