// pipe_bench.cuh — instruction-class throughput micro-benchmarks (roofline denominators).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_lean.cuh"

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
    } else if (KIND == 5 || KIND == 8) {  // packed no-FMA Horner step: FMUL2 + FFMA2(acc, 1.0, c); KIND 8 adds one LOP3 per packed op
        unsigned long long a0, a1, a2, a3, a4, a5, a6, a7;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"((float)t), "f"(t + 0.5f));
        a1 = a0 + 1; a2 = a0 + 2; a3 = a0 + 3; a4 = a0 + 4; a5 = a0 + 5; a6 = a0 + 6; a7 = a0 + 7;
        unsigned long long m, one, nz, c;
        asm("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(0.999f));
        asm("mov.b64 %0, {%1, %1};" : "=l"(one) : "f"(1.0f));
        asm("mov.b64 %0, {%1, %1};" : "=l"(nz) : "f"(-0.0f));
        asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(1.5f));
        uint32_t x0 = t, x1 = t + 1, x2 = t + 2, x3 = t + 3, x4 = t + 4, x5 = t + 5, x6 = t + 6, x7 = t + 7;
#define MCP_PB(a, x)                                                                      \
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(m), "l"(nz));        \
        if (KIND == 8) x = (x ^ (uint32_t)i) & (x2 | 0x55u);                              \
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(one), "l"(c));       \
        if (KIND == 8) x = (x ^ x1) | (uint32_t)i;
        for (int i = 0; i < iters; i++) {
            MCP_PB(a0, x0) MCP_PB(a1, x3) MCP_PB(a2, x4) MCP_PB(a3, x5) MCP_PB(a4, x6) MCP_PB(a5, x7) MCP_PB(a6, x0) MCP_PB(a7, x3)
        }
#undef MCP_PB
        if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x12345ull || (x0 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7) == 0x12345u) sink[0] = 1.0;
    } else if (KIND == 6) {  // IMAD.WIDE.U32 only
        uint32_t c0 = t, c1 = t + 1, c2 = t + 2, c3 = t + 3, c4 = t + 4, c5 = t + 5, c6 = t + 6, c7 = t + 7;
        for (int i = 0; i < iters; i++) {
#define MCP_PB(c) { const uint64_t p = (uint64_t)0xD2511F53u * c; c = (uint32_t)(p >> 32) + (uint32_t)p; }
            // the add folds into the next IMAD.WIDE's addend only partly: count IADD as a second op below
            MCP_PB(c0) MCP_PB(c1) MCP_PB(c2) MCP_PB(c3) MCP_PB(c4) MCP_PB(c5) MCP_PB(c6) MCP_PB(c7)
#undef MCP_PB
        }
        if ((c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7) == 0x12345u) sink[0] = c0;
    } else if (KIND == 7) {  // LOP3 only (3-input xor with a loop-varying operand)
        uint32_t c0 = t, c1 = t + 1, c2 = t + 2, c3 = t + 3, c4 = t + 4, c5 = t + 5, c6 = t + 6, c7 = t + 7;
        for (int i = 0; i < iters; i++) {
            const uint32_t u = (uint32_t)i;
            c0 = (c0 & c1) ^ u; c1 = (c1 | c2) ^ u; c2 = (c2 & c3) ^ u; c3 = (c3 | c4) ^ u;
            c4 = (c4 & c5) ^ u; c5 = (c5 | c6) ^ u; c6 = (c6 & c7) ^ u; c7 = (c7 | c0) ^ u;
        }
        if ((c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7) == 0x12345u) sink[0] = c0;
    } else if (KIND == 9) {  // the generation phase alone: Philox4x32-10 (hoisted round keys) -> 4 packed inverse-normal draws
        uint32_t rk[20];
#pragma unroll
        for (int r = 0; r < 10; r++) {
            rk[2 * r] = (uint32_t)t + (uint32_t)r * 0x9E3779B9u;
            rk[2 * r + 1] = (uint32_t)(t >> 3) + (uint32_t)r * 0xBB67AE85u;
        }
        float s0 = 0.f;
        unsigned tails = 0u;
#pragma unroll 3
        for (int i = 0; i < iters; i++) {
            const U4 w = philox4x32_10_rk((uint32_t)i, 3u, 1u, 0u, rk);
            float4 z;
            float v0, v1, v2, v3;
            ni_central_pair_packed(w.x, w.y, z.x, z.y, v0, v1);
            ni_central_pair_packed(w.z, w.w, z.z, z.w, v2, v3);
            unsigned tb = 0u;
            if (fabsf(v0) > MCP_NI_BP) { z.x = v0; tb |= 1u; }
            if (fabsf(v1) > MCP_NI_BP) { z.y = v1; tb |= 2u; }
            if (fabsf(v2) > MCP_NI_BP) { z.z = v2; tb |= 4u; }
            if (fabsf(v3) > MCP_NI_BP) { z.w = v3; tb |= 8u; }
            tails += tb;
            s0 = __uint_as_float(__float_as_uint(s0) ^ __float_as_uint(z.x) ^ __float_as_uint(z.y) ^ __float_as_uint(z.z) ^ __float_as_uint(z.w));
        }
        if (s0 == 0.12345f && tails == 77u) sink[0] = s0;
    } else if (KIND == 10 || KIND == 11 || KIND == 12) {  // co-issue: 8 DFMA + 8 of {LOP3 | packed FP32 | IMAD.WIDE} per iteration
        double a0 = t, a1 = t + 1, a2 = t + 2, a3 = t + 3, a4 = t + 4, a5 = t + 5, a6 = t + 6, a7 = t + 7;
        const double m = 1.0000001, c = 1e-9;
        uint32_t c0 = t, c1 = t + 1, c2 = t + 2, c3 = t + 3, c4 = t + 4, c5 = t + 5, c6 = t + 6, c7 = t + 7;
        unsigned long long p0 = t, p1 = t + 1, p2 = t + 2, p3 = t + 3, p4 = t + 4, p5 = t + 5, p6 = t + 6, p7 = t + 7, mm, nz;
        asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(0.999f));
        asm("mov.b64 %0, {%1, %1};" : "=l"(nz) : "f"(-0.0f));
        for (int i = 0; i < iters; i++) {
            const uint32_t u = (uint32_t)i;
#define MCP_PB(a, x, y, p)                                                                              \
            a = fma(a, m, c);                                                                           \
            if (KIND == 10) x = (x & y) ^ u;                                                            \
            if (KIND == 11) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(mm), "l"(nz));  \
            if (KIND == 12) { const uint64_t w = (uint64_t)0xD2511F53u * x; x = (uint32_t)(w >> 32) ^ (uint32_t)w; }
            MCP_PB(a0, c0, c1, p0) MCP_PB(a1, c1, c2, p1) MCP_PB(a2, c2, c3, p2) MCP_PB(a3, c3, c4, p3)
            MCP_PB(a4, c4, c5, p4) MCP_PB(a5, c5, c6, p5) MCP_PB(a6, c6, c7, p6) MCP_PB(a7, c7, c0, p7)
#undef MCP_PB
        }
        if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 0.12345 || (c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7) == 0x12345u ||
            (p0 ^ p1 ^ p2 ^ p3 ^ p4 ^ p5 ^ p6 ^ p7) == 0x12345ull)
            sink[0] = a0;
    } else if (KIND == 13 || KIND == 14 || KIND == 15) {  // 13: mul.hi.u32 only, 14: mul.lo.u32 only, 15: the pair (a Philox half-round without IMAD.WIDE)
        uint32_t c0 = t, c1 = t + 1, c2 = t + 2, c3 = t + 3, c4 = t + 4, c5 = t + 5, c6 = t + 6, c7 = t + 7;
        for (int i = 0; i < iters; i++) {
#define MCP_PB(c)                                                                                  \
    {                                                                                              \
        uint32_t h = 0u, l = 0u;                                                                   \
        if (KIND != 14) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(c), "r"(0xD2511F53u)); \
        if (KIND != 13) asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(l) : "r"(c), "r"(0xD2511F53u)); \
        c = h + l + (uint32_t)i;                                                                   \
    }
            MCP_PB(c0) MCP_PB(c1) MCP_PB(c2) MCP_PB(c3) MCP_PB(c4) MCP_PB(c5) MCP_PB(c6) MCP_PB(c7)
#undef MCP_PB
        }
        if ((c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7) == 0x12345u) sink[0] = c0;
    } else if (KIND == 16 || KIND == 17) {  // scalar FFMA: 16 = three register operands, 17 = immediate addend (guide: imm form issues every cycle)
        float a0 = t, a1 = t + 1, a2 = t + 2, a3 = t + 3, a4 = t + 4, a5 = t + 5, a6 = t + 6, a7 = t + 7;
        const float m = 0.999f;
        float cc = KIND == 16 ? (float)(iters & 7) * 0.25f + 1.5f : 1.5f;
        for (int i = 0; i < iters; i++) {
            a0 = fmaf(a0, m, cc); a1 = fmaf(a1, m, cc); a2 = fmaf(a2, m, cc); a3 = fmaf(a3, m, cc);
            a4 = fmaf(a4, m, cc); a5 = fmaf(a5, m, cc); a6 = fmaf(a6, m, cc); a7 = fmaf(a7, m, cc);
        }
        if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 0.12345f) sink[0] = a0;
    } else if (KIND == 18) {  // f32 -> f64 widening with integer ops (ALU pipe) instead of F2F (XU), + DADD as in kind 3
        float f0 = t, f1 = t + 1, f2 = t + 2, f3 = t + 3;
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int i = 0; i < iters; i++) {
#define MCP_W(f) __hiloint2double((int)(((__float_as_uint(f) >> 3) & 0x0fffffffu) + 0x38000000u) | (int)(__float_as_uint(f) & 0x80000000u), (int)(__float_as_uint(f) << 29))
            s0 += MCP_W(f0); s1 += MCP_W(f1); s2 += MCP_W(f2); s3 += MCP_W(f3);
#undef MCP_W
            f0 = __int_as_float(__float_as_int(f0) + 1); f1 = __int_as_float(__float_as_int(f1) + 1);
            f2 = __int_as_float(__float_as_int(f2) + 1); f3 = __int_as_float(__float_as_int(f3) + 1);
        }
        if (s0 + s1 + s2 + s3 == 0.12345) sink[0] = s0;
    } else if (KIND == 19) {  // shared-memory pipe: LDS.128 per thread (conflict-free), 8 per iteration
        extern __shared__ double2 pbs[];
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) pbs[i] = make_double2((double)i, 1.0);
        __syncthreads();
        double s0 = 0.0, s1 = 0.0;
        int idx = threadIdx.x;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const double2 v = pbs[(idx + 256 * u) & 2047];
                s0 += v.x;
                s1 += v.y;
            }
            idx = (idx + 1) & 2047;
        }
        if (s0 + s1 == 0.12345) sink[0] = s0;
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
