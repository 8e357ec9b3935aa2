// kernels_lean.cuh — the headline kernel: OLS power for the "plain linear model" spec family, one simulated study
// per lane quad, written for a small register footprint (4–5 warps per scheduler) and a short instruction stream.
//
// Spec family (host-checked, api.cu lean_ok): OLS, continuous outcome, optimistic scenario, Normal marginals, Normal
// residual, no cluster, factors only with the exact (sim-invariant) level walk; correlated predictors (MIX), factor
// dummies (FACT) and interaction columns (INTER) are compile-time options.  BASELINE.json's metric shape (n = 500,
// 8 N(0,1) predictors, P = 9), config 1 (y ~ x1 + x2 + x1:x2) and config 2 (sample-size sweep, 8 correlated
// predictors + 3-level factor, Bonferroni / BH) are in the family; everything else runs on the general kernels
// (kernels.cuh / kernels_q4.cuh).  Arithmetic and operation order are those of the general kernels and of the oracle — the data
// plane follows rng.rs:154-364 / data_gen.rs:334-363,643-649,781-888, the statistics are one sequential fma chain per
// accumulator — so every output bit is the same whichever kernel runs (tests/test_gpu_parity.py runs all of them on
// the same inputs).
//
// Work split inside a quad (lane q = threadIdx & 3), per "quad step" of 16 rows:
//   A. lane q generates Philox block 4·sb + q (rows 16·sb + 4q … +3) for every column: one Philox4x32-10 call →
//      4 words → 4 inverse-normal draws (packed f32x2 Horner: FMUL2 for the product, FFMA2(acc, 1.0, c) for the
//      separately rounded add — bit-identical to mul + add, half the issue slots) → one STS.128 into the warp's f32
//      slab.  Tail draws (4.85 %) are flagged in a bit mask and patched afterwards (B).
//   C. four sub-steps qq = 0..3: lane q finishes row 16·sb + 4·qq + q from lane qq's slab entries — widen to f64
//      ONCE, Cholesky mix (MIX), interaction products (INTER), linear predictor, y — and stages the row in shared
//      memory as Z = [1 | x | y] in f64, low half forwards + high half reversed (kernels_q4.cuh explains the
//      reflected-quarter trick);
//   D. all four lanes run the same 4-row accumulation pattern on their quarter of Z'Z (M(M+1)/2 accumulators per
//      lane: 15 at P = 9 instead of 56 per thread).
// At a grid point the quad hands its entries to lane 0 by shuffle and lane 0 solves (ols_fit_and_emit).
//
// Slab layout (per warp, floats): [column c][producer lane qq][slot i ^ 2·qq][row r], i = quad index in the warp.
// The XOR swizzle makes both the producer's STS.128 (8 lanes of a quarter-warp → 8 distinct 16-byte slots) and the
// consumer's LDS.32 (32 lanes → 32 distinct banks) conflict-free.
#pragma once
#include "kernels.cuh"

namespace mcp {

#ifndef MCP_LEAN_MIN_BLOCKS
#define MCP_LEAN_MIN_BLOCKS 4
#endif
#ifndef MCP_LEAN_UNROLL
#define MCP_LEAN_UNROLL 3
#endif
// Instruction-count switches of the headline kernel (profiles/r02_lean_kernel.md has the A/B numbers; every setting computes
// the same bits):
//   MCP_LEAN_DYNLP   linear predictor of the plain kernels.  0: DMUL + DADD + two selects per column (ptxas if-converts the
//                    per-column zero-beta test).  1: a uniform loop over the non-zero effects that re-reads those x from the
//                    slab (LDC -> LDS -> F2F -> DMUL -> DADD: few instructions, one long dependent chain per row).  2 (default):
//                    every column, no test — z·(±0) leaves the +0-seeded sum unchanged bit for bit; 17 FP64 instructions on
//                    values that already sit in registers.  Measured on top of the interior path: 0 -> 8.97e7, 1 -> 8.90e7,
//                    2 -> 9.26e7 fits/s.  A two-slot variant of 1 (at most two non-zero slopes, 10 instructions per row) ran
//                    SLOWER than 2 (9.43e7 against 9.56e7): the chain's latency costs more than the FP64 issue slots
//   MCP_LEAN_PUSH32  tail-queue push loop on two 32-bit masks instead of one 64-bit mask
//   MCP_LEAN_SCAN    tail-queue slots from a warp prefix sum (shuffles) instead of one shared-memory atomic per lane
#ifndef MCP_LEAN_DYNLP
#define MCP_LEAN_DYNLP 2
#endif
#ifndef MCP_LEAN_PUSH32
#define MCP_LEAN_PUSH32 1
#endif
#ifndef MCP_LEAN_SCAN
#define MCP_LEAN_SCAN 0
#endif
//   MCP_LEAN_SCALAR_PAIRS  how many of a column task's two draw pairs run the central polynomial in scalar FMUL / FADD (the
//                    other(s) packed f32x2): packed halves the issue slots but runs on the FMA-heavy pipe next to Philox's
//                    IMAD.WIDE; scalar can use the FMA-lite pipe
#ifndef MCP_LEAN_SCALAR_PAIRS
#define MCP_LEAN_SCALAR_PAIRS 0
#endif
//   MCP_LEAN_INTERIOR  a quad step that lies wholly inside the design (16 more rows exist) and ahead of the next grid point
//                    runs its four finishing / accumulation sub-steps fully unrolled and without the per-row bounds and
//                    grid-point tests (the slab addresses of the sub-step fold to constants).  Plain kernels only (no MIX /
//                    INTER / FACT): measured 8.55e7 -> 8.91e7 fits/s on the headline shape at 126 registers, while the
//                    MIX / FACT instantiations spill with it (config-2 sweep 5.96e8 -> 5.29e8) and keep the rolled loop
#ifndef MCP_LEAN_INTERIOR
#define MCP_LEAN_INTERIOR 1
#endif
//   MCP_LEAN_PAD     value stored in the padding slots of a staged row (never read): 0.0 is one CS2R, z[0] two MOVs
#ifndef MCP_LEAN_PAD
#define MCP_LEAN_PAD 0.0
#endif
//   MCP_LEAN_PUSH_UNROLL  unroll factor of the tail-queue push loops (4 = what ptxas picked on its own; 1 measured +1 % with the
//                    rolled generation loop and -0.5 % with the fully unrolled one)
#ifndef MCP_LEAN_PUSH_UNROLL
#define MCP_LEAN_PUSH_UNROLL 4
#endif
//   MCP_LEAN_FULL_GEN  plain kernels of at most this many columns unroll the generation loop completely (k is a compile-time
//                    constant there: the residual / column-half selects of a task fold away): 9.27e7 -> 9.55e7 fits/s at P = 9
#ifndef MCP_LEAN_FULL_GEN
#define MCP_LEAN_FULL_GEN 10
#endif
//   MCP_LEAN_LP_ALL_EVERYWHERE  the MIX / FACT / INTER instantiations multiply every column into the linear predictor as well
//                    (instead of DMUL + DADD + two selects per column): config-2 sweep 6.61e8 -> 6.90e8 fits/s
#ifndef MCP_LEAN_LP_ALL_EVERYWHERE
#define MCP_LEAN_LP_ALL_EVERYWHERE 1
#endif
//   MCP_LEAN_FACT_ONLY  designs with factor dummies but no interaction run <P, MIX, INTER = false, FACT = true>: the finished x of
//                    a row are not written back to the slab (only interaction products read them there): config-2 sweep
//                    6.61e8 -> 7.41e8, with the switch above 7.58e8 fits/s (kernel_tu.cu picks the instantiation)
#ifndef MCP_LEAN_FACT_ONLY
#define MCP_LEAN_FACT_ONLY 1
#endif


// Philox4x32-10 with the 10 round keys precomputed per study (rk[2r] = k0 + r·W0, rk[2r+1] = k1 + r·W1).
__device__ __forceinline__ U4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return U4{c0, c1, c2, c3};
}

template <int PT>
struct LeanGeom {
    static constexpr int N = PT;                  // index of y in Z = [1 | x_1 .. x_{P-1} | y]
    static constexpr int M = (PT + 2) / 2;        // slots per half = ceil((P+1)/2)
    static constexpr int NACCQ = M * (M + 1) / 2; // accumulators per lane
    static constexpr int MH = (M + 1) & ~1;       // half padded to an even count (16-byte vector loads)
    static constexpr int HOFF = mcp_stage_geom(M).hoff;    // offset of the reversed high half inside a staged row
    static constexpr int ROW = mcp_stage_geom(M).row;      // doubles per staged row: [z_0..z_{M-1} pad | z_N, z_{N-1} .. z_{N-M+1} pad]
    static constexpr int STUDY = mcp_stage_geom(M).study;  // 4 staged rows per study; strides chosen conflict-free (plan.h)
    static_assert(HOFF >= MH && ROW >= HOFF + MH && STUDY >= 4 * ROW, "staging geometry");
    static constexpr int STUDIES = MCP_BLOCK / 4; // studies per CTA
    static constexpr int WARPS = MCP_BLOCK / 32;
    static constexpr int SLAB_WARP = PT * 128;    // floats per warp: [PT columns][4 producers][8 slots][4 rows]
    static constexpr int E = PT * (PT + 1) / 2 + PT + 1;  // statistics per fit: X'X (packed lower) | X'y | y'y
    static size_t smem_bytes() {  // slabs | staged rows (aliased by the tail queues) | one tail-queue counter per warp
        return (size_t)WARPS * SLAB_WARP * sizeof(float) + (size_t)STUDIES * STUDY * sizeof(double) + WARPS * sizeof(unsigned);
    }
    static_assert(8 * STUDY * sizeof(double) >= (size_t)SLAB_WARP * sizeof(unsigned short), "tail queue must fit the staged rows");
};

// One staged row into the lane's local pattern: acc[a][c] += p[a]·r[c], c <= a < M; p and r are the low or the
// reversed-high half of the row depending on the lane.
template <int PT>
__device__ __forceinline__ void lean_acc_row(const double* __restrict__ prow, const double* __restrict__ rrow,
                                             double (&acc)[LeanGeom<PT>::NACCQ]) {
    constexpr int M = LeanGeom<PT>::M, MH = LeanGeom<PT>::MH;
    double pv[MH], rv[MH];
    const double2* const pp = reinterpret_cast<const double2*>(prow);
    const double2* const rp = reinterpret_cast<const double2*>(rrow);
#pragma unroll
    for (int j = 0; j < MH / 2; j++) {
        const double2 a2 = pp[j], b2 = rp[j];
        pv[2 * j] = a2.x;
        pv[2 * j + 1] = a2.y;
        rv[2 * j] = b2.x;
        rv[2 * j + 1] = b2.y;
    }
#pragma unroll
    for (int a = 0; a < M; a++) {
#pragma unroll
        for (int c = 0; c <= a; c++) acc[tri(a, c)] = fma(pv[a], rv[c], acc[tri(a, c)]);
    }
}

// One column task of the generation phase: Philox block b of column t (t == k: the residual) -> 4 packed draws ->
// the lane's float4 in the slab; tail draws keep v and are flagged in tmask (bits 4t.., residual at 60..).
template <int PT>
__device__ __forceinline__ void lean_gen_task(int t, int k, uint32_t b, const uint32_t (&rk)[20], float* own, unsigned long long& tmask,
                                              unsigned& tlo, unsigned& thi) {
    const bool is_x = t < k;
    const U4 w = philox4x32_10_rk(b, is_x ? (uint32_t)t : 0u, is_x ? 1u : 2u, 0u, rk);
    float4 z;
    float v0, v1, v2, v3;
#if MCP_LEAN_SCALAR_PAIRS >= 2
    v0 = word_to_v(w.x); v1 = word_to_v(w.y);
    z.x = ni_central(v0); z.y = ni_central(v1);
#else
    ni_central_pair_packed(w.x, w.y, z.x, z.y, v0, v1);
#endif
#if MCP_LEAN_SCALAR_PAIRS >= 1
    v2 = word_to_v(w.z); v3 = word_to_v(w.w);
    z.z = ni_central(v2); z.w = ni_central(v3);
#else
    ni_central_pair_packed(w.z, w.w, z.z, z.w, v2, v3);
#endif
    unsigned tb = 0u;
    if (fabsf(v0) > MCP_NI_BP) { z.x = v0; tb |= 1u; }
    if (fabsf(v1) > MCP_NI_BP) { z.y = v1; tb |= 2u; }
    if (fabsf(v2) > MCP_NI_BP) { z.z = v2; tb |= 4u; }
    if (fabsf(v3) > MCP_NI_BP) { z.w = v3; tb |= 8u; }
#if MCP_LEAN_PUSH32
    // columns 0..7 in the low word, 8..14 in the high word; the residual (slab column PT - 1) takes the top nibble
    if (is_x && t < 8) tlo |= tb << (4 * t);
    else thi |= tb << (is_x ? 4 * (t - 8) : 28);
#else
    tmask |= (unsigned long long)tb << (is_x ? 4 * t : 60);
#endif
    *reinterpret_cast<float4*>(own + 128 * (is_x ? t : PT - 1)) = z;
}

// grid = ⌈n_sims / 32⌉ CTAs of MCP_BLOCK threads (32 studies each), 4 CTAs per SM.  All cooperation is inside a warp
// (8 studies): full-mask warp syncs only, no block-level barrier.
// Measured alternatives that were slower on B200 (tools/tune_lean.sh, gpurun_out/tune_lean*.log): 3 or 5-6 CTAs per SM;
// a persistent one-CTA-per-SM variant with staggered warp starts; f32 staging (4x the F2F.F64.F32 count: the FP64
// pipe is the shared one); 16-row staging with the accumulation fused into the next step's generation code.
template <int PT, bool MIX, bool INTER, int UNR = MCP_LEAN_UNROLL, bool FACT = INTER>
__global__ void __launch_bounds__(MCP_BLOCK, MCP_LEAN_MIN_BLOCKS) ols_lean_kernel(uint64_t base_seed, uint64_t sim_begin, uint32_t n_sims,
                                                                                  double* __restrict__ stats) {
    using G = LeanGeom<PT>;
    constexpr int NACC = PT * (PT + 1) / 2;
    constexpr int N = G::N, M = G::M, MH = G::MH, HOFF = G::HOFF;
    constexpr int PUSH_UNR = MCP_LEAN_PUSH_UNROLL;
    const DevicePlan& pl = c_plan;
    if (mcp_cancelled()) return;
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5;
    const int q = threadIdx.x & 3;
    const int qi = (threadIdx.x & 31) >> 2;  // quad (study) index inside the warp
    float* const wslab = smem + (size_t)warp * G::SLAB_WARP;
    double* const stage_all = reinterpret_cast<double*>(smem + (size_t)G::WARPS * G::SLAB_WARP);
    double* const stage = stage_all + (size_t)(threadIdx.x >> 2) * G::STUDY;  // this study's 4 staged rows
    unsigned short* const wq = reinterpret_cast<unsigned short*>(stage_all + (size_t)warp * 8 * G::STUDY);
    unsigned* const wq_cnt = reinterpret_cast<unsigned*>(stage_all + (size_t)G::STUDIES * G::STUDY) + warp;
    if ((threadIdx.x & 31u) == 0u) *wq_cnt = 0u;
    __syncwarp();
    // Every warp runs with all 32 lanes in lock-step (full-mask warp syncs, warp-cooperative tail queue): a quad past
    // the end of the wave recomputes the last study and simply does not write statistics.
    const uint32_t sim_raw = blockIdx.x * G::STUDIES + (threadIdx.x >> 2);
    const bool valid = sim_raw < n_sims;
    const uint32_t sim = valid ? sim_raw : n_sims - 1;
    const uint64_t key = pcg_mix64(base_seed, sim_begin + sim);
    uint32_t rk[20];
#pragma unroll
    for (int r = 0; r < 10; r++) {
        rk[2 * r] = (uint32_t)key + (uint32_t)r * 0x9E3779B9u;
        rk[2 * r + 1] = (uint32_t)(key >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    const int max_n = pl.max_n;
    const int k = (INTER || FACT) ? pl.k : PT - 1;  // P = 1 + k + factor dummies + interactions
    double acc[G::NACCQ];
#pragma unroll
    for (int i = 0; i < G::NACCQ; i++) acc[i] = 0.0;
    int g = 0;
    uint32_t n_next = pl.ss[0];
    const double* const p_base = stage + ((q == 1 || q == 2) ? HOFF : 0);
    const double* const r_base = stage + (q >= 2 ? HOFF : 0);
    const int own_off = q * 32 + ((qi ^ (2 * q)) << 2);
    float* const own = wslab + own_off;  // + 128·column: this lane's float4 of 4 rows
    const int n_qsteps = (max_n + 15) >> 4;
    for (int sb = 0; sb < n_qsteps; sb++) {
        // ---- A. generate Philox block 4·sb + q (rows 16·sb + 4q .. +3) for every column
        unsigned long long tmask = 0ull;
        unsigned tlo = 0u, thi = 0u;
        if (16 * sb + 4 * q < max_n) {
            const uint32_t b = (uint32_t)(4 * sb + q);
            if constexpr (!MIX && !INTER && !FACT && PT <= MCP_LEAN_FULL_GEN) {
#pragma unroll
                for (int t = 0; t < PT; t++) lean_gen_task<PT>(t, PT - 1, b, rk, own, tmask, tlo, thi);
            } else {
#pragma unroll UNR
                for (int t = 0; t <= k; t++) lean_gen_task<PT>(t, k, b, rk, own, tmask, tlo, thi);
            }
        }
        // ---- B. deferred tail branch of the inverse-normal kernel (rng.rs:354-364), warp-cooperative: a lane sees
        // Binomial(4(k+1), 0.0485) tail draws per block, so a per-lane loop runs max-over-32-lanes (~5 at k = 8)
        // iterations of the ~45-instruction tail path with most lanes idle.  Instead every lane queues the slab
        // offsets of its tail draws (the queue aliases the staged rows, dead in this phase) and the warp drains the
        // queue 32 entries at a time (~2 iterations).
        {
#if MCP_LEAN_PUSH32
            const int cnt = __popc(tlo) + __popc(thi);
#else
            const int cnt = __popcll(tmask);
#endif
#if MCP_LEAN_SCAN
            // exclusive prefix sum of the lanes' counts: five shuffles instead of up to 32 serialised shared-memory atomics
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)(threadIdx.x & 31u) >= d) incl += v;
            }
            unsigned slot = (unsigned)(incl - cnt);
            const unsigned total = (unsigned)__shfl_sync(0xffffffffu, incl, 31);
#else
            unsigned slot = cnt ? atomicAdd(wq_cnt, (unsigned)cnt) : 0u;
#endif
#if MCP_LEAN_PUSH32
#pragma unroll PUSH_UNR
            while (tlo) {
                const int bit = __ffs((int)tlo) - 1;
                tlo &= tlo - 1u;
                wq[slot++] = (unsigned short)(own_off + 128 * (bit >> 2) + (bit & 3));
            }
#pragma unroll PUSH_UNR
            while (thi) {
                const int bit = __ffs((int)thi) - 1;
                thi &= thi - 1u;
                const int c = bit >> 2;
                wq[slot++] = (unsigned short)(own_off + 128 * (c == 7 ? PT - 1 : 8 + c) + (bit & 3));
            }
#else
            while (tmask) {
                const int bit = __ffsll((long long)tmask) - 1;
                tmask &= tmask - 1ull;
                const int c = bit >> 2;
                wq[slot++] = (unsigned short)(own_off + 128 * (c == 15 ? PT - 1 : c) + (bit & 3));
            }
#endif
            __syncwarp();
#if !MCP_LEAN_SCAN
            const unsigned total = *(volatile unsigned*)wq_cnt;
#endif
            for (unsigned e = threadIdx.x & 31u; e < total; e += 32u) {
                float* const p = wslab + wq[e];
                *p = ni_tail_inl(*p);
            }
            __syncwarp();
#if !MCP_LEAN_SCAN
            if ((threadIdx.x & 31u) == 0u) *wq_cnt = 0u;
#endif
        }
        // ---- C. lane q finishes row row0 + q (generated by lane qq) — Cholesky mix (MIX), interaction products
        // (INTER), linear predictor, y — and stages it in f64 (widened ONCE) as [z_0..z_{M-1} | z_N .. z_{N-M+1}]
        auto finish_row = [&](const int qq, const int row0) {
            {
                float* const src = wslab + qq * 32 + ((qi ^ (2 * qq)) << 2) + q;  // + 128·column
                double z[PT + 1];
                z[0] = 1.0;
                if (MIX) {  // x_j = f32(Σ_{m<=j} L[j][m] z_m), f64 accumulator, m ascending, no FMA (data_gen.rs:334-363)
                    double zr[PT];
#pragma unroll
                    for (int j = 0; j < PT - 1; j++) zr[j] = j < k ? (double)src[128 * j] : 0.0;
#pragma unroll
                    for (int j = 0; j < PT - 1; j++) {
                        if (j < k) {
                            double a = 0.0;
#pragma unroll
                            for (int m = 0; m <= j; m++) a = a + pl.L[j * MCP_MAX_K + m] * zr[m];
                            const float xf = (float)a;
                            if (INTER) src[128 * j] = xf;
                            z[j + 1] = (double)xf;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < PT - 1; j++)
                        if (j < k) z[j + 1] = (double)src[128 * j];
                }
                if (FACT) {  // factor dummies from the exact largest-remainder walk (data_gen.rs:509-551, no RNG): the
                             // host-built level table is sim-invariant
#pragma unroll
                    for (int j = 0; j < PT - 1; j++) {
                        const int f = pl.col_factor[j + 1];
                        if (j >= k && f >= 0) {
                            const float xf = pl.levels[(size_t)f * max_n + (row0 + q)] == (uint8_t)pl.col_level[j + 1] ? 1.0f : 0.0f;
                            if (INTER) src[128 * j] = xf;
                            z[j + 1] = (double)xf;
                        }
                    }
                }
                if (INTER) {  // x = f32(Π f64(components)) (data_gen.rs:643-649); columns after the dummies
#pragma unroll
                    for (int j = 0; j < PT - 1; j++) {
                        if (j >= k + pl.n_fd) {
                            double prod = 1.0;
                            for (int c = pl.inter_ptr[j - k - pl.n_fd]; c < pl.inter_ptr[j - k - pl.n_fd + 1]; c++)
                                prod *= (double)src[128 * (pl.inter_cols[c] - 1)];
                            const float xf = (float)prod;
                            src[128 * j] = xf;
                            z[j + 1] = (double)xf;
                        }
                    }
                }
                double lp = 0.0;
#if MCP_LEAN_DYNLP == 2
                if (!MIX && !INTER && !FACT) {
                    // every column, no test: a zero effect contributes z·(±0) = ±0, and a +0-seeded sum of such terms never
                    // reaches -0 (round to nearest), so s + ±0 == s bit for bit — the same value as skipping the term
#pragma unroll
                    for (int j = 0; j < PT; j++) lp = lp + z[j] * pl.beta[j];
                } else
#elif MCP_LEAN_DYNLP
                if (!MIX && !INTER && !FACT) {
                    // every x of this row still sits in the slab: a uniform loop over the non-zero effects (ascending columns,
                    // separately rounded product and sum — the same chain as below)
                    for (int t = 0; t < pl.n_nz; t++) {
                        const int j = pl.nz_col[t];
                        const double zj = j == 0 ? 1.0 : (double)src[128 * (j - 1)];
                        lp = lp + zj * pl.nz_beta[t];
                    }
                } else
#endif
                {
#if MCP_LEAN_LP_ALL_EVERYWHERE
                    // the same unconditional chain in the MIX / FACT / INTER instantiations (every column of z is defined there:
                    // a column is continuous, a dummy or an interaction)
#pragma unroll
                    for (int j = 0; j < PT; j++) lp = lp + z[j] * pl.beta[j];
#else
                    const uint32_t bm = pl.beta_nonzero_mask;  // β_j == 0 terms add ±0 to a +0-seeded sum
#pragma unroll
                    for (int j = 0; j < PT; j++)
                        if (bm & (1u << j)) lp = lp + z[j] * pl.beta[j];
#endif
                }
                const double ef = (double)src[128 * (PT - 1)];
                z[PT] = (double)(float)(lp + ef + 0.0);  // same association as row_xy: (lp + e) + u_re, u_re = +0
                double2* const dst = reinterpret_cast<double2*>(stage + q * G::ROW);
#pragma unroll
                for (int j = 0; j < MH / 2; j++) {
                    const int j0 = 2 * j, j1 = 2 * j + 1;
                    // (padding slots are never read by the accumulation)
                    dst[j] = make_double2(j0 < M ? z[j0] : MCP_LEAN_PAD, j1 < M ? z[j1] : MCP_LEAN_PAD);
                    dst[HOFF / 2 + j] = make_double2(j0 < M ? z[N - j0] : MCP_LEAN_PAD, j1 < M ? z[N - j1] : MCP_LEAN_PAD);
                }
            }
        };
        if constexpr (MCP_LEAN_INTERIOR && !MIX && !INTER && !FACT) {
            if (16 * sb + 16 <= max_n && n_next > (uint32_t)(16 * sb + 16)) {
#pragma unroll
                for (int qq = 0; qq < 4; qq++) {
                    finish_row(qq, 16 * sb + 4 * qq);
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < 4; r++) lean_acc_row<PT>(p_base + r * G::ROW, r_base + r * G::ROW, acc);
                    __syncwarp();
                }
                continue;
            }
        }
#pragma unroll 1
        for (int qq = 0; qq < 4; qq++) {
            const int row0 = 16 * sb + 4 * qq;
            if (row0 >= max_n) break;
            if (row0 + q < max_n) finish_row(qq, row0);
            __syncwarp();
            // ---- D. row pass: the same local pattern in every lane
            const int nr = max_n - row0 < 4 ? max_n - row0 : 4;
            if (nr == 4 && n_next - (uint32_t)row0 > 4u) {
#pragma unroll
                for (int r = 0; r < 4; r++) lean_acc_row<PT>(p_base + r * G::ROW, r_base + r * G::ROW, acc);
            } else {
#pragma unroll 1
                for (int r = 0; r < nr; r++) {
                    lean_acc_row<PT>(p_base + r * G::ROW, r_base + r * G::ROW, acc);
                    if ((uint32_t)(row0 + r + 1) == n_next) {
                        // grid point g: every lane writes its quarter of Z'Z straight to the wave's statistics buffer,
                        // stats[(g·E + e)·n_sims + sim] (coalesced over the studies of a warp); the solve runs in
                        // ols_solve_kernel, one thread per fit
                        if (valid) {
                            double* const out = stats + (size_t)g * G::E * n_sims + sim;
#pragma unroll
                            for (int a = 0; a < M; a++) {
#pragma unroll
                                for (int c = 0; c <= a; c++) {
                                    const int i = q == 0 ? a : (q == 1 ? N - a : N - c);
                                    const int j = q == 0 ? c : (q == 1 ? c : (q == 2 ? N - a : a));
                                    const int e = i < N ? tri(i, j) : NACC + j;  // X'X | X'y | y'y (j == N)
                                    out[(size_t)e * n_sims] = acc[tri(a, c)];
                                }
                            }
                        }
                        g++;
                        n_next = g < pl.n_ss ? pl.ss[g] : 0xffffffffu;
                    }
                }
            }
            __syncwarp();
        }
    }
    mcp_report_warp(blockIdx.x * G::STUDIES + (uint32_t)warp * 8u, 8u, n_sims);
}

// One OLS fit per thread from the statistics ols_lean_kernel left in HBM: Cholesky, β, σ², t², decisions,
// corrections, overall F, counters / byte tensors (ols_fit_and_emit — the same code the thread-per-study kernel runs).
// grid = (⌈n_sims / MCP_BLOCK⌉, n_ss); Σy is X'y[0] (x_0 ≡ 1).
#ifndef MCP_SOLVE_MIN_BLOCKS
#define MCP_SOLVE_MIN_BLOCKS 3  // measured on the 50-point sweep (P = 11): 2 → 6.13e8, 3 → 6.25e8, 4 → 5.69e8 fits/s
#endif
template <int PT>
__global__ void __launch_bounds__(MCP_BLOCK, MCP_SOLVE_MIN_BLOCKS) ols_solve_kernel(uint32_t n_sims, const double* __restrict__ stats) {
    constexpr int NACC = PT * (PT + 1) / 2;
    constexpr int E = LeanGeom<PT>::E;
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
    const int g = blockIdx.y;
    if (sim < n_sims) {
        const double* const in = stats + (size_t)g * E * n_sims + sim;  // entry e at in[e * n_sims]: coalesced over the warp
        const double yty = in[(size_t)(NACC + PT) * n_sims];
        const double sum_y = in[(size_t)NACC * n_sims];
        uint8_t excl[MCP_MAX_FACTORS];
        for (int f = 0; f < MCP_MAX_FACTORS; f++) excl[f] = 0;
        SimOut so{cnt_s, (uint64_t)sim};
        // sparse-factor exclusion (batch.rs:2451-2562): with exact walks the per-level counts at a grid point are the
        // same for every study, so the host tabulates them once; otherwise (sampled / bootstrapped factors) they are read
        // off the statistics — a reference-coded dummy is exactly 0 / 1, so the count of level l >= 1 is the column sum
        // X'X[dummy_l][0] (an exact integer in f64) and level 0 takes the remaining rows
        const bool track_excl = pl.n_factors > 0 && pl.min_level_count > 0;
        uint32_t lc[MCP_MAX_LEVELS];
        const uint32_t* counts = nullptr;
        if (track_excl) {
            if (pl.exact_counts) {
                counts = pl.exact_counts + (size_t)g * MCP_MAX_LEVELS;
            } else {
                for (int f = 0; f < pl.n_factors; f++) {
                    uint32_t rest = pl.ss[g];
                    for (int l = 1; l < pl.f_levels[f]; l++) {
                        const int col = pl.f_col[f] + l - 1;
                        const uint32_t c = (uint32_t)in[(size_t)tri(col, 0) * n_sims];
                        lc[pl.f_off[f] + l] = c;
                        rest -= c;
                    }
                    lc[pl.f_off[f]] = rest;
                }
                counts = lc;
            }
        }
        ols_fit_and_emit<PT>(in, yty, sum_y, pl.ss[g], so, g, counts, excl, track_excl, (size_t)n_sims);
    }
    if (cnt_s) {
        __syncthreads();
        for (int i = threadIdx.x; i < pl.n_ss * pl.counter_width; i += MCP_BLOCK) {
            const unsigned v = cnt_s[i];
            if (v) atomicAdd(&pl.counters[i], (unsigned long long)v);
        }
    }
}

}  // namespace mcp
