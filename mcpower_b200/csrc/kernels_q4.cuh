// kernels_q4.cuh — the OLS power kernel, lane-quad variant (Q4): one simulated study per FOUR adjacent lanes.
//
// Why (profiles/r01b_critical_path.md): with one study per thread the P(P+1)/2 + P + 2 FP64 accumulators (56 at P = 9)
// pin ols_power_kernel at 255 registers → 2 warps per scheduler → ~49 % issue utilisation, and ~88 % of a study's
// instructions (Philox + inverse normal) run while those 112 registers just sit there.  Here a study's state is spread
// over a quad, so each thread holds a quarter of the accumulators and the kernel fits 4–5 warps per scheduler:
//
//   * generation is split by ROWS, not columns: lane q of the quad runs the unmodified gen_block() for Philox block
//     4·sb + q (rows 16·sb + 4q … +3) into its own f32 slab — perfectly balanced for any column count;
//   * rows are then consumed in order, four at a time: in sub-step qq lane q finishes row 16·sb + 4·qq + q from lane
//     qq's slab (widen to f64 once, linear predictor, outcome — row_xy()) and stages it in shared memory as
//     Z = [1 | x_1 … x_{P−1} | y] in f64, laid out as [z_0 … z_{M−1} | z_N, z_{N−1}, … z_{N−M+1}] (N = P, M = ⌈(P+1)/2⌉):
//     the low half forwards and the high half REVERSED;
//   * every lane accumulates the SAME local pattern acc[a][c] += p[a]·r[c], c ≤ a < M, where p and r are the low
//     or the reversed-high half depending on the lane:
//         lane 0: p = lo, r = lo  → Z'Z[a][c]          (top-left triangle)
//         lane 1: p = hi, r = lo  → Z'Z[N−a][c]        (bottom-left square, on and below its diagonal)
//         lane 2: p = hi, r = hi  → Z'Z[N−c][N−a]      (bottom-right triangle)
//         lane 3: p = lo, r = hi  → Z'Z[N−c][a]        (bottom-left square, on and above its diagonal)
//     one instruction stream, lane-dependent shared-memory base addresses, M(M+1)/2 accumulators per lane (15 at
//     P = 9).  X'y and y'y are the last row of Z'Z; Σy is its (N, 0) entry (x_0 ≡ 1, so fma(1, y, acc) ≡ acc + y).
// Every accumulator is still ONE sequential fma chain over the rows, so the statistics are bit-identical to the
// thread-per-study kernel and to the oracle.  At a grid point the quad hands its entries to lane 0 by shuffle and
// lane 0 runs the solve (3 of 4 lanes idle there: the host picks this kernel only when fits are rare relative to
// rows — single-N power runs like the headline shape; dense sample-size sweeps keep the thread-per-study kernel).
#pragma once
#include "kernels.cuh"

namespace mcp {

#ifndef MCP_Q4_MIN_BLOCKS
#define MCP_Q4_MIN_BLOCKS 4
#endif

template <int PT>
struct Q4Geom {
    static constexpr int N = PT;                 // index of y in Z
    static constexpr int M = (PT + 2) / 2;       // slots per half = ceil((P+1)/2)
    static constexpr int MH = (M + 1) & ~1;      // padded to an even count (16-byte vector loads)
    static constexpr int HOFF = mcp_stage_geom(M).hoff;    // offset of the reversed high half inside a staged row
    static constexpr int ROW = mcp_stage_geom(M).row;      // doubles per staged row
    static constexpr int STUDY = mcp_stage_geom(M).study;  // doubles per study: 4 staged rows; strides chosen conflict-free (plan.h)
    static_assert(HOFF >= MH && ROW >= HOFF + MH && STUDY >= 4 * ROW, "staging geometry");
    static constexpr int NACCQ = M * (M + 1) / 2;
    static constexpr int STUDIES = MCP_BLOCK / 4;
    // owner lane / local (a, c) of the true entry Z'Z[i][j], j <= i <= N
    __host__ __device__ static constexpr int owner(int i, int j) {
        return i < M ? 0 : (j >= N - M + 1 ? 2 : (j <= N - i ? 1 : 3));
    }
    __host__ __device__ static constexpr int la(int i, int j) {
        return i < M ? i : (j >= N - M + 1 ? N - j : (j <= N - i ? N - i : j));
    }
    __host__ __device__ static constexpr int lc(int i, int j) {
        return i < M ? j : (j >= N - M + 1 ? N - i : (j <= N - i ? j : N - i));
    }
    static size_t smem_bytes(int n_ss, int counter_width, bool smem_counters) {
        size_t b = (size_t)MCP_BLOCK * 4 * (PT | 1) * sizeof(float) + (size_t)STUDIES * STUDY * sizeof(double);
        if (smem_counters) b += (size_t)n_ss * counter_width * sizeof(unsigned);
        return b;
    }
};

// One staged row into the lane's local pattern: acc[a][c] += p[a]·r[c], c <= a < M.
template <int PT>
__device__ __forceinline__ void q4_acc_row(const double* __restrict__ prow, const double* __restrict__ rrow,
                                           double (&acc)[Q4Geom<PT>::NACCQ]) {
    constexpr int M = Q4Geom<PT>::M, MH = Q4Geom<PT>::MH;
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

// Hand the quad's accumulators to every lane (lane 0 uses them): loop over the LOCAL pattern (static register
// indices) and the four owners; the true Z'Z index only addresses the memory-resident st[] array.
//   owner 0: Z'Z[a][c]   owner 1: Z'Z[N-a][c]   owner 2: Z'Z[N-c][N-a]   owner 3: Z'Z[N-c][a]
template <int PT>
__device__ __forceinline__ void q4_gather(const double (&acc)[Q4Geom<PT>::NACCQ], unsigned qmask, double* st, double& yty) {
    constexpr int N = Q4Geom<PT>::N, M = Q4Geom<PT>::M, NACC = PT * (PT + 1) / 2;
#pragma unroll
    for (int a = 0; a < M; a++) {
#pragma unroll
        for (int c = 0; c <= a; c++) {
#pragma unroll
            for (int o = 0; o < 4; o++) {
                const double v = __shfl_sync(qmask, acc[tri(a, c)], o, 4);
                const int i = o == 0 ? a : (o == 1 ? N - a : N - c);
                const int j = o == 0 ? c : (o == 1 ? c : (o == 2 ? N - a : a));
                if (i < N) st[tri(i, j)] = v;
                else if (j < N) st[NACC + j] = v;
                else yty = v;
            }
        }
    }
}

// Per-level counts of slab row r (batch.rs:2451-2496); each lane of the quad counts one row of a sub-step.
static __device__ __noinline__ void q4_count_levels(const Slab<MCP_BLOCK> ps, int r, uint32_t* lvl_counts) {
    const DevicePlan& pl = c_plan;
    for (int f = 0; f < pl.n_factors; f++) {
        const int c0 = pl.f_col[f] - 1, nl = pl.f_levels[f];
        int lvl = 0;
        for (int d = 0; d < nl - 1; d++)
            if (ps[4 * (c0 + d) + r] == 1.0f) { lvl = d + 1; break; }
        lvl_counts[pl.f_off[f] + lvl]++;
    }
}

// Warp-cooperative marginal transforms of one generation step (scenario path).  Every lane of the warp has just generated and
// mixed a 4-row block of ITS study (quad); which continuous columns carry a non-Normal marginal is a per-study draw
// (data_gen.rs:1040-1060), so a per-lane column loop leaves the lanes of studies whose column j stayed Normal idle (19-22 of 32
// active, profiles/r01k_split_path.md).  Here the warp's (lane, column) pairs that need a transform are listed in shared memory
// (exclusive prefix sum of the per-lane counts) and dealt round-robin to all 32 lanes: each transforms four values of some
// lane's slab in place.  An element sees exactly the arithmetic of gen_block_pre's own loop, so the plane is bit-identical.
// `nn` = this lane's mask of non-Normal continuous columns (0 for a lane without a block this step).
template <int PT>
__device__ __forceinline__ void q4_coop_marginals(unsigned wmask, unsigned nn, const int* __restrict__ vt_of, float* __restrict__ smem,
                                                  uint16_t* __restrict__ items) {
    const DevicePlan& pl = c_plan;
    const int lane = threadIdx.x & 31;
    const int cnt = __popc(nn);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(wmask, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(wmask, incl, 31 - __clz(wmask));
    int pos = incl - cnt;  // (a rank-major list — every round 32 different owners — measured slower: the ballot loop costs more than the
                           //  bank conflicts of one owner's consecutive items, profiles/r02_config3.md)
    for (unsigned m = nn; m; m &= m - 1u) {
        const int j = __ffs(m) - 1;
        items[pos++] = (uint16_t)(lane | (j << 5) | (vt_of[j] << 10));
    }
    __syncwarp(wmask);
    for (int t = lane; t < total; t += 32) {
        const unsigned it = items[t];
        const int j = (it >> 5) & 31, vt = it >> 10;
        const Slab<MCP_BLOCK> ps{smem + (threadIdx.x & ~31u) + (it & 31u)};
        const float2 lo = apply_marginal2(ps[4 * j + 0], ps[4 * j + 1], vt | (j << 8), pl.var_param[j], pl.t3_table);
        const float2 hi = apply_marginal2(ps[4 * j + 2], ps[4 * j + 3], vt | (j << 8), pl.var_param[j], pl.t3_table);
        ps[4 * j + 0] = lo.x;
        ps[4 * j + 1] = lo.y;
        ps[4 * j + 2] = hi.x;
        ps[4 * j + 3] = hi.y;
    }
    __syncwarp(wmask);
}

// SPLIT = true: generate + accumulate only.  At a grid point every lane writes its quarter of Z'Z straight to the wave's
// statistics buffer stats[(g·E + e)·n_sims + sim] (the lean path's layout, coalesced over the studies of a warp) and
// ols_solve_kernel (kernels_lean.cuh) fits from there, one thread per fit.  No solver call in the streaming kernel, so
// ptxas does not park the accumulators in local memory around it, and no per-level counting either: the dummies are exact
// 0 / 1 (data_gen.rs:545-549), so the per-level counts of batch.rs:2451-2496 ARE the dummy column sums X'X[dummy][0]
// (exact integers in f64), and the solve kernel reads them from the statistics.
// The SPLIT scenario path also takes the per-study scenario design (perturbed correlation → PSD repair → Cholesky, marginal
// and residual swaps, jittered effects) from scen_design_kernel below instead of deriving it in every lane of the quad: the
// Jacobi repair runs once per study, in its own kernel, and the streaming kernel reads the factor through the caches.
// One thread per study: the per-study scenario design of populate_design (data_gen.rs:990-1093) into HBM.
// TAG = the translation unit's kernel family (OLS / GLM / LMM): every family launches this kernel for the same P, and identical
// template instantiations in two translation units are merged by the linker — the launch could then run the other module's copy,
// which reads THAT module's __constant__ plan (see the note in kernel_tu.cu).  The tag keeps the instantiations distinct.
template <int PT, int TAG>
__global__ void __launch_bounds__(MCP_BLOCK) scen_design_kernel(uint64_t base_seed, uint64_t sim_begin, uint32_t n_sims, ScenDesign<PT>* __restrict__ out) {
    const uint32_t sim = blockIdx.x * MCP_BLOCK + threadIdx.x;
    if (sim >= n_sims || mcp_cancelled()) return;
    ScenDesign<PT> sd;
    scen_populate<PT>(sd, base_seed, sim_begin + sim);
    if (!sd.ok) atomicExch(c_plan.error_flag, 2);  // EngineError::CorrelationNotPSD (data_gen.rs:1049)
    out[sim] = sd;
}

//
// CLUS = true (with SPLIT; the mixed-model arm's generator): the study's per-cluster sums of [1 | x | y] also live in shared
// memory, csm[cluster][column].  After a sub-step's four rows are staged, lane q adds them — rows in order, so every sum is
// the same sequential chain as in the oracle — into the columns j ≡ q (mod 4); at a grid point the quad copies the sums of
// the clusters seen so far to clus_snap[(g · clus_doubles + e) · stride + study] for lmm_fit_kernel.  `stat_stride` is the
// study stride of both buffers (0 → n_sims, the OLS layout).
//
// PLANE = true (with SPLIT; the GLM arm's generator): nothing is accumulated — lane q writes the finished row row0 + q
// (x_1 … x_{P−1}, y as f32) to the data plane plane[(row · PT + column) · stride + study] that glm_fit_kernel re-reads on
// every IRLS pass.
template <int PT, bool SCEN, bool SPLIT = false, bool CLUS = false, bool PLANE = false>
__global__ void __launch_bounds__(MCP_BLOCK, MCP_Q4_MIN_BLOCKS) ols_power_kernel_q4(uint64_t base_seed, uint64_t sim_begin, uint32_t n_sims,
                                                                                    double* __restrict__ stats = nullptr,
                                                                                    const ScenDesign<PT>* __restrict__ designs = nullptr,
                                                                                    size_t stat_stride = 0, double* __restrict__ clus_snap = nullptr,
                                                                                    int clus_doubles = 0, float* __restrict__ plane = nullptr) {
    static_assert(!CLUS || SPLIT, "cluster sums are only kept on the split path");
    static_assert(!PLANE || (SPLIT && !CLUS), "the data plane is written on the split path");
    using G = Q4Geom<PT>;
    constexpr int NACC = PT * (PT + 1) / 2;
    constexpr int SP = PT | 1;
    constexpr int N = G::N, M = G::M, MH = G::MH, HOFF = G::HOFF;
    const DevicePlan& pl = c_plan;
    if (mcp_cancelled()) return;
    extern __shared__ __align__(16) float smem[];
    const int q = threadIdx.x & 3;
    const Slab<MCP_BLOCK> slab{smem + threadIdx.x};
    double* const stage = reinterpret_cast<double*>(smem + (size_t)MCP_BLOCK * 4 * SP) + (size_t)(threadIdx.x >> 2) * G::STUDY;
    // odd study stride: the eight quads of a warp hit eight different bank pairs
    double* const csm = reinterpret_cast<double*>(smem + (size_t)MCP_BLOCK * 4 * SP) + (size_t)G::STUDIES * G::STUDY +
                        (size_t)(threadIdx.x >> 2) * (size_t)(clus_doubles | 1);
    const size_t sstride = stat_stride ? stat_stride : (size_t)n_sims;
    unsigned int* cnt_s = nullptr;
    if (!SPLIT && pl.smem_counters) {  // SPLIT emits nothing: the solve kernel owns the counters
        cnt_s = reinterpret_cast<unsigned int*>(reinterpret_cast<double*>(smem + (size_t)MCP_BLOCK * 4 * SP) + (size_t)G::STUDIES * G::STUDY);
        for (int i = threadIdx.x; i < pl.n_ss * pl.counter_width; i += MCP_BLOCK) cnt_s[i] = 0u;
        __syncthreads();
    }
    const uint32_t sim = blockIdx.x * G::STUDIES + (threadIdx.x >> 2);
    // the scenario path's warp-cooperative marginal pass: (lane, column, kind) items of one generation step
    __shared__ uint16_t coop_items[SCEN ? MCP_BLOCK * (PT > 1 ? PT - 1 : 1) : 1];
    // quads never straddle a warp; a quad past the end idles as a whole, so the quad-masked syncs below see full quads
    if (sim < n_sims) {
        const unsigned qmask = 0xFu << ((threadIdx.x & 31u) & ~3u);
        // the warp's live lanes: its studies are consecutive, so the quads past the end of the wave are the highest ones
        const uint32_t warp_first = blockIdx.x * G::STUDIES + ((threadIdx.x >> 5) << 3);
        const unsigned wmask = n_sims - warp_first >= 8u ? 0xffffffffu : (1u << (4u * (n_sims - warp_first))) - 1u;
        const uint64_t key = pcg_mix64(base_seed, sim_begin + sim);
        const int max_n = pl.max_n;
        SimOut so{cnt_s, (uint64_t)sim};
        ScenDesign<PT> sdv;
        const ScenDesign<PT>* sd = nullptr;
        if (SCEN) {
            if (SPLIT) {
                sd = designs + sim;  // scen_design_kernel ran first
            } else {  // every lane of the quad derives the same per-study design (pure function of the sim id)
                scen_populate<PT>(sdv, base_seed, sim_begin + sim);
                sd = &sdv;
                if (!sdv.ok) atomicExch(pl.error_flag, 2);
            }
        }
        SimCtx cx;
        init_sim_ctx<PT>(cx, key, sim, sd);
        double acc[G::NACCQ];
#pragma unroll
        for (int i = 0; i < G::NACCQ; i++) acc[i] = 0.0;
        const int clus_cs = pl.sizing_value > 0 ? pl.sizing_value : 1;
        if (CLUS) {
            for (int e = q; e < clus_doubles; e += 4) csm[e] = 0.0;
            // the study's cluster effects u_c = f32(normal(word c) · tau) (data_gen.rs:146-161), tabulated once instead of one
            // Philox block + inverse normal per row; clusters that were never drawn (trailing partial cluster) have none
            const int nc_cap = clus_doubles / (PT + 1);
            double* const utab = reinterpret_cast<double*>(smem + (size_t)MCP_BLOCK * 4 * SP) + (size_t)G::STUDIES * G::STUDY +
                                 (size_t)G::STUDIES * (size_t)(clus_doubles | 1) + (size_t)(threadIdx.x >> 2) * (size_t)(nc_cap | 1);
            for (int c = q; c < nc_cap; c += 4) utab[c] = c < pl.n_clusters ? cluster_effect(cx, (uint32_t)c) : 0.0;
            cx.utab = utab;
            __syncwarp(qmask);
        }
        // cluster sums of staged row r (row index row0 + r): this lane's columns, plain adds (data_gen.rs cluster_of_row)
        auto clus_add = [&](int row, const double* zr) {
            const int kc = pl.sizing_fixed_size ? row / clus_cs : row % clus_cs;
            double* const base = csm + kc * (PT + 1);
            for (int j = q; j <= PT; j += 4) base[j] += j < M ? zr[j] : zr[HOFF + (N - j)];
        };
        uint32_t lvl_counts[MCP_MAX_LEVELS];  // this lane's share (rows with r == q); summed over the quad at a fit
        uint8_t excl[MCP_MAX_FACTORS];
        const bool track_excl = !SPLIT && pl.n_factors > 0 && pl.min_level_count > 0;
        if (track_excl)
            for (int i = 0; i < MCP_MAX_LEVELS; i++) lvl_counts[i] = 0;
        for (int f = 0; f < MCP_MAX_FACTORS; f++) excl[f] = 0;
        int g = 0;
        uint32_t n_next = pl.ss[0];
        // lane-dependent halves of the staged row
        const double* const p_base = stage + ((q == 1 || q == 2) ? HOFF : 0);
        const double* const r_base = stage + (q >= 2 ? HOFF : 0);
        unsigned nn_cols = 0u;  // this study's non-Normal continuous columns
        if (SCEN)
            for (int j = 0; j < pl.k; j++) nn_cols |= (sd->vt[j] != D_NORMAL ? 1u : 0u) << j;
        const int n_qsteps = (max_n + 15) >> 4;
        for (int sb = 0; sb < n_qsteps; sb++) {
            {
                const bool mine = 16 * sb + 4 * q < max_n;
                if (SCEN) {
                    if (mine) gen_block_pre<PT, SCEN, 1>(4 * sb + q, slab, cx, sd);
                    if (pl.coop_marginals) {
                        __syncwarp(wmask);
                        q4_coop_marginals<PT>(wmask, mine ? nn_cols : 0u, sd->vt, smem, coop_items + (threadIdx.x >> 5) * 32 * (PT > 1 ? PT - 1 : 1));
                    } else if (mine) {
                        gen_marginals<PT, SCEN>(slab, sd);
                    }
                    if (mine) gen_block_pre<PT, SCEN, 2>(4 * sb + q, slab, cx, sd);
                    if (pl.n_between > 0) __syncwarp(qmask);
                    if (mine) gen_block_post<PT, SCEN>(4 * sb + q, slab, cx);
                } else if (pl.n_between == 0) {
                    if (mine) gen_block<PT, SCEN>(4 * sb + q, slab, cx, sd);
                } else {  // cluster-level predictors: the step's representative rows are stored before any lane loads one
                    if (mine) gen_block_pre<PT, SCEN, 0>(4 * sb + q, slab, cx, sd);
                    __syncwarp(qmask);
                    if (mine) gen_block_post<PT, SCEN>(4 * sb + q, slab, cx);
                }
            }
            __syncwarp(qmask);
#pragma unroll 1
            for (int qq = 0; qq < 4; qq++) {
                const int row0 = 16 * sb + 4 * qq;
                if (row0 >= max_n) break;
                const Slab<MCP_BLOCK> ps{smem + (threadIdx.x & ~3u) + qq};  // slab of the lane that generated these rows
                // ---- Y. lane q finishes row row0 + q and stages it in f64
                if (row0 + q < max_n) {
                    double z[PT + 1];
                    {
                        double xr[PT];
                        const double yd = row_xy<PT, SCEN>(4 * sb + qq, q, ps, cx, xr, sd);
#pragma unroll
                        for (int j = 0; j < PT; j++) z[j] = xr[j];
                        z[PT] = yd;
                    }
                    if (PLANE) {
                        float* const rowp = plane + (size_t)(row0 + q) * PT * sstride + sim;
#pragma unroll
                        for (int j = 1; j <= PT; j++) rowp[(size_t)(j - 1) * sstride] = (float)z[j];
                    } else {
                        double2* const dst = reinterpret_cast<double2*>(stage + q * G::ROW);
#pragma unroll
                        for (int j = 0; j < MH / 2; j++) {
                            const int j0 = 2 * j, j1 = 2 * j + 1;
                            dst[j] = make_double2(z[j0 < M ? j0 : 0], z[j1 < M ? j1 : 0]);
                            dst[HOFF / 2 + j] = make_double2(z[j0 < M ? N - j0 : 0], z[j1 < M ? N - j1 : 0]);
                        }
                    }
                }
                __syncwarp(qmask);
                if (PLANE) continue;
                // ---- E. row pass: the same local pattern in every lane
                const int nr = max_n - row0 < 4 ? max_n - row0 : 4;
                if (nr == 4 && n_next - (uint32_t)row0 > 4u) {
                    // fast path: four rows, no grid point inside
#pragma unroll
                    for (int r = 0; r < 4; r++) q4_acc_row<PT>(p_base + r * G::ROW, r_base + r * G::ROW, acc);
                    if (CLUS) {
#pragma unroll 1
                        for (int r = 0; r < 4; r++) clus_add(row0 + r, stage + r * G::ROW);
                    }
                    if (track_excl) q4_count_levels(ps, q, lvl_counts);
                } else {
#pragma unroll 1
                    for (int r = 0; r < nr; r++) {
                        q4_acc_row<PT>(p_base + r * G::ROW, r_base + r * G::ROW, acc);
                        if (CLUS) clus_add(row0 + r, stage + r * G::ROW);
                        if (track_excl && r == q) q4_count_levels(ps, r, lvl_counts);
                        if (SPLIT && (uint32_t)(row0 + r + 1) == n_next) {
                            // ---- grid point g: this lane's quarter of Z'Z → the statistics buffer (entry e = X'X packed lower
                            // | X'y | y'y, as LeanGeom<PT>::E)
                            double* const out = stats + (size_t)g * (NACC + PT + 1) * sstride + sim;
#pragma unroll
                            for (int a = 0; a < M; a++) {
#pragma unroll
                                for (int c = 0; c <= a; c++) {
                                    const int i = q == 0 ? a : (q == 1 ? N - a : N - c);
                                    const int j = q == 0 ? c : (q == 1 ? c : (q == 2 ? N - a : a));
                                    const int e = i < N ? tri(i, j) : NACC + j;  // j == N: y'y
                                    out[(size_t)e * sstride] = acc[tri(a, c)];
                                }
                            }
                            if (CLUS) {  // snapshot of the cluster sums after the first row0 + r + 1 rows
                                const int n_now = row0 + r + 1;
                                const int ncl = pl.sizing_fixed_size ? (n_now + clus_cs - 1) / clus_cs : clus_cs;
                                __syncwarp(qmask);
                                double* const sn = clus_snap + (size_t)g * clus_doubles * sstride + sim;
                                for (int e = q; e < ncl * (PT + 1); e += 4) sn[(size_t)e * sstride] = csm[e];
                                __syncwarp(qmask);
                            }
                            g++;
                            n_next = g < pl.n_ss ? pl.ss[g] : 0xffffffffu;
                        } else if ((uint32_t)(row0 + r + 1) == n_next) {
                            // ---- fit at grid point g: the quad hands its entries to lane 0, which solves through memory
                            double st[NACC + PT];
                            double yty = 0.0;
                            q4_gather<PT>(acc, qmask, st, yty);
                            uint32_t lc_tot[MCP_MAX_LEVELS];
                            if (track_excl) {
                                int n_lv = 0;
                                for (int f = 0; f < pl.n_factors; f++) n_lv += pl.f_levels[f];
                                for (int i = 0; i < n_lv; i++) {
                                    uint32_t v = lvl_counts[i];
                                    v += __shfl_xor_sync(qmask, v, 1);
                                    v += __shfl_xor_sync(qmask, v, 2);
                                    lc_tot[i] = v;
                                }
                            }
                            if (q == 0) ols_fit_and_emit<PT>(st, yty, st[NACC], (uint32_t)(row0 + r + 1), so, g, lc_tot, excl, track_excl);
                            __syncwarp(qmask);
                            g++;
                            n_next = g < pl.n_ss ? pl.ss[g] : 0xffffffffu;
                        }
                    }
                }
                __syncwarp(qmask);
            }
        }
    }
    // progress is counted where the OLS arm spends its time; the GLM / mixed-model arms count finished fits in their fit kernels
    if (!CLUS && !PLANE) mcp_report_warp(blockIdx.x * G::STUDIES + ((threadIdx.x & ~31u) >> 2), 8u, n_sims);
    if (cnt_s) {
        __syncthreads();
        for (int i = threadIdx.x; i < pl.n_ss * pl.counter_width; i += MCP_BLOCK) {
            const unsigned v = cnt_s[i];
            if (v) atomicAdd(&pl.counters[i], (unsigned long long)v);
        }
    }
}

}  // namespace mcp
