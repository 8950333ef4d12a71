// Visible -> hidden pass of the cRBM on packed 2-bit DNA, both strands, as a k-mer table gather.
//
// Reference semantics (secomo/convRBM.py):
//   A1 :238-243  x[n,k,j]  = b_k + sum_m W[k, base(n,j+m), m]            (forward strand)
//                x'[n,k,j] = b_k + sum_m W[k, 3-base(n,j+m), M-1-m]      (flip_motif=True: reverse complement)
//   A2 :245-257  p = exp(x) / (P + sum_{pool group} exp(x))
//   A3 :259-267  one uniform per pool group, first unit whose running sum exceeds it, none if u >= sum
//   A8-A10 :327-371  VH[k,a,m] = sum_{n,j} p[n,k,j] [base(n,j+m)==a]  (+ strand fold (a,m) -> (3-a, M-1-m))
//   A11 :440-451 Q = p (1 - sum_group p) on the forward strand (closed form of the T.grad sparsity term)
//   A13 :657-697 free energy: sum_groups log(1 + sum_group exp(x))
//
// Mapping to the SM. One lane = one motif k (32 motifs per k-tile), one warp walks a block of 16 consecutive
// hidden positions of one sequence, so the base under the read head is warp-uniform and every shared-memory
// access is a conflict-free row read:
//   * the motif bank tile (32 motifs x 4 x M floats, contiguous in W) is staged global->shared with one
//     cp.async.bulk (TMA engine) and expanded in shared memory into a 3-mer table
//         T[g][idx][lane] = (sum_{r<3} W[k, a_r, 3g+r],  sum_{r<3} Wrc[k, a_r, 3g+r]),   idx = a_0 | a_1<<2 | a_2<<4
//     (bias folded into g=0; Wrc[k,a,m] = W[k,3-a,M-1-m]; zero beyond M), MW = 3*ceil(M/3), G = MW/3;
//   * x and x' for a position are G LDS.64 + 2G FADD instead of 2M dependent one-hot products;
//   * the statistics are accumulated in registers, organised by sequence position so that the (warp-uniform)
//     base selects one of four fully unrolled accumulator updates: acc[a][m] += p[s-m].
//     The reverse-strand fold uses a base stream shifted by MW-M so that its target index MW-1-m is static.
#pragma once
#include "common.cuh"

namespace crbm {

enum : int { F_STAT_P = 1, F_STAT_Q = 2, F_SAMPLE = 4, F_FE = 8, F_COUNT = 16 };

struct V2HParams {
    const uint32_t *packed;   // rows of sequences
    int row_words, n_rows, L, Lh;
    const float *W, *b;       // (K,4,M), (K)
    int K, M, pooling, ds;
    int bv, nblk, n_items;    // valid block length (multiple of pooling), blocks per sequence, n_rows*nblk
    // outputs
    float *partial;           // [ktiles][gridDim.y][nacc][32]
    uint32_t *bits;           // [2][n_rows][ktiles][FANTASY_WORDS] sampled hidden bit planes (F_SAMPLE; layout: common.cuh)
    float *fe_motif;          // [n_rows][K] accumulated (F_FE), may be null
    float *fe_seq;            // [n_rows] accumulated (F_FE), may be null
    float *count;             // [n_rows] accumulated number of sampled forward units per sequence (F_COUNT)
    // randomness
    const float *u_fwd, *u_rc; // injected uniforms [(n*K+k)*(Lh/P)+group] or null
    uint64_t seed;
    uint32_t ctr_hi;          // training step (or the caller's offset): third word of the Philox counter
    uint32_t iter;            // Gibbs sweep of the step: bits [8,32) of the counter's tag word (any cd_k; the sweeps of
                              // consecutive steps never share a counter)
    // device-resident cursors (CUDA-graph replay): minibatch index into `packed`, training step
    const int *batch_index;
    const unsigned long long *step_dev;
    // programmatic dependent launch (common.cuh): 0 = wait for the predecessor grid before the main loop (the kernel reads
    // what it wrote); 1 = only before exit (independent of the predecessor: keeps stream order transitive, overlaps its tail)
    int pdl_late;
};

template <int MW> struct V2HShape {
    static constexpr int G = MW / 3;
    static constexpr int B = BLOCK_POS;
    static constexpr int T = B + MW - 3;          // 3-mer start offsets touched by one block
    static constexpr int S = B + MW - 1;          // sequence positions touched by one block
    static constexpr int TABLE_FLOAT2 = G * 64 * KTILE;
    static constexpr size_t table_bytes = (size_t)TABLE_FLOAT2 * sizeof(float2);
    static constexpr size_t wtile_bytes = (size_t)KTILE * 4 * CRBM_MAX_MOTIF_LEN * sizeof(float);
};

template <int MW, int FLAGS> struct V2HAcc {
    static constexpr int NP = (FLAGS & F_STAT_P) ? 4 * MW : 0;
    static constexpr int NQ = (FLAGS & F_STAT_Q) ? 4 * MW : 0;
    static constexpr int NACC = NP + NQ + 4;      // + sumP fwd, sumP rc, sumQ, spare
};

// Stage the W tile of k-tile `kt` (TMA bulk copy) and expand it into the 3-mer table. All threads call.
template <int MW>
__device__ __forceinline__ void build_table(float2 *tab, float *wtile, uint64_t *bar,
                                            const float *__restrict__ W, const float *__restrict__ b,
                                            int K, int M, int kt) {
    constexpr int G = MW / 3;
    const int k0 = kt * KTILE;
    const int nk = min(KTILE, K - k0);
    const uint32_t bytes = (uint32_t)nk * 4u * (uint32_t)M * 4u;       // multiple of 16
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        tma_load_1d(wtile, W + (size_t)k0 * 4 * M, bytes, bar);
    }
    mbar_wait(bar, 0);
    const int fourM = 4 * M;
    for (int e = threadIdx.x; e < G * 64 * KTILE; e += blockDim.x) {
        const int lane = e & 31, idx = (e >> 5) & 63, g = e >> 11;
        float f = 0.f, r = 0.f;
        if (lane < nk) {
            const float *w = wtile + lane * fourM;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int m = 3 * g + q, a = (idx >> (2 * q)) & 3;
                if (m < M) {
                    f += w[a * M + m];
                    r += w[(3 - a) * M + (M - 1 - m)];
                }
            }
            if (g == 0) {
                const float bk = __ldg(b + k0 + lane);
                f += bk;
                r += bk;
            }
        }
        tab[e] = make_float2(f, r);
    }
    __syncthreads();
}

// x[r], xr[r] for the 16 hidden positions of a block whose first base window is `a` (bases j0 .. j0+47).
template <int MW>
__device__ __forceinline__ void block_activities(const float2 *__restrict__ tab_lane, const uint32_t (&a)[3],
                                                 float (&x)[BLOCK_POS], float (&xr)[BLOCK_POS]) {
    constexpr int G = MW / 3, B = BLOCK_POS, T = B + MW - 3;
#pragma unroll
    for (int r = 0; r < B; ++r) { x[r] = 0.f; xr[r] = 0.f; }
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const uint32_t idx = window_bits<3, 3>(a, t);
        const float2 *p = tab_lane + idx * KTILE;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int r = t - 3 * g;
            if (r >= 0 && r < B) {
                const float2 v = p[g * 64 * KTILE];
                x[r] += v.x;
                xr[r] += v.y;
            }
        }
    }
}

struct PoolMasks { uint32_t start, end, valid; };   // bit r: first / last of its pool group / position is real

// start/end depend only on the pooling factor: computed once per kernel.
__device__ __forceinline__ PoolMasks make_pool_masks(int pooling) {
    PoolMasks m{0u, 0u, 0u};
    for (int r = 0; r < BLOCK_POS; ++r) {
        if (r % pooling == 0) m.start |= 1u << r;
        if (r % pooling == pooling - 1) m.end |= 1u << r;
    }
    return m;
}
__device__ __forceinline__ uint32_t valid_mask(int bv, int j0, int Lh) {
    const int n = min(bv, Lh - j0);
    return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
}

// In place: x[r] (activity) -> p[r] (pooled probability, 0 where invalid). Also returns, per position, the
// quantities later stages need: cum[r]*inv is the running sum of p inside the group; pe[r] = P*exp(-gmax)/norm
// (so Q = p*pe, and 1 - sum_group p = pe). fe += sum over groups of log(1 + sum exp(x)).
template <bool WANT_FE>
__device__ __forceinline__ void pool_block(float (&x)[BLOCK_POS], float (&cumn)[BLOCK_POS], float (&pe)[BLOCK_POS],
                                           const PoolMasks mk, float fpool, float &fe) {
    constexpr int B = BLOCK_POS;
    float gm[B];
    float run = 0.f;
#pragma unroll
    for (int r = 0; r < B; ++r) {
        if ((mk.start >> r) & 1u) run = 0.f;
        run = fmaxf(run, ((mk.valid >> r) & 1u) ? x[r] : 0.f);
        gm[r] = run;
    }
#pragma unroll
    for (int r = B - 1; r >= 0; --r) {
        if ((mk.end >> r) & 1u) run = gm[r];
        gm[r] = run;
    }
    float cum[B];
    run = 0.f;
#pragma unroll
    for (int r = 0; r < B; ++r) {
        const float e = ((mk.valid >> r) & 1u) ? __expf(x[r] - gm[r]) : 0.f;
        x[r] = e;
        if ((mk.start >> r) & 1u) run = 0.f;
        run += e;
        cum[r] = run;
    }
#pragma unroll
    for (int r = B - 1; r >= 0; --r) {
        if ((mk.end >> r) & 1u) run = cum[r];
        const float tot = run;
        const float eg = __expf(-gm[r]);
        const float inv = 1.0f / (fpool * eg + tot);
        if (WANT_FE) {
            if (((mk.end & mk.valid) >> r) & 1u) fe += (gm[r] == 0.f) ? log1pf(tot) : gm[r] + __logf(eg + tot);
        }
        x[r] *= inv;                 // probability
        cumn[r] = cum[r] * inv;      // running sum of probabilities within the group
        pe[r] = fpool * eg * inv;    // 1 - sum_group p
    }
}

// One uniform per pool group for rows (n,k): injected array or Philox keyed by (seed; row, group/4, tag).
struct GroupUniforms {
    const float *inj;      // already offset to this lane's row: inj[group]
    uint2 key;
    uint32_t row, tag, hi;
    uint4 cache;
    int cached_q;
    __device__ __forceinline__ float get(int group) {
        if (inj) return inj[group];
        const int q = group >> 2;
        if (q != cached_q) {
            cache = philox4x32(make_uint4(row, (uint32_t)q, hi, tag), key);
            cached_q = q;
        }
        const int c = group & 3;
        const uint32_t v = c == 0 ? cache.x : c == 1 ? cache.y : c == 2 ? cache.z : cache.w;
        return u01(v);
    }
};

// Sample within pool groups from probabilities p and their running sums; returns a bit mask over r.
__device__ __forceinline__ uint32_t sample_block(const float (&cumn)[BLOCK_POS], const PoolMasks mk,
                                                 GroupUniforms &gu, int group0) {
    uint32_t bits = 0u;
    bool done = false;
    float u = 0.f;
    int g = group0 - 1;
#pragma unroll
    for (int r = 0; r < BLOCK_POS; ++r) {
        if ((mk.start >> r) & 1u) {
            ++g;
            done = false;
            if ((mk.valid >> r) & 1u) u = gu.get(g);
        }
        const bool s = ((mk.valid >> r) & 1u) && !done && (cumn[r] > u);
        done = done || s;
        if (s) bits |= 1u << r;
    }
    return bits;
}

// Sampled units of one lane (= motif) for the block's positions -> bit planes (layout: common.cuh). The planes were
// zeroed before the launch; hits are rare (target sparsity rho), so one atomicOr per hit is cheap.
__device__ __forceinline__ void scatter_sample_bits(uint32_t *tile_bits, uint32_t sb, int j0, int lane) {
    while (sb) {
        const int r = __ffs(sb) - 1;
        sb &= sb - 1;
        atomicOr(tile_bits + hbit_word(j0 + r, lane), 1u << hbit_bit(j0 + r, lane));
    }
}

// acc[a][m] += v[s-m] over the block, a = base at position s of window array `w`, warp-uniform switch.
// REVERSE: target column MW-1-m and base 3-a (the strand fold of convRBM.py:367).
template <int MW, bool REVERSE, bool WITH_Q>
__device__ __forceinline__ void accumulate_stats(float (&acc)[4][MW], float (&accq)[4][MW],
                                                 const float (&v)[BLOCK_POS], const float (&q)[BLOCK_POS],
                                                 const uint32_t (&w)[3]) {
    constexpr int B = BLOCK_POS, S = B + MW - 1;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        int a = (int)window_bits<1, 3>(w, s);
        if (REVERSE) a = 3 - a;
#define CRBM_ACC_CASE(A)                                                          \
        case A: {                                                                 \
            _Pragma("unroll") for (int m = 0; m < MW; ++m) {                       \
                const int r = REVERSE ? s - (MW - 1 - m) : s - m;                 \
                if (r >= 0 && r < B) {                                            \
                    acc[A][m] += v[r];                                            \
                    if (WITH_Q) accq[A][m] += q[r];                               \
                }                                                                 \
            }                                                                     \
        } break;
        switch (a) {
            CRBM_ACC_CASE(0)
            CRBM_ACC_CASE(1)
            CRBM_ACC_CASE(2)
            CRBM_ACC_CASE(3)
        }
#undef CRBM_ACC_CASE
    }
}

// ----------------------------------------------------------------------------------------------------------
// Fused kernel: grid (ktiles, ctas_per_tile), 256 threads. Dynamic smem: table | wtile | mbarrier.
// ----------------------------------------------------------------------------------------------------------
template <int MW, int FLAGS>
__global__ void __launch_bounds__(256, 1) v2h_fused_kernel(const V2HParams p) {
    using Sh = V2HShape<MW>;
    using Ac = V2HAcc<MW, FLAGS>;
    constexpr int B = BLOCK_POS;
    constexpr bool STAT_P = FLAGS & F_STAT_P, STAT_Q = FLAGS & F_STAT_Q, SAMPLE = FLAGS & F_SAMPLE,
                   WANT_FE = FLAGS & F_FE, COUNT = FLAGS & F_COUNT;
    extern __shared__ __align__(128) unsigned char smem[];
    float2 *tab = reinterpret_cast<float2 *>(smem);
    float *wtile = reinterpret_cast<float *>(smem + Sh::table_bytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + Sh::table_bytes + Sh::wtile_bytes);

    const int kt = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int k = kt * KTILE + lane;
    const bool kvalid = k < p.K;
    build_table<MW>(tab, wtile, bar, p.W, p.b, p.K, p.M, kt);
    const float2 *tab_lane = tab + lane;

    float accp[4][MW], accq[4][MW];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int m = 0; m < MW; ++m) { accp[a][m] = 0.f; accq[a][m] = 0.f; }
    float sum_p = 0.f, sum_pr = 0.f, sum_q = 0.f;

    const int d2 = MW - p.M;
    const float fpool = (float)p.pooling;
    const int groups_per_seq = p.Lh / p.pooling;
    const int ktiles = gridDim.x;
    const PoolMasks mk0 = make_pool_masks(p.pooling);
    GroupUniforms gu_f, gu_r;
    if (SAMPLE || COUNT) {
        gu_f.key = gu_r.key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        gu_f.hi = gu_r.hi = p.ctr_hi + (p.step_dev ? (uint32_t)(*p.step_dev) : 0u);
        gu_f.tag = (COUNT ? RNG_EVAL : RNG_HIDDEN) | (p.iter << 8);
        gu_r.tag = RNG_HIDDEN_RC | (p.iter << 8);
    }

    const uint32_t *packed = p.packed + (p.batch_index ? (size_t)(*p.batch_index) * p.n_rows * p.row_words : 0);
    for (int item = blockIdx.y * nwarps + warp; item < p.n_items; item += gridDim.y * nwarps) {
        const int n = item / p.nblk, blk = item - n * p.nblk, j0 = blk * p.bv;
        const uint32_t *row = packed + (size_t)n * p.row_words;
        uint32_t wf[3];
        load_windows<3>(row, j0, wf);
        float x[B], xr[B];
        block_activities<MW>(tab_lane, wf, x, xr);
        PoolMasks mk = mk0;
        mk.valid = valid_mask(p.bv, j0, p.Lh);
        float cumn[B], pe[B];
        float fe = 0.f;
        // ---- forward strand -------------------------------------------------------------------------
        pool_block<WANT_FE>(x, cumn, pe, mk, fpool, fe);
        if (SAMPLE || COUNT) {
            gu_f.inj = p.u_fwd ? p.u_fwd + (size_t)(n * p.K + (kvalid ? k : 0)) * groups_per_seq : nullptr;
            gu_f.row = (uint32_t)(n * p.K + k);
            gu_f.cached_q = -1;
            uint32_t sb = sample_block(cumn, mk, gu_f, j0 / p.pooling);
            if (!kvalid) sb = 0u;
            if (COUNT) {
                const float tot = warp_sum((float)__popc(sb));
                if (lane == 0 && tot != 0.f) atomicAdd(p.count + n, tot);
            }
            if (SAMPLE) scatter_sample_bits(p.bits + ((size_t)(0 * p.n_rows + n) * ktiles + kt) * FANTASY_WORDS, sb, j0, lane);
        }
        if (STAT_P || STAT_Q) {
            float q[B];
#pragma unroll
            for (int r = 0; r < B; ++r) {
                q[r] = x[r] * pe[r];
                sum_p += x[r];
                sum_q += q[r];
            }
            if (STAT_P && STAT_Q) accumulate_stats<MW, false, true>(accp, accq, x, q, wf);
            else if (STAT_P) accumulate_stats<MW, false, false>(accp, accq, x, q, wf);
            else accumulate_stats<MW, false, false>(accq, accp, q, x, wf);
        }
        // ---- reverse-complement strand ----------------------------------------------------------------
        if (p.ds) {
            pool_block<WANT_FE>(xr, cumn, pe, mk, fpool, fe);
            if (SAMPLE) {
                gu_r.inj = p.u_rc ? p.u_rc + (size_t)(n * p.K + (kvalid ? k : 0)) * groups_per_seq : nullptr;
                gu_r.row = (uint32_t)(n * p.K + k);
                gu_r.cached_q = -1;
                uint32_t sb = sample_block(cumn, mk, gu_r, j0 / p.pooling);
                if (!kvalid) sb = 0u;
                scatter_sample_bits(p.bits + ((size_t)(1 * p.n_rows + n) * ktiles + kt) * FANTASY_WORDS, sb, j0, lane);
            }
            if (STAT_P) {
#pragma unroll
                for (int r = 0; r < B; ++r) sum_pr += xr[r];
                uint32_t wr[3];
                if (d2 == 0) { wr[0] = wf[0]; wr[1] = wf[1]; wr[2] = wf[2]; }
                else load_windows<3>(row, j0 - d2, wr);
                accumulate_stats<MW, true, false>(accp, accq, xr, xr, wr);
            }
        }
        if (WANT_FE) {
            if (p.fe_motif && kvalid) atomicAdd(p.fe_motif + (size_t)n * p.K + k, fe);
            if (p.fe_seq) {
                const float tot = warp_sum(kvalid ? fe : 0.f);
                if (lane == 0) atomicAdd(p.fe_seq + n, tot);
            }
        }
    }

    // ---- reduce the per-warp register accumulators over the CTA (tree through shared memory) -----------
    if (STAT_P || STAT_Q) {
        constexpr int NACC = Ac::NACC;
        __syncthreads();                                     // table no longer needed: reuse as scratch
        float *scr = reinterpret_cast<float *>(smem);        // [slot][NACC][32]
        auto put = [&](float *dst) {
            int o = 0;
            if (STAT_P) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int m = 0; m < MW; ++m) dst[(o++) * 32] = accp[a][m];
            }
            if (STAT_Q) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int m = 0; m < MW; ++m) dst[(o++) * 32] = accq[a][m];
            }
            dst[(o++) * 32] = sum_p;
            dst[(o++) * 32] = sum_pr;
            dst[(o++) * 32] = sum_q;
            dst[(o++) * 32] = 0.f;
        };
        auto add = [&](const float *src) {
            int o = 0;
            if (STAT_P) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int m = 0; m < MW; ++m) accp[a][m] += src[(o++) * 32];
            }
            if (STAT_Q) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int m = 0; m < MW; ++m) accq[a][m] += src[(o++) * 32];
            }
            sum_p += src[(o++) * 32];
            sum_pr += src[(o++) * 32];
            sum_q += src[(o++) * 32];
        };
        for (int half = nwarps >> 1; half >= 1; half >>= 1) {
            if (warp >= half && warp < 2 * half) put(scr + (size_t)(warp - half) * NACC * 32 + lane);
            __syncthreads();
            if (warp < half) add(scr + (size_t)warp * NACC * 32 + lane);
            __syncthreads();
        }
        if (warp == 0) put(p.partial + ((size_t)(kt * gridDim.y + blockIdx.y) * NACC) * 32 + lane);
    }
}

// ----------------------------------------------------------------------------------------------------------
// Dense-output kernel (scan / eager builders): out[n][k][j] = activity or pooled probability.
// Each warp computes a 32-motif x 16-position tile and writes it transposed through shared memory so that
// global stores are contiguous along j.
// ----------------------------------------------------------------------------------------------------------
struct DenseParams {
    const uint32_t *packed;
    int row_words, n_rows, L, Lh;
    const float *W, *b;
    int K, M, pooling, strand_mode, out_kind;
    int bv, nblk, n_items;
    float *out;
};

template <int MW>
__global__ void __launch_bounds__(256, 1) v2h_dense_kernel(const DenseParams p) {
    using Sh = V2HShape<MW>;
    constexpr int B = BLOCK_POS;
    extern __shared__ __align__(128) unsigned char smem[];
    float2 *tab = reinterpret_cast<float2 *>(smem);
    float *wtile = reinterpret_cast<float *>(smem + Sh::table_bytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + Sh::table_bytes + Sh::wtile_bytes);
    float *stage = reinterpret_cast<float *>(smem + Sh::table_bytes + Sh::wtile_bytes + 16);  // [warp][32][17]

    const int kt = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    build_table<MW>(tab, wtile, bar, p.W, p.b, p.K, p.M, kt);
    const float2 *tab_lane = tab + lane;
    float *st = stage + warp * (KTILE * (B + 1));
    const float fpool = (float)p.pooling;
    const int k0 = kt * KTILE;
    const PoolMasks mk0 = make_pool_masks(p.pooling);

    for (int item = blockIdx.y * nwarps + warp; item < p.n_items; item += gridDim.y * nwarps) {
        const int n = item / p.nblk, blk = item - n * p.nblk, j0 = blk * p.bv;
        const uint32_t *row = p.packed + (size_t)n * p.row_words;
        uint32_t wf[3];
        load_windows<3>(row, j0, wf);
        float x[B], xr[B];
        block_activities<MW>(tab_lane, wf, x, xr);
        if (p.strand_mode == CRBM_STRAND_RC) {
#pragma unroll
            for (int r = 0; r < B; ++r) x[r] = xr[r];
        } else if (p.strand_mode == CRBM_STRAND_SUM) {
#pragma unroll
            for (int r = 0; r < B; ++r) x[r] += xr[r];
        }
        PoolMasks mk = mk0;
        mk.valid = valid_mask(p.bv, j0, p.Lh);
        if (p.out_kind == CRBM_OUT_PROBABILITY) {
            float cumn[B], pe[B], fe = 0.f;
            pool_block<false>(x, cumn, pe, mk, fpool, fe);
        }
#pragma unroll
        for (int r = 0; r < B; ++r) st[lane * (B + 1) + r] = x[r];
        __syncwarp();
        // 32 motifs x 16 positions -> two motif rows per pass, 16 consecutive floats each
        const int rr = lane & 15, half = lane >> 4;
#pragma unroll
        for (int kk = 0; kk < KTILE; kk += 2) {
            const int kl = kk + half;
            if (k0 + kl < p.K && ((mk.valid >> rr) & 1u))
                p.out[((size_t)n * p.K + k0 + kl) * p.Lh + j0 + rr] = st[kl * (B + 1) + rr];
        }
        __syncwarp();
    }
}

template <int MW> constexpr size_t fused_smem_bytes() {
    return V2HShape<MW>::table_bytes + V2HShape<MW>::wtile_bytes + 16;
}
template <int MW> constexpr size_t dense_smem_bytes() {
    return fused_smem_bytes<MW>() + (size_t)8 * KTILE * (BLOCK_POS + 1) * sizeof(float);
}

// Host-side launchers, one translation unit per MW (v2h_inst_*.cu) so the heavy unrolled kernels build in parallel.
template <int MW> int launch_v2h_fused(int flags, const V2HParams &p, dim3 grid, cudaStream_t stream);
template <int MW> int launch_v2h_dense(const DenseParams &p, dim3 grid, cudaStream_t stream);

}  // namespace crbm
