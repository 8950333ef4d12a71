// One PCD/CD-k training step of the cRBM (reference: secomo/convRBM.py:373-438 `_updateWeightsOnMinibatch`),
// free energy / evaluation (:657-697, :466-472) and the dense visible->hidden entry point, built on the fused
// kernels of v2h.cuh plus the small kernels below:
//   h2v_bits_kernel     A5+A6: sparse hidden bit planes -> visible activity -> 4-way softmax -> sampled packed V
//   reduce_partials     per-CTA register accumulators -> the rank's `stats` vector
//   apply_update        A11 sparsity factor + A12 momentum update of W, b, c
#include "v2h_mma.cuh"

namespace crbm {

#define CRBM_FOR_EACH_MW(X) X(3) X(6) X(9) X(12) X(15) X(18) X(21) X(24) X(27) X(30) X(33)
#define CRBM_DECLARE(MW)                                                                             \
    template <> int launch_v2h_fused<MW>(int, const V2HParams &, dim3, cudaStream_t);                \
    template <> int launch_v2h_dense<MW>(const DenseParams &, dim3, cudaStream_t);
CRBM_FOR_EACH_MW(CRBM_DECLARE)
#undef CRBM_DECLARE

static int dispatch_fused(int MW, int flags, const V2HParams &p, dim3 grid, cudaStream_t s) {
    switch (MW) {
#define CRBM_CASE(MWV) case MWV: return launch_v2h_fused<MWV>(flags, p, grid, s);
        CRBM_FOR_EACH_MW(CRBM_CASE)
#undef CRBM_CASE
    }
    set_error("motif length bucket %d not built", MW);
    return CRBM_ERR_UNSUPPORTED;
}
#define CRBM_FOR_EACH_KS(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9)
#define CRBM_DECLARE_KS(KSV) template <> int launch_v2h_mma<KSV>(int, const V2HParams &, dim3, cudaStream_t);
CRBM_FOR_EACH_KS(CRBM_DECLARE_KS)
#undef CRBM_DECLARE_KS
static int dispatch_mma(int KS, int flags, const V2HParams &p, dim3 grid, cudaStream_t s) {
    switch (KS) {
#define CRBM_CASE(KSV) case KSV: return launch_v2h_mma<KSV>(flags, p, grid, s);
        CRBM_FOR_EACH_KS(CRBM_CASE)
#undef CRBM_CASE
    }
    set_error("MMA k-step bucket %d not built", KS);
    return CRBM_ERR_UNSUPPORTED;
}
static int dispatch_dense(int MW, const DenseParams &p, dim3 grid, cudaStream_t s) {
    switch (MW) {
#define CRBM_CASE(MWV) case MWV: return launch_v2h_dense<MWV>(p, grid, s);
        CRBM_FOR_EACH_MW(CRBM_CASE)
#undef CRBM_CASE
    }
    set_error("motif length bucket %d not built", MW);
    return CRBM_ERR_UNSUPPORTED;
}

static int check_shape(const char *who, int K, int M, int pooling, int Lh) {
    CRBM_CHECK_ARG(K > 0 && M > 0 && pooling > 0, "%s: K, M, pooling must be positive", who);
    CRBM_CHECK_ARG(Lh > 0, "%s: sequences shorter than the motif", who);
    if (M > CRBM_MAX_MOTIF_LEN) { set_error("%s: motif_length %d > %d is not covered by the sm_100a kernels", who, M, CRBM_MAX_MOTIF_LEN); return CRBM_ERR_UNSUPPORTED; }
    if (pooling > CRBM_MAX_POOLING) { set_error("%s: pooling %d > %d is not covered by the sm_100a kernels", who, pooling, CRBM_MAX_POOLING); return CRBM_ERR_UNSUPPORTED; }
    return CRBM_OK;
}

struct Tiling { int MW, ktiles, bv, nblk, n_items, gy; };
static Tiling make_tiling(int K, int M, int pooling, int Lh, int n_rows) {
    Tiling t;
    t.MW = 3 * ceil_div(M, 3);
    t.ktiles = ceil_div(K, KTILE);
    t.bv = (BLOCK_POS / pooling) * pooling;
    t.nblk = ceil_div(Lh, t.bv);
    t.n_items = n_rows * t.nblk;
    const int per_tile = num_sms() / t.ktiles > 0 ? num_sms() / t.ktiles : 1;
    t.gy = ceil_div(t.n_items, 8);
    if (t.gy > per_tile) t.gy = per_tile;
    if (t.gy < 1) t.gy = 1;
    return t;
}
static int max_gy(int K) { const int kt = ceil_div(K, KTILE); return num_sms() / kt > 0 ? num_sms() / kt : 1; }

// ---- A5 + A6 on bit planes ------------------------------------------------------------------------------------
// grid (particles, 2 halves of the visible sequence), 128 threads = one visible position each (Lv = 199+M <= 232).
//  1. the particle's bit planes (MMA-fragment order, common.cuh) are re-ordered in shared memory into one word of 32
//     motif bits per hidden position;
//  2. W is staged as float4 {A,C,G,T} per (k, m) when it fits, so a hit costs one LDS.128 and four adds;
//  3. every thread gathers the hits that reach its position, takes the 4-way softmax, samples one base and the
//     half-warps assemble the packed 2-bit words.
struct H2VParams {
    const uint32_t *bits;     // [2][B][kw][FANTASY_WORDS]
    int Bn, K, kw, M, ds, w4_in_smem;
    const float *W, *c;
    uint32_t *v_packed;       // rows of Lv bases
    int v_row_words;
    const float *u_v;         // injected uniforms [n*Lv + i] or null
    uint64_t seed;
    uint32_t ctr_hi;
    const unsigned long long *step_dev;
    float *base_counts;       // stats slot (4 floats) or null
};
constexpr int H2V_THREADS = 128, H2V_SPLIT = 112;      // positions [0,112) and [112, Lv): both multiples of 16 bases/word

__device__ __forceinline__ uint32_t gather_stride4(uint32_t x) {   // bits 0,4,..,28 -> bits 0..7
    x = (x | (x >> 3)) & 0x03030303u;
    x = (x | (x >> 6)) & 0x000F000Fu;
    return (x | (x >> 12)) & 0xFFu;
}

__global__ void __launch_bounds__(H2V_THREADS) h2v_bits_kernel(const H2VParams p) {
    extern __shared__ __align__(16) unsigned char h2v_smem[];
    const int M = p.M, Lv = FANTASY_LEN + M - 1;
    const int strands = p.ds ? 2 : 1;
    uint32_t *posw = reinterpret_cast<uint32_t *>(h2v_smem);                                    // [strand][kw][FANTASY_WORDS]
    float4 *w4 = reinterpret_cast<float4 *>(h2v_smem + (size_t)strands * p.kw * FANTASY_WORDS * 4);  // [K][M]
    const int n = blockIdx.x, half = blockIdx.y;
    const int i = half * H2V_SPLIT + threadIdx.x;
    const int ilo = half * H2V_SPLIT, ihi = half == 0 ? min(H2V_SPLIT, Lv) : Lv;               // visible range of this CTA
    const int jlo = max(0, ilo - (M - 1)), jhi = min(FANTASY_LEN, ihi);                         // hidden range it needs
    for (int e = threadIdx.x; e < strands * p.kw * (jhi - jlo); e += blockDim.x) {
        const int j = jlo + e % (jhi - jlo), rest = e / (jhi - jlo);
        const int t = rest % p.kw, s = rest / p.kw;
        const uint32_t *src = p.bits + (((size_t)s * p.Bn + n) * p.kw + t) * FANTASY_WORDS + (j >> 4) * 16;
        const int nt = (j >> 3) & 1, q = (j >> 1) & 3, ee = j & 1;
        uint32_t word = 0u;
#pragma unroll
        for (int mh = 0; mh < 4; ++mh)      // mh = 2*mt + h -> motifs 8*mh .. 8*mh+7 of the k-tile
            word |= gather_stride4((__ldg(src + (mh * 2 + nt) * 2 + ee) >> q) & 0x11111111u) << (8 * mh);
        posw[(s * p.kw + t) * FANTASY_WORDS + j] = word;
    }
    if (p.w4_in_smem) {
        for (int e = threadIdx.x; e < p.K * M; e += blockDim.x) {
            const int k = e / M, m = e - k * M;
            const float *wk = p.W + (size_t)k * 4 * M + m;
            w4[e] = make_float4(__ldg(wk), __ldg(wk + M), __ldg(wk + 2 * M), __ldg(wk + 3 * M));
        }
    }
    __syncthreads();
    int pick = 0;
    const bool active = i < ihi;
    if (active) {
        float v0 = __ldg(p.c), v1 = __ldg(p.c + 1), v2 = __ldg(p.c + 2), v3 = __ldg(p.c + 3);
        const int mlo = max(0, i - (FANTASY_LEN - 1)), mhi = min(M - 1, i);
        for (int m = mlo; m <= mhi; ++m) {
            const int j = i - m;
            for (int t = 0; t < p.kw; ++t) {
                uint32_t w = posw[t * FANTASY_WORDS + j];
                while (w) {
                    const int k = t * 32 + __ffs(w) - 1;
                    w &= w - 1;
                    float4 f;
                    if (p.w4_in_smem) f = w4[k * M + m];
                    else { const float *wk = p.W + (size_t)k * 4 * M + m; f = make_float4(__ldg(wk), __ldg(wk + M), __ldg(wk + 2 * M), __ldg(wk + 3 * M)); }
                    v0 += f.x; v1 += f.y; v2 += f.z; v3 += f.w;
                }
                if (p.ds) {
                    uint32_t w2 = posw[(p.kw + t) * FANTASY_WORDS + j];
                    while (w2) {
                        const int k = t * 32 + __ffs(w2) - 1;
                        w2 &= w2 - 1;
                        float4 f;                                                    // W[k, 3-a, M-1-m]
                        if (p.w4_in_smem) f = w4[k * M + (M - 1 - m)];
                        else { const float *wk = p.W + (size_t)k * 4 * M + (M - 1 - m); f = make_float4(__ldg(wk), __ldg(wk + M), __ldg(wk + 2 * M), __ldg(wk + 3 * M)); }
                        v0 += f.w; v1 += f.z; v2 += f.y; v3 += f.x;
                    }
                }
            }
        }
        const float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
        const float e0 = __expf(v0 - mx), e1 = __expf(v1 - mx), e2 = __expf(v2 - mx), e3 = __expf(v3 - mx);
        const float tot = e0 + e1 + e2 + e3;
        float u;
        if (p.u_v) u = p.u_v[(size_t)n * Lv + i];
        else {
            const uint32_t hi = p.ctr_hi + (p.step_dev ? (uint32_t)(*p.step_dev) * 64u : 0u);
            u = u01(philox4x32(make_uint4((uint32_t)(n * Lv + i), 0u, hi, RNG_VISIBLE),
                               make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32))).x);
        }
        const float ut = u * tot;
        pick = (e0 > ut) ? 0 : (e0 + e1 > ut) ? 1 : (e0 + e1 + e2 > ut) ? 2 : 3;
    }
    // assemble 16 bases per word inside each half warp
    const int lane = threadIdx.x & 31;
    uint32_t v = active ? ((uint32_t)pick << (2 * (lane & 15))) : 0u;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    if ((lane & 15) == 0 && i < ihi)
        p.v_packed[(size_t)n * p.v_row_words + FRONT_PAD_WORDS + (i >> 4)] = v;
    if (p.base_counts) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int cnt = __popc(__ballot_sync(0xffffffffu, active && pick == a));
            if (lane == 0 && cnt) atomicAdd(p.base_counts + a, (float)cnt);
        }
    }
}

// ---- base counts of a packed batch (A9 `_collectVStatistics` numerator) -----------------------------------------------
__global__ void count_bases_kernel(const uint32_t *__restrict__ packed, int row_words, int n_rows, int L,
                                   const int *__restrict__ batch_index, float *__restrict__ counts) {
    const int data_words = ceil_div(L, 16);
    const long long total = (long long)n_rows * data_words;
    const uint32_t *base = packed + (batch_index ? (size_t)(*batch_index) * n_rows * row_words : 0);
    int c[4] = {0, 0, 0, 0};
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(g / data_words), w = (int)(g - (long long)n * data_words);
        const uint32_t word = base[(size_t)n * row_words + FRONT_PAD_WORDS + w];
        const int nb = min(16, L - 16 * w);
        // count codes 1,2,3 with bit tricks; code 0 = nb - others
        const uint32_t lo = word & 0x55555555u, hi = (word >> 1) & 0x55555555u;
        const int c3 = __popc(lo & hi), c1 = __popc(lo & ~hi), c2 = __popc(hi & ~lo);
        c[1] += c1; c[2] += c2; c[3] += c3; c[0] += nb - c1 - c2 - c3;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const float t = warp_sum((float)c[a]);
        if ((threadIdx.x & 31) == 0 && t != 0.f) atomicAdd(counts + a, t);
    }
}

// ---- per-CTA partial accumulators -> stats ------------------------------------------------------------------------------
struct ReduceParams {
    const float *partial;   // [ktiles][gy][nacc][32]
    int ktiles, gy, nacc, np, nq, MW, M, K;
    float *vh_p, *vh_q;     // K*4*M destinations (np/nq accumulators), may be null when np/nq == 0
    float *h_sum, *h_fwd, *h_q;   // K each; h_sum += sum_p + sum_pr; may be null
};
__global__ void reduce_partials_kernel(const ReduceParams p) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = p.ktiles * p.nacc * 32;
    if (gid >= total) return;
    const int lane = gid & 31, o = (gid >> 5) % p.nacc, kt = gid / (32 * p.nacc);
    const int k = kt * KTILE + lane;
    if (k >= p.K) return;
    const float *src = p.partial + ((size_t)kt * p.gy * p.nacc + o) * 32 + lane;
    float v = 0.f;
    for (int c = 0; c < p.gy; ++c) v += src[(size_t)c * p.nacc * 32];
    if (o < p.np + p.nq) {
        const bool isq = o >= p.np;
        const int oo = isq ? o - p.np : o;
        const int a = oo / p.MW, m = oo - a * p.MW;
        if (m < p.M) (isq ? p.vh_q : p.vh_p)[((size_t)k * 4 + a) * p.M + m] = v;
    } else {
        const int w = o - p.np - p.nq;
        if (w == 0) { if (p.h_fwd) p.h_fwd[k] = v; if (p.h_sum) atomicAdd(p.h_sum + k, v); }
        else if (w == 1) { if (p.h_sum) atomicAdd(p.h_sum + k, v); }
        else if (w == 2) { if (p.h_q) p.h_q[k] = v; }
    }
}

// Same for the MMA kernels, whose accumulators are in fragment order (v2h_mma.cuh): register o of lane (g,q) holds
// motif 16mt + 8(r>>1) + g, column c = 8t + 2q + (r&1) = 4m + a.
__global__ void reduce_partials_mma_kernel(const ReduceParams p, int KS) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = p.ktiles * p.nacc * 32;
    if (gid >= total) return;
    const int lane = gid & 31, o = (gid >> 5) % p.nacc, kt = gid / (32 * p.nacc);
    const int g = lane >> 2, q = lane & 3;
    const float *src = p.partial + ((size_t)kt * p.gy * p.nacc + o) * 32 + lane;
    float v = 0.f;
    for (int c = 0; c < p.gy; ++c) v += src[(size_t)c * p.nacc * 32];
    if (o < p.np + p.nq) {
        const bool isq = o >= p.np;
        const int oo = isq ? o - p.np : o;
        const int r = oo & 3, t = (oo >> 2) % (2 * KS), mt = oo / (8 * KS);
        const int k = kt * KTILE + 16 * mt + 8 * (r >> 1) + g;
        const int col = 8 * t + 2 * q + (r & 1), m = col >> 2, a = col & 3;
        if (k < p.K && m < p.M) (isq ? p.vh_q : p.vh_p)[((size_t)k * 4 + a) * p.M + m] = v;
    } else if (q == 0) {
        const int w = o - p.np - p.nq, which = w >> 2, r4 = w & 3;
        const int k = kt * KTILE + 16 * (r4 >> 1) + 8 * (r4 & 1) + g;
        if (k < p.K) {
            if (which == 0) { if (p.h_fwd) p.h_fwd[k] = v; if (p.h_sum) atomicAdd(p.h_sum + k, v); }
            else if (which == 1) { if (p.h_sum) atomicAdd(p.h_sum + k, v); }
            else if (p.h_q) p.h_q[k] = v;
        }
    }
}

__global__ void write_counts_kernel(float *stats, crbm_stats_layout lay, float nhd, float nhm, float nvd, float nvm) {
    if (threadIdx.x == 0) {
        stats[lay.n_hidden_data] = nhd; stats[lay.n_hidden_model] = nhm;
        stats[lay.n_vis_data] = nvd; stats[lay.n_vis_model] = nvm;
    }
}

// ---- A11 + A12: sparsity factor and momentum update ------------------------------------------------------------------------
struct ApplyParams {
    float *W, *b, *c, *vW, *vb, *vc;
    const float *stats;
    crbm_stats_layout lay;
    int K, M, ds;
    float lr, mu, rho, lam;
    unsigned long long *step_dev;
    int *batch_index_dev;
    int n_batches;
};
__global__ void apply_update_kernel(const ApplyParams p) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nW = p.K * 4 * p.M;
    const float *st = p.stats;
    const float nhd = st[p.lay.n_hidden_data], nhm = st[p.lay.n_hidden_model];
    const float strands = p.ds ? 2.f : 1.f;
    if (gid < nW + p.K) {
        const int k = gid < nW ? gid / (4 * p.M) : gid - nW;
        const float pk = st[p.lay.h_fwd + k] / nhd;
        const float gk = -(p.rho / pk - (1.f - p.rho) / (1.f - pk)) / ((float)p.K * nhd);
        if (gid < nW) {
            const float gd = st[p.lay.vh_data + gid] / (strands * nhd);
            const float gm = st[p.lay.vh_model + gid] / (strands * nhm);
            const float reg = gk * st[p.lay.vh_q + gid];
            const float v = p.mu * p.vW[gid] + p.lr * (gd - gm - p.lam * reg);
            p.vW[gid] = v;
            p.W[gid] += v;
        } else {
            const float gd = st[p.lay.h_data + k] / (strands * nhd);
            const float gm = st[p.lay.h_model + k] / (strands * nhm);
            const float reg = gk * st[p.lay.h_q + k];
            const float v = p.mu * p.vb[k] + p.lr * (gd - gm - p.lam * reg);
            p.vb[k] = v;
            p.b[k] += v;
        }
    } else if (gid < nW + p.K + 4) {
        const int a = gid - nW - p.K;
        const float nvd = st[p.lay.n_vis_data], nvm = st[p.lay.n_vis_model];
        const float gd = (st[p.lay.v_data + a] + st[p.lay.v_data + 3 - a]) / nvd;      // f_a + f_{3-a}, convRBM.py:345
        const float gm = (st[p.lay.v_model + a] + st[p.lay.v_model + 3 - a]) / nvm;
        const float v = p.mu * p.vc[a] + p.lr * (gd - gm);
        p.vc[a] = v;
        p.c[a] += v;
    } else if (gid == nW + p.K + 4) {
        if (p.step_dev) *p.step_dev += 1ull;
        if (p.batch_index_dev) { const int nb = *p.batch_index_dev + 1; *p.batch_index_dev = nb >= p.n_batches ? 0 : nb; }
    }
}

// ---- free energy finish: add the visible-bias term, negate, normalise -----------------------------------------------------
__global__ void fe_finish_kernel(const uint32_t *__restrict__ packed, int row_words, int N, int L, int K,
                                 const float *__restrict__ c, float *fe_seq, float *fe_motif) {
    const int n = blockIdx.x;
    __shared__ float sh[32];
    float acc = 0.f;
    const float c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
    const uint32_t *row = packed + (size_t)n * row_words;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const int a = base_at(row, i);
        acc += a == 0 ? c0 : a == 1 ? c1 : a == 2 ? c2 : c3;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    float cterm = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) cterm += sh[w];
    if (fe_seq && threadIdx.x == 0) fe_seq[n] = -(fe_seq[n] + cterm) / (float)L;
    if (fe_motif)
        for (int k = threadIdx.x; k < K; k += blockDim.x) fe_motif[(size_t)n * K + k] = -fe_motif[(size_t)n * K + k] - cterm;
}

static size_t partial_floats(int K, int M) {
    const int MW = 3 * ceil_div(M, 3);
    const size_t nacc_cc = 8 * (size_t)MW + 4, nacc_mma = 32 * (size_t)ceil_div(M, 4) + 12;
    const size_t nacc = nacc_cc > nacc_mma ? nacc_cc : nacc_mma;
    return (size_t)ceil_div(K, KTILE) * max_gy(K) * nacc * 32;
}

static void fill_layout(int K, int M, crbm_stats_layout *l) {
    int o = 0;
    const int nW = K * 4 * M;
    l->vh_data = o; o += nW;
    l->vh_model = o; o += nW;
    l->vh_q = o; o += nW;
    l->h_data = o; o += K;
    l->h_model = o; o += K;
    l->h_q = o; o += K;
    l->h_fwd = o; o += K;
    l->v_data = o; o += 4;
    l->v_model = o; o += 4;
    l->n_hidden_data = o++; l->n_hidden_model = o++; l->n_vis_data = o++; l->n_vis_model = o++;
    l->total = o;
}

}  // namespace crbm

using namespace crbm;

extern "C" {

int crbm_train_stats_layout(int K, int M, crbm_stats_layout *out_host) {
    CRBM_CHECK_ARG(out_host && K > 0 && M > 0, "crbm_train_stats_layout: bad arguments");
    fill_layout(K, M, out_host);
    return CRBM_OK;
}
size_t crbm_train_stats_floats(int K, int M) {
    crbm_stats_layout l;
    fill_layout(K, M, &l);
    return (size_t)l.total;
}
size_t crbm_train_workspace_bytes(int K, int M) {
    return 3 * partial_floats(K, M) * sizeof(float) + 256;
}

int crbm_v2h_dense(const uint32_t *packed, int row_words, int N, int L, const float *W, const float *b, int K, int M,
                   int pooling, int strand_mode, int out_kind, float *out, void *stream) {
    CRBM_CHECK_ARG(packed && W && b && out && N > 0, "crbm_v2h_dense: bad arguments");
    const int Lh = L - M + 1;
    int rc = check_shape("crbm_v2h_dense", K, M, pooling, Lh);
    if (rc) return rc;
    CRBM_CHECK_ARG(strand_mode >= 0 && strand_mode <= 2 && (out_kind == 0 || out_kind == 1), "crbm_v2h_dense: bad mode");
    CRBM_CHECK_ARG(row_words >= packed_row_words(L), "crbm_v2h_dense: row_words too small");
    CRBM_CHECK_ARG(((uintptr_t)W & 15) == 0, "crbm_v2h_dense: W must be 16-byte aligned");
    int pool = pooling;
    if (out_kind == CRBM_OUT_PROBABILITY)
        CRBM_CHECK_ARG(Lh % pooling == 0, "crbm_v2h_dense: hidden length %d not divisible by pooling %d", Lh, pooling);
    else pool = 1;
    Tiling t = make_tiling(K, M, pool, Lh, N);
    DenseParams p{packed, row_words, N, L, Lh, W, b, K, M, pool, strand_mode, out_kind, t.bv, t.nblk, t.n_items, out};
    return dispatch_dense(t.MW, p, dim3(t.ktiles, t.gy), (cudaStream_t)stream);
}

int crbm_free_energy(const uint32_t *packed, int row_words, int N, int L, const float *W, const float *b, const float *c,
                     int K, int M, int pooling, int doublestranded, float *fe_seq, float *fe_motif, float *hsample_count,
                     const float *uniforms, uint64_t seed, uint64_t offset, void *stream) {
    CRBM_CHECK_ARG(packed && W && b && c && N > 0 && (fe_seq || fe_motif || hsample_count), "crbm_free_energy: bad arguments");
    const int Lh = L - M + 1;
    int rc = check_shape("crbm_free_energy", K, M, pooling, Lh);
    if (rc) return rc;
    CRBM_CHECK_ARG(Lh % pooling == 0, "crbm_free_energy: hidden length %d not divisible by pooling %d", Lh, pooling);
    CRBM_CHECK_ARG(((uintptr_t)W & 15) == 0, "crbm_free_energy: W must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (fe_seq) CRBM_CHECK_CUDA(cudaMemsetAsync(fe_seq, 0, sizeof(float) * N, s));
    if (fe_motif) CRBM_CHECK_CUDA(cudaMemsetAsync(fe_motif, 0, sizeof(float) * (size_t)N * K, s));
    if (hsample_count) CRBM_CHECK_CUDA(cudaMemsetAsync(hsample_count, 0, sizeof(float) * N, s));
    Tiling t = make_tiling(K, M, pooling, Lh, N);
    V2HParams p{};
    p.packed = packed; p.row_words = row_words; p.n_rows = N; p.L = L; p.Lh = Lh;
    p.W = W; p.b = b; p.K = K; p.M = M; p.pooling = pooling; p.ds = doublestranded ? 1 : 0;
    p.bv = t.bv; p.nblk = t.nblk; p.n_items = t.n_items;
    p.fe_motif = fe_motif; p.fe_seq = fe_seq; p.count = hsample_count;
    p.u_fwd = uniforms; p.seed = seed; p.ctr_hi = (uint32_t)offset;
    rc = dispatch_fused(t.MW, hsample_count ? (F_FE | F_COUNT) : F_FE, p, dim3(t.ktiles, t.gy), s);
    if (rc) return rc;
    fe_finish_kernel<<<N, 128, 0, s>>>(packed, row_words, N, L, K, c, fe_seq, fe_motif);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_train_accumulate(const crbm_train_args *a, void *stream) {
    CRBM_CHECK_ARG(a && a->W && a->b && a->c && a->data && a->fantasy_bits && a->v_model && a->stats && a->workspace,
                   "crbm_train_accumulate: null pointer in args");
    const int K = a->K, M = a->M, P = a->pooling, Lh = a->L - M + 1;
    int rc = check_shape("crbm_train_accumulate", K, M, P, Lh);
    if (rc) return rc;
    CRBM_CHECK_ARG(a->n_data >= 0 && a->batchsize > 0 && a->cd_k > 0, "crbm_train_accumulate: bad batch / cd_k");
    CRBM_CHECK_ARG(Lh % P == 0, "crbm_train_accumulate: hidden length %d not divisible by pooling %d (truncate as fit() does)", Lh, P);
    CRBM_CHECK_ARG(FANTASY_LEN % P == 0, "crbm_train_accumulate: pooling %d does not divide the fantasy length 200 "
                   "(the reference fails on this reshape too)", P);
    CRBM_CHECK_ARG(a->workspace_bytes >= crbm_train_workspace_bytes(K, M), "crbm_train_accumulate: workspace too small");
    CRBM_CHECK_ARG(a->data_row_words >= packed_row_words(a->L), "crbm_train_accumulate: data_row_words too small");
    const int Lv = FANTASY_LEN + M - 1;
    CRBM_CHECK_ARG(a->v_row_words >= packed_row_words(Lv), "crbm_train_accumulate: v_row_words too small");
    CRBM_CHECK_ARG(((uintptr_t)a->W & 15) == 0, "crbm_train_accumulate: W must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    crbm_stats_layout lay;
    fill_layout(K, M, &lay);
    float *st = a->stats;
    CRBM_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(float) * lay.total, s));
    const int kw = ceil_div(K, 32);
    const int ds = a->doublestranded ? 1 : 0;
    float *part0 = reinterpret_cast<float *>(a->workspace);
    float *part1 = part0 + partial_floats(K, M);
    float *part2 = part1 + partial_floats(K, M);
    // pooling == 1: both contractions on the tensor pipe (v2h_mma.cuh); otherwise the CUDA-core kernel (v2h.cuh)
    const bool mma = P == 1;
    const int KS = ceil_div(M, 4);
    const int nthreads = 256;

    auto run_fused = [&](int flags, const V2HParams &p, const Tiling &t) -> int {
        return mma ? dispatch_mma(KS, flags, p, dim3(t.ktiles, t.gy), s) : dispatch_fused(t.MW, flags, p, dim3(t.ktiles, t.gy), s);
    };
    auto run_reduce = [&](ReduceParams r, const Tiling &t, bool has_p, bool has_q) -> int {
        r.ktiles = t.ktiles; r.gy = t.gy; r.MW = t.MW; r.M = M; r.K = K;
        const int per = mma ? 16 * KS : 4 * t.MW;
        r.np = has_p ? per : 0; r.nq = has_q ? per : 0; r.nacc = r.np + r.nq + (mma ? 12 : 4);
        const int blocks = ceil_div(r.ktiles * r.nacc * 32, nthreads);
        if (mma) reduce_partials_mma_kernel<<<blocks, nthreads, 0, s>>>(r, KS);
        else reduce_partials_kernel<<<blocks, nthreads, 0, s>>>(r);
        CRBM_CHECK_LAUNCH();
        return CRBM_OK;
    };

    // ---- positive phase on the minibatch (convRBM.py:377-388) and the sparsity statistic (:440-451) ----------
    Tiling t = make_tiling(K, M, P, Lh, a->n_data);
    V2HParams p{};
    p.packed = a->data; p.row_words = a->data_row_words; p.n_rows = a->n_data; p.L = a->L; p.Lh = Lh;
    p.W = a->W; p.b = a->b; p.K = K; p.M = M; p.pooling = P; p.ds = ds;
    p.bv = t.bv; p.nblk = t.nblk; p.n_items = t.n_items;
    p.batch_index = a->batch_index_dev;
    const bool combined = mma ? KS <= 4 : t.MW <= 15;
    if (a->n_data == 0) {
        // a data-parallel rank whose shard of a short batch is empty contributes zeros
    } else if (combined) {
        p.partial = part0;
        if ((rc = run_fused(F_STAT_P | F_STAT_Q, p, t))) return rc;
        ReduceParams r{};
        r.partial = part0; r.vh_p = st + lay.vh_data; r.vh_q = st + lay.vh_q;
        r.h_sum = st + lay.h_data; r.h_fwd = st + lay.h_fwd; r.h_q = st + lay.h_q;
        if ((rc = run_reduce(r, t, true, true))) return rc;
    } else {
        p.partial = part0;
        if ((rc = run_fused(F_STAT_P, p, t))) return rc;
        ReduceParams r{};
        r.partial = part0; r.vh_p = st + lay.vh_data; r.h_sum = st + lay.h_data; r.h_fwd = st + lay.h_fwd;
        if ((rc = run_reduce(r, t, true, false))) return rc;
        p.partial = part2;
        if ((rc = run_fused(F_STAT_Q, p, t))) return rc;
        ReduceParams rq{};
        rq.partial = part2; rq.vh_q = st + lay.vh_q; rq.h_q = st + lay.h_q;
        if ((rc = run_reduce(rq, t, false, true))) return rc;
    }
    if (a->n_data > 0) {
        const int cb = ceil_div(a->n_data * ceil_div(a->L, 16), 256);
        count_bases_kernel<<<cb > 64 ? 64 : cb, 256, 0, s>>>(a->data, a->data_row_words, a->n_data, a->L,
                                                            a->batch_index_dev, st + lay.v_data);
        CRBM_CHECK_LAUNCH();
    }
    write_counts_kernel<<<1, 32, 0, s>>>(st, lay, (float)a->n_data * Lh, (float)a->batchsize * FANTASY_LEN,
                                         (float)a->n_data * a->L, (float)a->batchsize * Lv);
    CRBM_CHECK_LAUNCH();

    // ---- negative phase: cd_k Gibbs sweeps of the persistent chain (convRBM.py:391-413) -----------------------------
    Tiling tn = make_tiling(K, M, P, FANTASY_LEN, a->batchsize);
    const size_t n_u_v = (size_t)a->batchsize * Lv, n_u_h = (size_t)a->batchsize * K * (FANTASY_LEN / P);
    const size_t bits_bytes = (size_t)2 * a->batchsize * kw * FANTASY_WORDS * sizeof(uint32_t);
    const size_t w4_bytes = (size_t)K * M * sizeof(float4);
    const size_t posw_bytes = (size_t)(ds ? 2 : 1) * kw * FANTASY_WORDS * sizeof(uint32_t);
    const int w4_in_smem = posw_bytes + w4_bytes <= 160 * 1024;
    const size_t h2v_smem = posw_bytes + (w4_in_smem ? w4_bytes : 0);
    static size_t h2v_smem_configured = 0;
    if (h2v_smem > h2v_smem_configured) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(h2v_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h2v_smem));
        h2v_smem_configured = h2v_smem;
    }
    for (int it = 0; it < a->cd_k; ++it) {
        const bool last = it == a->cd_k - 1;
        H2VParams h{};
        h.bits = a->fantasy_bits; h.Bn = a->batchsize; h.K = K; h.kw = kw; h.M = M; h.ds = ds; h.w4_in_smem = w4_in_smem;
        h.W = a->W; h.c = a->c; h.v_packed = a->v_model; h.v_row_words = a->v_row_words;
        h.u_v = a->u_v ? a->u_v + it * n_u_v : nullptr;
        h.seed = a->seed; h.step_dev = a->step_dev;
        h.ctr_hi = a->step_dev ? (uint32_t)it : (uint32_t)(a->step * 64u + it);
        h.base_counts = last ? st + lay.v_model : nullptr;
        h2v_bits_kernel<<<dim3(a->batchsize, 2), H2V_THREADS, h2v_smem, s>>>(h);
        CRBM_CHECK_LAUNCH();

        V2HParams q{};
        q.packed = a->v_model; q.row_words = a->v_row_words; q.n_rows = a->batchsize; q.L = Lv; q.Lh = FANTASY_LEN;
        q.W = a->W; q.b = a->b; q.K = K; q.M = M; q.pooling = P; q.ds = ds;
        q.bv = tn.bv; q.nblk = tn.nblk; q.n_items = tn.n_items;
        q.partial = part1; q.bits = a->fantasy_bits;
        q.u_fwd = a->u_h ? a->u_h + it * n_u_h : nullptr;
        q.u_rc = a->u_hp ? a->u_hp + it * n_u_h : nullptr;
        q.seed = a->seed; q.step_dev = a->step_dev;
        q.ctr_hi = a->step_dev ? (uint32_t)it : (uint32_t)(a->step * 64u + it);
        if (!mma) CRBM_CHECK_CUDA(cudaMemsetAsync(a->fantasy_bits, 0, bits_bytes, s));   // CUDA-core sampler ORs its hits in
        if ((rc = run_fused(last ? (F_STAT_P | F_SAMPLE) : F_SAMPLE, q, tn))) return rc;
    }
    ReduceParams rn{};
    rn.partial = part1; rn.vh_p = st + lay.vh_model; rn.h_sum = st + lay.h_model;
    if ((rc = run_reduce(rn, tn, true, false))) return rc;
    return CRBM_OK;
}

int crbm_train_apply(const crbm_train_args *a, void *stream) {
    CRBM_CHECK_ARG(a && a->W && a->b && a->c && a->vW && a->vb && a->vc && a->stats, "crbm_train_apply: null pointer in args");
    ApplyParams p{};
    p.W = a->W; p.b = a->b; p.c = a->c; p.vW = a->vW; p.vb = a->vb; p.vc = a->vc; p.stats = a->stats;
    fill_layout(a->K, a->M, &p.lay);
    p.K = a->K; p.M = a->M; p.ds = a->doublestranded ? 1 : 0;
    p.lr = a->learning_rate; p.mu = a->momentum; p.rho = a->rho; p.lam = a->lambda_rate;
    p.step_dev = a->step_dev; p.batch_index_dev = a->batch_index_dev; p.n_batches = a->n_batches;
    const int n = a->K * 4 * a->M + a->K + 5;
    apply_update_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(p);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

}  // extern "C"
