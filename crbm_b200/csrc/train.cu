// One PCD/CD-k training step of the cRBM (reference: secomo/convRBM.py:373-438 `_updateWeightsOnMinibatch`),
// free energy / evaluation (:657-697, :466-472) and the dense visible->hidden entry point, built on the fused
// kernels of v2h.cuh plus the small kernels below:
//   h2v_bits_kernel     A5+A6: sparse hidden bit planes -> visible activity -> 4-way softmax -> sampled packed V
//   reduce_partials     per-CTA register accumulators -> the rank's `stats` vector
//   apply_update        A11 sparsity factor + A12 momentum update of W, b, c
#include <stdlib.h>
#include "v2h_mma.cuh"
#include "generic.cuh"

namespace crbm {

#define CRBM_FOR_EACH_MW(X) X(3) X(6) X(9) X(12) X(15) X(18) X(21) X(24) X(27) X(30) X(33)
#define CRBM_DECLARE(MW)                                                                             \
    template <> int launch_v2h_fused<MW>(int, const V2HParams &, dim3, cudaStream_t);                \
    template <> int launch_v2h_dense<MW>(const DenseParams &, dim3, cudaStream_t);
CRBM_FOR_EACH_MW(CRBM_DECLARE)
#undef CRBM_DECLARE

int launch_v2h_dense_tc(const DenseParams &, long long row_stride, cudaStream_t);          // scan_tc.cu: tcgen05 / TMEM scan kernel
static bool scan_uses_tcgen05() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("CRBM_SCAN_KERNEL"); v = (e && e[0] == 'm') ? 0 : 1; }   // "mma" -> mma.sync kernel
    return v == 1;
}

static bool pool_uses_mma() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("CRBM_POOL_KERNEL"); v = (e && e[0] == 'c') ? 0 : 1; }      // "cuda" -> CUDA-core kernel
    return v == 1;
}

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
#define CRBM_DECLARE_KS(KSV) template <> int launch_v2h_mma<KSV>(int, const V2HParams &, dim3, cudaStream_t); \
    template <> int launch_v2h_mma_dual<KSV>(const V2HDualParams &, dim3, cudaStream_t); \
    template <> int launch_v2h_dense_mma<KSV>(const DenseParams &, dim3, cudaStream_t); \
    template <> int launch_v2h_scan_reduce<KSV>(const ScanReduceParams &, cudaStream_t);
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
static int dispatch_mma_dual(int KS, const V2HDualParams &d, dim3 grid, cudaStream_t s) {
    switch (KS) {
#define CRBM_CASE(KSV) case KSV: return launch_v2h_mma_dual<KSV>(d, grid, s);
        CRBM_FOR_EACH_KS(CRBM_CASE)
#undef CRBM_CASE
    }
    set_error("MMA k-step bucket %d not built", KS);
    return CRBM_ERR_UNSUPPORTED;
}
static bool dual_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("CRBM_DUAL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
static int dispatch_dense_mma(int KS, const DenseParams &p, dim3 grid, cudaStream_t s) {
    switch (KS) {
#define CRBM_CASE(KSV) case KSV: return launch_v2h_dense_mma<KSV>(p, grid, s);
        CRBM_FOR_EACH_KS(CRBM_CASE)
#undef CRBM_CASE
    }
    set_error("MMA k-step bucket %d not built", KS);
    return CRBM_ERR_UNSUPPORTED;
}
static int dispatch_scan_reduce(int KS, const ScanReduceParams &p, cudaStream_t s) {
    switch (KS) {
#define CRBM_CASE(KSV) case KSV: return launch_v2h_scan_reduce<KSV>(p, s);
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
    return CRBM_OK;
}
// motif_length > CRBM_MAX_MOTIF_LEN or pooling > CRBM_MAX_POOLING: the shape-generic kernels (generic.cuh). CRBM_GENERIC=1 sends
// every shape there (tests compare them with the fixtures of the tuned kernels).
static bool shape_is_generic(int M, int P) {
    static int force = -1;
    if (force < 0) { const char *e = getenv("CRBM_GENERIC"); force = (e && e[0] == '1') ? 1 : 0; }
    return force == 1 || M > CRBM_MAX_MOTIF_LEN || P > CRBM_MAX_POOLING;
}
static int launch_gen_v2h(const GenParams &g, cudaStream_t s) {
    const size_t smem = (size_t)8 * g.M * sizeof(float);
    static size_t configured[64];
    int dev_id = 0;
    CRBM_CHECK_CUDA(cudaGetDevice(&dev_id));
    if (smem > 48 * 1024 && smem > configured[dev_id & 63]) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(gen_v2h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_id & 63] = smem;
    }
    const bool per_row = g.mode == GEN_REDUCE || g.mode == GEN_FE;
    const bool prob = g.mode != GEN_DENSE && g.mode != GEN_REDUCE ? true : g.out_kind == CRBM_OUT_PROBABILITY;
    const long long items = per_row ? g.n_rows : (long long)g.n_rows * (g.Lh / (prob ? g.P : 1));
    long long gy = (items + 255) / 256;
    const long long cap = (8ll * num_sms()) / g.K + 1;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    gen_v2h_kernel<<<dim3(g.K, (unsigned)gy), 256, smem, s>>>(g);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

struct Tiling { int MW, ktiles, bv, nblk, n_items, gy; };
static Tiling make_tiling(int K, int M, int pooling, int Lh, int n_rows) {
    Tiling t;
    t.MW = 3 * ceil_div(M, 3);
    t.ktiles = ceil_div(K, KTILE);
    t.bv = pooling <= BLOCK_POS ? (BLOCK_POS / pooling) * pooling : pooling;      // (pooling > 16 runs on the shape-generic kernels)
    t.nblk = ceil_div(Lh, t.bv);
    t.n_items = n_rows * t.nblk;
    const int per_tile = num_sms() / t.ktiles > 0 ? num_sms() / t.ktiles : 1;
    t.gy = ceil_div(t.n_items, 8);
    if (t.gy > per_tile) t.gy = per_tile;
    if (t.gy < 1) t.gy = 1;
    return t;
}
// tensor-pipe kernels: one CTA per (16-motif m-tile, slice of the position blocks), two CTAs per SM
static int mma_gy(int K, int n_items) {
    const int mtiles = ceil_div(K, 16);
    int gy = ceil_div(n_items, 8);
    const int cap = (2 * num_sms()) / mtiles > 0 ? (2 * num_sms()) / mtiles : 1;
    return gy > cap ? cap : (gy < 1 ? 1 : gy);
}
static int max_gy(int K) { const int kt = ceil_div(K, KTILE); return num_sms() / kt > 0 ? num_sms() / kt : 1; }

// ---- A5 + A6 on bit planes ------------------------------------------------------------------------------------
// Hit-parallel: one CTA per particle, one warp per 16 visible positions (= one packed output word). The warp scans the
// bit-plane words of the 2-3 hidden position blocks that can reach its positions (MMA-fragment order, common.cuh: a word
// holds 8 motifs x 4 positions), and for every set bit (k, j, strand) — warp-uniform — lane (p, ah) adds
// W[k, a, i-j] for its position i = 16w+p and its two bases a = 2ah, 2ah+1 into two registers (reverse strand:
// W[k, 3-a, M-1-(i-j)], convRBM.py:285). No shared memory, no atomics, no divergence beyond the range predicate.
// Then the two lanes of a position exchange their halves, take the 4-way softmax, draw the base and the warp ORs the
// 2-bit codes into the packed word.
struct H2VParams {
    const uint32_t *bits;     // [2][B][kw][FANTASY_WORDS]
    int Bn, K, kw, M, ds, w_in_smem, words_per_row, list_offset, split, split_log2;
    const float *W, *c;
    uint32_t *v_packed;       // rows of Lv bases
    int v_row_words;
    const float *u_v;         // injected uniforms [n*Lv + i] or null
    uint64_t seed;
    uint32_t ctr_hi, iter;    // training step and Gibbs sweep (see V2HParams)
    const unsigned long long *step_dev;
};

constexpr int H2V_LIST = 1024 + 4;     // a round holds at most 32 words x 32 bits hits (+ padding quad)

__global__ void __launch_bounds__(512) h2v_bits_kernel(const H2VParams p) {
    extern __shared__ __align__(128) unsigned char h2v_smem[];
    const int M = p.M, Lv = FANTASY_LEN + M - 1;
    // programmatic dependent launch: W (update kernel) and the bit planes (previous sweep) come from the predecessors
    pdl_wait();
    pdl_launch_dependents();
    // the motif bank is staged once per CTA by the TMA engine: L1 is cold at kernel start, and every hit reads two
    // short rows of W, so fetching them one miss at a time would serialise L2 round trips
    const float *Wsm = p.W;
    if (p.w_in_smem) {
        float *wbuf = reinterpret_cast<float *>(h2v_smem + 16);
        uint64_t *bar = reinterpret_cast<uint64_t *>(h2v_smem);
        const uint32_t bytes = (uint32_t)p.K * 16u * (uint32_t)M;
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, bytes);
            tma_load_1d(wbuf, p.W, bytes, bar);
        }
        mbar_wait(bar, 0);
        Wsm = wbuf;
    }
    // work item = (particle n, 16-position word w), shared by `split` adjacent warps of the CTA (a power of two): warp `part`
    // scans the k-tiles part, part + split, ...; the partial activities are combined through shared memory in a fixed order.
    // (One warp per item left 10-13 warps per SM walking 2 * kw * 2 rounds each: latency bound, 83 us per launch at K = 512.)
    const int wslot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int part = wslot & (p.split - 1);
    const int item_raw = wslot >> p.split_log2;
    const bool live = item_raw < p.Bn * p.words_per_row;
    const int item = live ? item_raw : 0;
    const int n = item / p.words_per_row, w = item - n * p.words_per_row, lane = threadIdx.x & 31;
    const int pidx = lane & 15, ah = lane >> 4;
    const int i = 16 * w + pidx;
    const int strands = p.ds ? 2 : 1;
    const int jb_lo = max(0, 16 * w - (M - 1)) >> 4, jb_hi = min(FANTASY_BLOCKS - 1, w);
    float acc0 = 0.f, acc1 = 0.f;
    // One round = the 2 x 16 bit-plane words of two position blocks of one (strand, k-tile), one word per lane.
    // The lanes first compact their set bits into a warp-private list of (motif row offset, hidden position) in
    // shared memory (warp prefix sum over popcounts), then the whole warp consumes the list four hits at a time:
    // the descriptor reads are broadcast LDS.128 and the eight weight loads of a quad of hits are independent.
    uint32_t *list = reinterpret_cast<uint32_t *>(h2v_smem + p.list_offset) + (threadIdx.x >> 5) * H2V_LIST;
    for (int s = 0; s < strands; ++s) {
        const float *Wbase = s == 0 ? Wsm + (size_t)(2 * ah) * M : Wsm + (size_t)(3 - 2 * ah) * M;
        const int dir = s == 0 ? M : -M;                     // second base row: +M (forward) / -M (reverse)
        for (int t = live ? part : p.kw; t < p.kw; t += p.split) {
            const uint32_t *tile = p.bits + (((size_t)s * p.Bn + n) * p.kw + t) * FANTASY_WORDS;
            for (int jb0 = jb_lo; jb0 <= jb_hi; jb0 += 2) {
                const int jbl = jb0 + (lane >> 4), f = lane & 15;
                uint32_t word = jbl <= jb_hi ? __ldg(tile + jbl * 16 + f) : 0u;
                if (__ballot_sync(0xffffffffu, word != 0u) == 0u) continue;
                const int cnt = __popc(word);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                int pos = incl - cnt;
                const int jbase = 16 * jbl + 8 * ((f >> 1) & 1) + (f & 1);
                const int krow = (32 * t + 8 * (f >> 2)) * 4 * M;
                while (word) {
                    const int b = __ffs(word) - 1;
                    word &= word - 1;
                    list[pos++] = ((uint32_t)(krow + (b >> 2) * 4 * M) << 8) | (uint32_t)(jbase + 2 * (b & 3));
                }
                if (lane < 4) list[total + lane] = 0xFFu;            // padding entries: position 255 never matches
                __syncwarp();
                for (int h0 = 0; h0 < total; h0 += 4) {
                    const uint4 d = *reinterpret_cast<const uint4 *>(list + h0);
                    const uint32_t dd[4] = {d.x, d.y, d.z, d.w};
                    float x0[4], x1[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int m = i - (int)(dd[u] & 0xFFu);
                        const bool ok = m >= 0 && m < M;
                        const float *wk = Wbase + (dd[u] >> 8) + (s == 0 ? m : M - 1 - m);
                        x0[u] = ok ? wk[0] : 0.f;
                        x1[u] = ok ? wk[dir] : 0.f;
                    }
                    acc0 += (x0[0] + x0[1]) + (x0[2] + x0[3]);
                    acc1 += (x1[0] + x1[1]) + (x1[2] + x1[3]);
                }
                __syncwarp();
            }
        }
    }
    if (p.split > 1) {
        // the list area of a warp is free now: it carries the warp's partial sums to the item's first warp
        float2 *mine = reinterpret_cast<float2 *>(list);
        mine[lane] = make_float2(acc0, acc1);
        __syncthreads();
        if (part == 0) {
            for (int i = 1; i < p.split; ++i) {
                const float2 o = reinterpret_cast<const float2 *>(list + (size_t)i * H2V_LIST)[lane];
                acc0 += o.x;
                acc1 += o.y;
            }
        }
    }
    if (part != 0 || !live) return;
    // lanes (p, 0) and (p, 1) hold bases {A,C} and {G,T} of position i
    const float o0 = __shfl_xor_sync(0xffffffffu, acc0, 16), o1 = __shfl_xor_sync(0xffffffffu, acc1, 16);
    const bool active = ah == 0 && i < Lv;
    int pick = 0;
    if (active) {
        const float v0 = acc0 + __ldg(p.c), v1 = acc1 + __ldg(p.c + 1), v2 = o0 + __ldg(p.c + 2), v3 = o1 + __ldg(p.c + 3);
        const float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
        const float e0 = __expf(v0 - mx), e1 = __expf(v1 - mx), e2 = __expf(v2 - mx), e3 = __expf(v3 - mx);
        const float tot = e0 + e1 + e2 + e3;
        float u;
        if (p.u_v) u = p.u_v[(size_t)n * Lv + i];
        else {
            const uint32_t hi = p.ctr_hi + (p.step_dev ? (uint32_t)(*p.step_dev) : 0u);
            u = u01(philox4x32(make_uint4((uint32_t)(n * Lv + i), 0u, hi, RNG_VISIBLE | (p.iter << 8)),
                               make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32))).x);
        }
        const float ut = u * tot;
        pick = (e0 > ut) ? 0 : (e0 + e1 > ut) ? 1 : (e0 + e1 + e2 > ut) ? 2 : 3;
    }
    uint32_t v = active ? ((uint32_t)pick << (2 * pidx)) : 0u;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) p.v_packed[(size_t)n * p.v_row_words + FRONT_PAD_WORDS + w] = v;
}

// ---- A0 inside the step: one-hot fp32 minibatch -> packed rows, read where it lies (device memory or mapped host memory) ----------
// The host-fed training call (crbm_train_step_hostfed) has no staging copy: this kernel reads the caller's (n,1,4,L) fp32 array
// itself. A warp owns 128 consecutive bases of one sequence: lane l loads the float4 of bases 4l..4l+3 of each of the four channel
// rows, i.e. every warp-wide load is 512 contiguous bytes (over PCIe: full-size read requests, all of a minibatch in flight at
// once), turns them into four 2-bit codes, and four lanes assemble one packed word by shuffles. Rows whose length or address is not
// a multiple of 4 floats take scalar loads. Pad words of the rows are never written (the caller zeroes the buffer once).
// Programmatic dependent launch: the kernel depends on nothing its predecessors in the step write (the packed rows were last read by
// the previous step, which had completed before the predecessor could trigger), so it starts while the chain sweeps run and only
// waits before it exits.
struct PackRowsParams {
    const float *src;
    int n, L, vec;
    uint32_t *dst;
    int row_words;
    int *bad;
    const unsigned int *ready_flag;      // staged variant: src is a device buffer a DMA copy fills; wait until *ready_flag == ready_value
    unsigned int ready_value;
};
__global__ void __launch_bounds__(256) pack_rows_kernel(const PackRowsParams p) {
    pdl_launch_dependents();
    if (p.ready_flag) {
        // the producer is the copy engine (crbm_train_step_staged): it makes progress independently of the SMs, so spinning here
        // cannot dead-lock; bounded all the same (~2 s), a missing copy poisons the validation counter instead of hanging the GPU
        __shared__ int arrived;
        if (threadIdx.x == 0) {
            unsigned int seen = 0u;
            long long spins = 0;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.ready_flag) : "memory");
                if (seen == p.ready_value) break;
                if (++spins > (1ll << 25)) { *p.bad = (int)0x80000000; break; }
                __nanosleep(64);
            }
            arrived = 1;
        }
        __syncthreads();
    }
    const int segs = ceil_div(p.L, 128);
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (wid < p.n * segs) {
        const int n = wid / segs, seg = wid - n * segs;
        const int i0 = 128 * seg + 4 * lane;
        const float *row = p.src + (size_t)n * 4 * p.L;
        float v[4][4];
        if (p.vec && i0 + 4 <= p.L) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float4 f = __ldcv(reinterpret_cast<const float4 *>(row + (size_t)a * p.L + i0));
                v[a][0] = f.x; v[a][1] = f.y; v[a][2] = f.z; v[a][3] = f.w;
            }
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[a][e] = (i0 + e < p.L) ? __ldcv(row + (size_t)a * p.L + i0 + e) : (a == 0 ? 1.f : 0.f);
        }
        uint32_t byte = 0u;
        int nbad = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a = v[0][e], c = v[1][e], g = v[2][e], t = v[3][e];
            const int ones = (a == 1.f) + (c == 1.f) + (g == 1.f) + (t == 1.f);
            const int zeros = (a == 0.f) + (c == 0.f) + (g == 0.f) + (t == 0.f);
            if (ones != 1 || zeros != 3) ++nbad;
            const uint32_t code = (c == 1.f) ? 1u : (g == 1.f) ? 2u : (t == 1.f) ? 3u : 0u;
            byte |= ((i0 + e < p.L) ? code : 0u) << (2 * e);
        }
        uint32_t word = byte << (8 * (lane & 3));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        if ((lane & 3) == 0 && i0 < p.L) p.dst[(size_t)n * p.row_words + FRONT_PAD_WORDS + 8 * seg + (lane >> 2)] = word;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nbad += __shfl_xor_sync(0xffffffffu, nbad, o);
        if (lane == 0 && nbad) atomicAdd(p.bad, nbad);
    }
    pdl_wait();
}
struct PackJob { const float *src; int *bad; const unsigned int *ready_flag; unsigned int ready_value; };
static int launch_pack_job(const PackJob &job, const crbm_train_args *a, cudaStream_t s) {
    PackRowsParams p{};
    p.src = job.src; p.n = a->n_data; p.L = a->L; p.dst = const_cast<uint32_t *>(a->data); p.row_words = a->data_row_words; p.bad = job.bad;
    p.vec = (a->L % 4 == 0 && ((uintptr_t)job.src & 15) == 0) ? 1 : 0;
    p.ready_flag = job.ready_flag; p.ready_value = job.ready_value;
    const int warps = a->n_data * ceil_div(a->L, 128);
    CRBM_CHECK_CUDA(launch_chain(pack_rows_kernel, dim3(ceil_div(warps, 8)), dim3(256), 0, s, p));
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

// ---- per-CTA partial accumulators -> stats ------------------------------------------------------------------------------
// One launch reduces every partial buffer of the step (positive phase P|Q, optional separate Q, negative phase).
// One CTA of 128 threads per (set, k-tile, accumulator slot): its four warps each add a quarter of the CTA partials
// (coalesced over the 32 lanes), combine through shared memory, and warp 0 scatters the 32 totals to `stats`.
struct ReduceSet {
    const float *partial;   // [ktiles][gy][nacc][32]
    int gy, nacc, np, nq;
    float *vh_p, *vh_q;     // K*4*M destinations (np/nq accumulators), may be null when np/nq == 0
    float *h_sum, *h_fwd, *h_q;   // K each; h_sum += sum_p + sum_pr; may be null
};
struct ReduceParams {
    ReduceSet set[3];
    int nsets, ktiles, MW, KS, mma, M, K;
    int block_start[4];     // first blockIdx.x of each set
    float *counts;          // stats + n_hidden_data (4 consecutive scalars) or null
    float c0, c1, c2, c3;
    // base counts (A9 `_collectVStatistics` numerators) are taken by two extra CTAs: [0] minibatch, [1] sampled V of the model
    const uint32_t *cnt_rows[2];
    int cnt_row_words[2], cnt_n[2], cnt_L[2];
    const int *batch_index;
    float *cnt_out[2];      // 4 floats each
};

// Counts of the four base codes over n rows of L packed bases; one CTA, result written (not accumulated).
__device__ void count_bases_block(const uint32_t *__restrict__ rows, int row_words, int n_rows, int L, float *out) {
    __shared__ int tot[4];
    if (threadIdx.x < 4) tot[threadIdx.x] = 0;
    __syncthreads();
    const int data_words = ceil_div(L, 16);
    int c[4] = {0, 0, 0, 0};
    for (int e = threadIdx.x; e < n_rows * data_words; e += blockDim.x) {
        const int n = e / data_words, w = e - n * data_words;
        const uint32_t word = __ldg(rows + (size_t)n * row_words + FRONT_PAD_WORDS + w);
        const int nb = min(16, L - 16 * w);
        const uint32_t lo = word & 0x55555555u, hi = (word >> 1) & 0x55555555u;
        const int c3 = __popc(lo & hi), c1 = __popc(lo & ~hi), c2 = __popc(hi & ~lo);
        c[1] += c1; c[2] += c2; c[3] += c3; c[0] += nb - c1 - c2 - c3;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        int v = c[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&tot[a], v);
    }
    __syncthreads();
    if (threadIdx.x < 4) out[threadIdx.x] = (float)tot[threadIdx.x];
}
__device__ __forceinline__ void reduce_partials_body(const ReduceParams &p);
// This warp's share (CTA partials warp, warp + 4, ...) of accumulator slot `slot` of k-tile `kt`, 32 lanes wide. Four independent loads
// in flight per pass (the partials are L2 resident; the loop is latency bound). The summation order is part of the contract: the plain
// reduction and the fused reduce + update kernel use this function and combine the four warps as (w0 + w1) + (w2 + w3).
__device__ __forceinline__ float warp_partial(const ReduceSet &S, int kt, int slot, int lane, int warp) {
    const float *src = S.partial + ((size_t)kt * S.gy * S.nacc + slot) * 32 + lane;
    const size_t cs = (size_t)S.nacc * 32;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int c = warp; c < S.gy; c += 16) {
        const float t0 = src[c * cs];
        const float t1 = c + 4 < S.gy ? src[(c + 4) * cs] : 0.f;
        const float t2 = c + 8 < S.gy ? src[(c + 8) * cs] : 0.f;
        const float t3 = c + 12 < S.gy ? src[(c + 12) * cs] : 0.f;
        a0 += t0; a1 += t1; a2 += t2; a3 += t3;
    }
    return (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(128) reduce_partials_kernel(const ReduceParams p) {
    pdl_wait();
    pdl_launch_dependents();
    reduce_partials_body(p);
}

__device__ __forceinline__ void reduce_partials_body(const ReduceParams &p) {
    __shared__ float sh[4][32], shr[4][32];
    if ((int)blockIdx.x >= p.block_start[p.nsets]) {
        const int w = blockIdx.x - p.block_start[p.nsets];
        if (p.cnt_n[w] > 0) {
            const uint32_t *rows = p.cnt_rows[w];
            if (w == 0 && p.batch_index) rows += (size_t)(*p.batch_index) * p.cnt_n[0] * p.cnt_row_words[0];
            count_bases_block(rows, p.cnt_row_words[w], p.cnt_n[w], p.cnt_L[w], p.cnt_out[w]);
        }
        return;
    }
    int si = 0;
    while (si + 1 < p.nsets && (int)blockIdx.x >= p.block_start[si + 1]) ++si;
    const ReduceSet &S = p.set[si];
    const int local = blockIdx.x - p.block_start[si];
    const int kt = local / S.nacc, o = local - kt * S.nacc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.counts) {
        p.counts[0] = p.c0; p.counts[1] = p.c1; p.counts[2] = p.c2; p.counts[3] = p.c3;
    }
    // h_sum = sum_p + sum_pr is written once, by the CTA of the sum_p slot, which also adds the partials of the matching
    // sum_pr slot (`pair` slots further); the CTA of the sum_pr slot has nothing left to do. No atomics, no zeroing.
    const int tail = o - S.np - S.nq, pair = p.mma ? 2 : 1;
    if (tail >= pair && tail < 2 * pair) return;
    const bool with_rc = tail >= 0 && tail < pair;
    float v = warp_partial(S, kt, o, lane, warp);
    float vr = with_rc ? warp_partial(S, kt, o + pair, lane, warp) : 0.f;
    sh[warp][lane] = v;
    shr[warp][lane] = vr;
    __syncthreads();
    if (warp != 0) return;
    v = (sh[0][lane] + sh[1][lane]) + (sh[2][lane] + sh[3][lane]);
    vr = (shr[0][lane] + shr[1][lane]) + (shr[2][lane] + shr[3][lane]);
    if (p.mma) {
        // fragment order (v2h_mma.cuh), kt = m-tile of 16 motifs: slot t*4+r of lane (g,q) = motif 16kt + 8(r>>1) + g,
        // column 8t + 2q + (r&1) = 4m + a; then sum_p[h], sum_pr[h], sum_q[h] on the q == 0 lanes
        const int g = lane >> 2, q = lane & 3;
        if (o < S.np + S.nq) {
            const bool isq = o >= S.np;
            const int oo = isq ? o - S.np : o;
            const int r = oo & 3, t = oo >> 2;
            const int k = kt * 16 + 8 * (r >> 1) + g;
            const int col = 8 * t + 2 * q + (r & 1), m = col >> 2, a = col & 3;
            if (k < p.K && m < p.M) (isq ? S.vh_q : S.vh_p)[((size_t)k * 4 + a) * p.M + m] = v;
        } else if (q == 0) {
            const int w = o - S.np - S.nq, which = w >> 1, h = w & 1;
            const int k = kt * 16 + 8 * h + g;
            if (k < p.K) {
                if (which == 0) { if (S.h_fwd) S.h_fwd[k] = v; if (S.h_sum) S.h_sum[k] = v + vr; }
                else if (which == 2 && S.h_q) S.h_q[k] = v;
            }
        }
    } else {
        const int k = kt * KTILE + lane;
        if (k >= p.K) return;
        if (o < S.np + S.nq) {
            const bool isq = o >= S.np;
            const int oo = isq ? o - S.np : o;
            const int a = oo / p.MW, m = oo - a * p.MW;
            if (m < p.M) (isq ? S.vh_q : S.vh_p)[((size_t)k * 4 + a) * p.M + m] = v;
        } else {
            const int w = o - S.np - S.nq;
            if (w == 0) { if (S.h_fwd) S.h_fwd[k] = v; if (S.h_sum) S.h_sum[k] = v + vr; }
            else if (w == 2) { if (S.h_q) S.h_q[k] = v; }
        }
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
    float *c_trace;                // optional per-step trace of c (crbm_train_args::c_trace; may be mapped host memory)
    int c_trace_slots;
    unsigned long long step;       // host-side step when step_dev is null
};
// Store the updated c[a] into the trace slot of this step (called by the thread that owns c[a], BEFORE the step cursor advances).
__device__ __forceinline__ void trace_c(const ApplyParams &p, int a, float v, unsigned long long step) {
    if (p.c_trace) p.c_trace[(size_t)(step % (unsigned long long)p.c_trace_slots) * 4 + a] = v;
}
// The update arithmetic, shared by the single-rank and the peer-memory kernel with explicit roundings (no FMA contraction that
// could differ between the two), so that a one-rank exchange reproduces the plain step bit for bit.
__device__ __forceinline__ float sparsity_factor(float h_fwd, float nhd, float rho, int K) {
    const float pk = __fdiv_rn(h_fwd, nhd);
    const float t = __fsub_rn(__fdiv_rn(rho, pk), __fdiv_rn(__fsub_rn(1.f, rho), __fsub_rn(1.f, pk)));
    return __fdiv_rn(-t, __fmul_rn((float)K, nhd));
}
// v <- mu v + lr (data / nd - model / nm - lam reg)
__device__ __forceinline__ float next_velocity(float vel, float data, float nd, float model, float nm, float reg, float mu, float lr, float lam) {
    const float g = __fsub_rn(__fsub_rn(__fdiv_rn(data, nd), __fdiv_rn(model, nm)), __fmul_rn(lam, reg));
    return __fadd_rn(__fmul_rn(mu, vel), __fmul_rn(lr, g));
}

__global__ void apply_update_kernel(const ApplyParams p) {
    pdl_wait();
    pdl_launch_dependents();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nW = p.K * 4 * p.M;
    const float *st = p.stats;
    const float nhd = st[p.lay.n_hidden_data], nhm = st[p.lay.n_hidden_model];
    const float strands = p.ds ? 2.f : 1.f;
    if (gid < nW + p.K) {
        const int k = gid < nW ? gid / (4 * p.M) : gid - nW;
        const float gk = sparsity_factor(st[p.lay.h_fwd + k], nhd, p.rho, p.K);
        const float nd = __fmul_rn(strands, nhd), nm = __fmul_rn(strands, nhm);
        if (gid < nW) {
            const float v = next_velocity(p.vW[gid], st[p.lay.vh_data + gid], nd, st[p.lay.vh_model + gid], nm,
                                          __fmul_rn(gk, st[p.lay.vh_q + gid]), p.mu, p.lr, p.lam);
            p.vW[gid] = v;
            p.W[gid] = __fadd_rn(p.W[gid], v);
        } else {
            const float v = next_velocity(p.vb[k], st[p.lay.h_data + k], nd, st[p.lay.h_model + k], nm,
                                          __fmul_rn(gk, st[p.lay.h_q + k]), p.mu, p.lr, p.lam);
            p.vb[k] = v;
            p.b[k] = __fadd_rn(p.b[k], v);
        }
    } else if (gid == nW + p.K) {
        // c (4 entries), its trace slot and the cursors of the next step: one thread, so that the trace sees the step before it advances
        const unsigned long long step = p.step_dev ? *p.step_dev : p.step;
        for (int a = 0; a < 4; ++a) {
            // f_a + f_{3-a}, convRBM.py:345
            const float v = next_velocity(p.vc[a], __fadd_rn(st[p.lay.v_data + a], st[p.lay.v_data + 3 - a]), st[p.lay.n_vis_data],
                                          __fadd_rn(st[p.lay.v_model + a], st[p.lay.v_model + 3 - a]), st[p.lay.n_vis_model], 0.f, p.mu, p.lr, 0.f);
            p.vc[a] = v;
            const float cn = __fadd_rn(p.c[a], v);
            p.c[a] = cn;
            trace_c(p, a, cn, step);
        }
        if (p.step_dev) *p.step_dev += 1ull;
        if (p.batch_index_dev) { const int nb = *p.batch_index_dev + 1; *p.batch_index_dev = nb >= p.n_batches ? 0 : nb; }
    }
}

// ---- reduction + update in one launch (single rank, tensor-pipe partial layout) ------------------------------------------------
// The update of W[k,a,m] needs three sums over the CTA partials (data P, data Q, model P) and the forward hidden sum of its motif:
// one CTA of 16 warps per (m-tile, accumulator slot) reduces those four rows (four warps per row, the summation order of
// reduce_partials_kernel) and applies the update to its 32 weights; two CTAs per m-tile do the 16 biases, one CTA the base counts, c
// and the step / minibatch cursors. (With 4 warps per CTA walking the four rows in turn the kernel was slower than the two launches.)
// Same arithmetic as reduce + apply (warp_partial, sparsity_factor, next_velocity): bit-identical parameters, one launch and one
// pass over the statistics less; the statistics vector is not written (crbm_train_accumulate still does).
struct FusedUpdateParams {
    ReduceParams red;     // the sets of the reduction: data P (| Q), optional separate Q, model
    ApplyParams ap;
    int sd, sq, sm;       // indices of the data-P set, of the set holding the Q slots, and of the model set
    int q_off;            // first Q slot inside set sq
    int per;              // accumulator slots per statistic (8 * KS)
};
__global__ void __launch_bounds__(512) reduce_update_kernel(const FusedUpdateParams p) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float sh[5][4][32];
    __shared__ float counts[8];
    const ReduceParams &R = p.red;
    const ApplyParams &ap = p.ap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = warp >> 2, sub = warp & 3;            // 16 warps: four per accumulator row, like the CTAs of reduce_partials_kernel
    const int g = lane >> 2, q = lane & 3;
    const ReduceSet &D = R.set[p.sd], &Q = R.set[p.sq], &Mo = R.set[p.sm];
    const float nhd = R.c0, nhm = R.c1;
    const float strands = ap.ds ? 2.f : 1.f;
    const float nd = __fmul_rn(strands, nhd), nm = __fmul_rn(strands, nhm);
    const int n_w = R.ktiles * p.per;
    auto total = [&](int r) { return (sh[r][0][lane] + sh[r][1][lane]) + (sh[r][2][lane] + sh[r][3][lane]); };
    if ((int)blockIdx.x < n_w) {
        const int kt = blockIdx.x / p.per, o = blockIdx.x - kt * p.per;
        const int r = o & 3, t = o >> 2, h = r >> 1;
        const ReduceSet &S = row == 0 ? D : row == 1 ? Q : row == 2 ? Mo : D;
        const int slot = row == 0 ? o : row == 1 ? p.q_off + o : row == 2 ? o : D.np + D.nq + h;   // row 3: sum_p[h], forward hidden sum
        sh[row][sub][lane] = warp_partial(S, kt, slot, lane, sub);
        __syncthreads();
        if (warp != 0) return;
        const float vd = total(0), vq = total(1), vm = total(2);
        const float hf = __shfl_sync(0xffffffffu, total(3), lane & 28);      // the sums sit on the q == 0 lanes
        const int k = kt * 16 + 8 * h + g;
        const int col = 8 * t + 2 * q + (r & 1), m = col >> 2, a = col & 3;
        if (k < ap.K && m < ap.M) {
            const int gid = (k * 4 + a) * ap.M + m;
            const float gk = sparsity_factor(hf, nhd, ap.rho, ap.K);
            const float v = next_velocity(ap.vW[gid], vd, nd, vm, nm, __fmul_rn(gk, vq), ap.mu, ap.lr, ap.lam);
            ap.vW[gid] = v;
            ap.W[gid] = __fadd_rn(ap.W[gid], v);
        }
    } else if ((int)blockIdx.x < n_w + 2 * R.ktiles) {
        // biases of motifs 16 kt + 8 h + g: rows sum_p, sum_pr (data), sum_p, sum_pr (model), sum_q
        const int kt = (blockIdx.x - n_w) >> 1, h = (blockIdx.x - n_w) & 1;
        const int bd = D.np + D.nq, bq = Q.np + Q.nq, bm = Mo.np + Mo.nq;
        if (row == 0) sh[0][sub][lane] = warp_partial(D, kt, bd + h, lane, sub);
        else if (row == 1) sh[1][sub][lane] = warp_partial(D, kt, bd + 2 + h, lane, sub);
        else if (row == 2) sh[2][sub][lane] = warp_partial(Mo, kt, bm + h, lane, sub);
        else sh[3][sub][lane] = warp_partial(Mo, kt, bm + 2 + h, lane, sub);
        if (row == 0) sh[4][sub][lane] = warp_partial(Q, kt, bq + 4 + h, lane, sub);
        __syncthreads();
        const int k = kt * 16 + 8 * h + g;
        if (warp != 0 || q != 0 || k >= ap.K) return;
        const float hf = total(0);
        const float hdata = hf + total(1), hmodel = total(2) + total(3), hq = total(4);
        const float gk = sparsity_factor(hf, nhd, ap.rho, ap.K);
        const float v = next_velocity(ap.vb[k], hdata, nd, hmodel, nm, __fmul_rn(gk, hq), ap.mu, ap.lr, ap.lam);
        ap.vb[k] = v;
        ap.b[k] = __fadd_rn(ap.b[k], v);
    } else {
        // base counts of the minibatch and of the sampled visible layer (A9), the update of c, the cursors of the next step
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const uint32_t *rows = R.cnt_rows[w];
            if (w == 0 && R.batch_index) rows += (size_t)(*R.batch_index) * R.cnt_n[0] * R.cnt_row_words[0];
            count_bases_block(rows, R.cnt_row_words[w], R.cnt_n[w], R.cnt_L[w], counts + 4 * w);
            __syncthreads();
        }
        if (threadIdx.x < 4) {
            const int a = threadIdx.x;
            const float v = next_velocity(ap.vc[a], __fadd_rn(counts[a], counts[3 - a]), R.c2, __fadd_rn(counts[4 + a], counts[7 - a]), R.c3,
                                          0.f, ap.mu, ap.lr, 0.f);
            ap.vc[a] = v;
            const float cn = __fadd_rn(ap.c[a], v);
            ap.c[a] = cn;
            trace_c(ap, a, cn, ap.step_dev ? *ap.step_dev : ap.step);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (ap.step_dev) *ap.step_dev += 1ull;
            if (ap.batch_index_dev) { const int nb = *ap.batch_index_dev + 1; *ap.batch_index_dev = nb >= ap.n_batches ? 0 : nb; }
        }
    }
}

// ---- reduction + all-reduce + update in ONE launch (data parallel, tensor-pipe partial layout) ---------------------------------
// reduce_update_kernel with the exchange between its two halves, at CTA granularity and without fences or separate flags: after a CTA
// has reduced its rows of the per-CTA partials, warp w PUSHES row w (32 values) into the receive buffers of every rank as 8-byte
// stores that carry the value and the step's sequence number together (the "LL" scheme of NCCL's low-latency protocol: an aligned
// 8-byte store is a single transaction, so a matching sequence number proves the value next to it has arrived), then polls its own
// receive buffer until the 32 x n_ranks pairs of its row carry the sequence number, adds them in rank order — bit-identical sums on
// every rank — and hands the sums to warp 0, which applies the update with the arithmetic of the single-rank kernel. One NVLink
// one-way latency per step instead of a flag round trip plus a pull round trip, no launch boundary between reduction and update, and the
// CTAs of fast m-tiles exchange while slow ones still reduce. Receive buffers are double buffered by step parity (a rank can be at most
// one exchange ahead of its slowest peer). CTA i of one rank only waits for CTA i of the others, which never waits for anything
// before it sends, and CTAs are dispatched in index order on every rank: no dead-lock. With one rank the result is bit-identical to
// reduce_update_kernel.
constexpr int LL_ROWS = 12;                                     // rows of 32 (value, sequence) pairs per (CTA, source rank)
struct LLView {
    int n_ranks, rank;
    unsigned long long *peer_ll[CRBM_MAX_PEERS];   // every rank's receive buffer of this step's parity: [n_cta][CRBM_MAX_PEERS][LL_ROWS][32]
    unsigned long long sequence;
    const unsigned long long *sequence_dev;
    int *error_flag;
    unsigned int *ticket;
};
struct FusedLLParams {
    FusedUpdateParams f;
    LLView ll;
};
__device__ __forceinline__ size_t ll_index(int cta, int src, int row, int lane) {
    return (((size_t)cta * CRBM_MAX_PEERS + src) * LL_ROWS + row) * 32 + lane;
}
// push this lane's value of `row` to every rank, then gather the row from every rank and add in rank order
__device__ __forceinline__ float ll_allreduce_row(const LLView &p, int cta, int row, int lane, float v, uint32_t seq, int &timed_out) {
    const unsigned long long packet = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(v);
#pragma unroll
    for (int r = 0; r < CRBM_MAX_PEERS; ++r)
        if (r < p.n_ranks) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p.peer_ll[r] + ll_index(cta, p.rank, row, lane)), "l"(packet) : "memory");
    const unsigned long long *mine = p.peer_ll[p.rank];
    unsigned long long x[CRBM_MAX_PEERS];
    long long spins = 0;
    for (;;) {
        bool all = true;
#pragma unroll
        for (int r = 0; r < CRBM_MAX_PEERS; ++r)
            if (r < p.n_ranks) {
                asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(x[r]) : "l"(mine + ll_index(cta, r, row, lane)) : "memory");
                all = all && ((uint32_t)(x[r] >> 32) == seq);
            }
        if (__all_sync(0xffffffffu, all)) break;
        if (++spins > (1ll << 24)) { timed_out = 1; break; }          // a peer is gone (seconds): do not hang the GPU
    }
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < CRBM_MAX_PEERS; ++r)
        if (r < p.n_ranks) sum += __uint_as_float((uint32_t)x[r]);
    return sum;
}
__global__ void __launch_bounds__(512) reduce_update_ll_kernel(const FusedLLParams P) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float sh[5][4][32];
    __shared__ float gs[LL_ROWS][32];
    __shared__ float counts[8];
    __shared__ int timed_out_s;
    const FusedUpdateParams &p = P.f;
    const ReduceParams &R = p.red;
    const ApplyParams &ap = p.ap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = warp >> 2, sub = warp & 3;
    const int g = lane >> 2, q = lane & 3;
    const ReduceSet &D = R.set[p.sd], &Q = R.set[p.sq], &Mo = R.set[p.sm];
    const float strands = ap.ds ? 2.f : 1.f;
    const int n_w = R.ktiles * p.per, cta = blockIdx.x;
    const uint32_t seq = (uint32_t)(P.ll.sequence_dev ? *P.ll.sequence_dev + 1ull : P.ll.sequence);
    if (threadIdx.x == 0) timed_out_s = 0;
    auto total = [&](int r) { return (sh[r][0][lane] + sh[r][1][lane]) + (sh[r][2][lane] + sh[r][3][lane]); };
    int timed_out = 0;
    if (cta < n_w) {
        const int kt = cta / p.per, o = cta - kt * p.per;
        const int r = o & 3, t = o >> 2, h = r >> 1;
        const ReduceSet &S = row == 0 ? D : row == 1 ? Q : row == 2 ? Mo : D;
        const int slot = row == 0 ? o : row == 1 ? p.q_off + o : row == 2 ? o : D.np + D.nq + h;
        sh[row][sub][lane] = warp_partial(S, kt, slot, lane, sub);
        __syncthreads();
        // rows 0-3: data P, data Q, model P, forward hidden sum; rows 4, 5: this rank's N*Lh and B*200
        if (warp < 6) {
            const float v = warp < 4 ? total(warp) : warp == 4 ? R.c0 : R.c1;
            gs[warp][lane] = ll_allreduce_row(P.ll, cta, warp, lane, v, seq, timed_out);
            if (timed_out) timed_out_s = 1;
        }
        __syncthreads();
        if (warp == 0) {
            const int k = kt * 16 + 8 * h + g;
            const int col = 8 * t + 2 * q + (r & 1), m = col >> 2, a = col & 3;
            if (k < ap.K && m < ap.M) {
                const int gid = (k * 4 + a) * ap.M + m;
                if (timed_out_s) { ap.W[gid] = __int_as_float(0x7fc00000); ap.vW[gid] = __int_as_float(0x7fc00000); }
                else {
                    const float vd = gs[0][lane], vq = gs[1][lane], vm = gs[2][lane], nhd = gs[4][lane], nhm = gs[5][lane];
                    const float hf = gs[3][lane & 28];
                    const float nd = __fmul_rn(strands, nhd), nm = __fmul_rn(strands, nhm);
                    const float gk = sparsity_factor(hf, nhd, ap.rho, ap.K);
                    const float v = next_velocity(ap.vW[gid], vd, nd, vm, nm, __fmul_rn(gk, vq), ap.mu, ap.lr, ap.lam);
                    ap.vW[gid] = v;
                    ap.W[gid] = __fadd_rn(ap.W[gid], v);
                }
            }
        }
    } else if (cta < n_w + 2 * R.ktiles) {
        const int kt = (cta - n_w) >> 1, h = (cta - n_w) & 1;
        const int bd = D.np + D.nq, bq = Q.np + Q.nq, bm = Mo.np + Mo.nq;
        if (row == 0) sh[0][sub][lane] = warp_partial(D, kt, bd + h, lane, sub);
        else if (row == 1) sh[1][sub][lane] = warp_partial(D, kt, bd + 2 + h, lane, sub);
        else if (row == 2) sh[2][sub][lane] = warp_partial(Mo, kt, bm + h, lane, sub);
        else sh[3][sub][lane] = warp_partial(Mo, kt, bm + 2 + h, lane, sub);
        if (row == 0) sh[4][sub][lane] = warp_partial(Q, kt, bq + 4 + h, lane, sub);
        __syncthreads();
        // rows 0-4: sum_p, sum_pr (data), sum_p, sum_pr (model), sum_q; rows 5, 6: N*Lh, B*200
        if (warp < 7) {
            const float v = warp < 5 ? total(warp) : warp == 5 ? R.c0 : R.c1;
            gs[warp][lane] = ll_allreduce_row(P.ll, cta, warp, lane, v, seq, timed_out);
            if (timed_out) timed_out_s = 1;
        }
        __syncthreads();
        const int k = kt * 16 + 8 * h + g;
        if (warp == 0 && q == 0 && k < ap.K) {
            if (timed_out_s) { ap.b[k] = __int_as_float(0x7fc00000); ap.vb[k] = __int_as_float(0x7fc00000); }
            else {
                const float hf = gs[0][lane], nhd = gs[5][lane], nhm = gs[6][lane];
                const float hdata = hf + gs[1][lane], hmodel = gs[2][lane] + gs[3][lane], hq = gs[4][lane];
                const float nd = __fmul_rn(strands, nhd), nm = __fmul_rn(strands, nhm);
                const float gk = sparsity_factor(hf, nhd, ap.rho, ap.K);
                const float v = next_velocity(ap.vb[k], hdata, nd, hmodel, nm, __fmul_rn(gk, hq), ap.mu, ap.lr, ap.lam);
                ap.vb[k] = v;
                ap.b[k] = __fadd_rn(ap.b[k], v);
            }
        }
    } else {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const uint32_t *rows = R.cnt_rows[w];
            if (w == 0 && R.batch_index) rows += (size_t)(*R.batch_index) * R.cnt_n[0] * R.cnt_row_words[0];
            count_bases_block(rows, R.cnt_row_words[w], R.cnt_n[w], R.cnt_L[w], counts + 4 * w);
            __syncthreads();
        }
        // row 0: the eight base counts in lanes 0-7; row 1: N*L and B*(199+M) in lanes 0, 1
        if (warp < 2) {
            const float v = warp == 0 ? (lane < 8 ? counts[lane] : 0.f) : (lane == 0 ? R.c2 : lane == 1 ? R.c3 : 0.f);
            gs[warp][lane] = ll_allreduce_row(P.ll, cta, warp, lane, v, seq, timed_out);
            if (timed_out) timed_out_s = 1;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            const int a = threadIdx.x;
            if (timed_out_s) { ap.c[a] = __int_as_float(0x7fc00000); ap.vc[a] = __int_as_float(0x7fc00000); }
            else {
                const float v = next_velocity(ap.vc[a], __fadd_rn(gs[0][a], gs[0][3 - a]), gs[1][0], __fadd_rn(gs[0][4 + a], gs[0][7 - a]), gs[1][1],
                                              0.f, ap.mu, ap.lr, 0.f);
                ap.vc[a] = v;
                const float cn = __fadd_rn(ap.c[a], v);
                ap.c[a] = cn;
                trace_c(ap, a, cn, ap.step_dev ? *ap.step_dev : ap.step);
            }
        }
    }
    // the step / minibatch cursors are advanced by the last CTA to finish: every CTA has read *sequence_dev by then
    __syncthreads();
    if (threadIdx.x == 0) {
        if (timed_out_s) *P.ll.error_flag = 1;
        const unsigned int t = atomicAdd(P.ll.ticket, 1u);
        if (t == gridDim.x - 1) {
            *P.ll.ticket = 0u;
            if (ap.step_dev) *ap.step_dev += 1ull;
            if (ap.batch_index_dev) { const int nb = *ap.batch_index_dev + 1; *ap.batch_index_dev = nb >= ap.n_batches ? 0 : nb; }
        }
    }
}

static void fill_ll_view(const crbm_train_args *a, const crbm_p2p_args *p2p, LLView *v) {
    v->n_ranks = p2p->n_ranks; v->rank = p2p->rank;
    for (int r = 0; r < p2p->n_ranks; ++r) v->peer_ll[r] = p2p->peer_ll[r] + (size_t)p2p->parity * p2p->ll_parity_stride;
    v->sequence = p2p->sequence; v->sequence_dev = a->step_dev; v->error_flag = p2p->error_flag; v->ticket = p2p->ticket;
}

// ---- A11 + A12 fused with a one-shot all-reduce over peer memory (data parallel) ------------------------------------------
// Every rank's statistics buffer is mapped into every other rank's address space (NVLink). A thread block first waits
// until all ranks have published this step's sequence number (their reduce kernels finished), then each thread adds its
// entries over the ranks in rank order — the sums are bit-identical on every rank, so the replicas stay identical — and
// applies the update. 18 420 floats x 8 ranks = 0.6 MB of remote reads per rank per step at C2.
struct ApplyP2PParams {
    ApplyParams ap;
    int n_ranks, parity;
    const float *peer_stats[CRBM_MAX_PEERS];      // every rank's statistics double buffer (peer-mapped)
    const unsigned long long *local_flags;        // this rank's flag array [n_ranks][2], written by the peers
    unsigned long long *peer_flags[CRBM_MAX_PEERS]; // every rank's flag array; entry [rank][parity] of each is ours to write
    int rank;
    size_t stats_stride;
    unsigned long long sequence;
    const unsigned long long *sequence_dev;
    int *error_flag;
    unsigned int *ticket;   // zero between launches (shared with the reduce kernel, which has finished by then)
};
// Sums of NI statistics entries over the ranks, in rank order. All NI * n_ranks loads are issued before the first add, so
// the thread pays one NVLink round trip, not one per term.
template <int NI>
__device__ __forceinline__ void p2p_sums(const ApplyP2PParams &p, const int (&idx)[NI], float (&out)[NI]) {
    float x[NI][CRBM_MAX_PEERS];
    const size_t off = (size_t)p.parity * p.stats_stride;
#pragma unroll
    for (int r = 0; r < CRBM_MAX_PEERS; ++r) {
        if (r < p.n_ranks) {
            const float *base = p.peer_stats[r] + off;
#pragma unroll
            for (int i = 0; i < NI; ++i) x[i][r] = __ldcv(base + idx[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < CRBM_MAX_PEERS; ++r) if (r < p.n_ranks) v += x[i][r];
        out[i] = v;
    }
}
__global__ void __launch_bounds__(256) apply_update_p2p_kernel(const ApplyP2PParams p) {
    pdl_wait();
    pdl_launch_dependents();
    const ApplyParams &ap = p.ap;
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    if (threadIdx.x < p.n_ranks) {
        const unsigned long long want = p.sequence_dev ? *p.sequence_dev + 1ull : p.sequence;
        // publish: this rank's statistics are complete (the reduction is the predecessor grid): one thread per peer stores the
        // step's sequence number into that peer's flag array (one-way NVLink store; every peer polls its own memory).
        // (The first version published from the last CTA of the reduction, elected by a ticket all 758 CTAs took after a
        // fence: measured with one rank as its own peer, the exchange cost 9.2 us per step.)
        if (blockIdx.x == 0)        // release at system scope: the statistics this thread can see are visible before the flag
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_flags[threadIdx.x] + 2 * p.rank + p.parity), "l"(want) : "memory");
        // acquire loads instead of a volatile spin + system fence (a MEMBAR.SYS per block was most of the exchange's cost)
        const unsigned long long *f = p.local_flags + 2 * threadIdx.x + p.parity;
        long long spins = 0;
        for (;;) {
            unsigned long long seen;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
            if (seen >= want) break;
            if (++spins > (1ll << 26)) { *p.error_flag = 1; timed_out = 1; break; }   // ~4 s: a peer is gone; do not hang the GPU
            __nanosleep(32);
        }
    }
    __syncthreads();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nW = ap.K * 4 * ap.M;
    const crbm_stats_layout &lay = ap.lay;
    const float strands = ap.ds ? 2.f : 1.f;
    if (timed_out) {
        // A peer's statistics never arrived: applying whatever its buffer holds would silently diverge this replica. The block
        // marks the parameters it owns invalid instead (NaN propagates into everything computed from them) and the host raises at
        // its next synchronisation point (error_flag; CRBM._p2p_check).
        const float nan = __int_as_float(0x7fc00000);
        if (gid < nW) { ap.W[gid] = nan; ap.vW[gid] = nan; }
        else if (gid < nW + ap.K) { ap.b[gid - nW] = nan; ap.vb[gid - nW] = nan; }
        else if (gid < nW + ap.K + 4) { ap.c[gid - nW - ap.K] = nan; ap.vc[gid - nW - ap.K] = nan; }
    } else if (gid < nW + ap.K) {
        const bool isW = gid < nW;
        const int k = isW ? gid / (4 * ap.M) : gid - nW;
        const int idx[6] = {lay.n_hidden_data, lay.n_hidden_model, lay.h_fwd + k,
                            isW ? lay.vh_data + gid : lay.h_data + k, isW ? lay.vh_model + gid : lay.h_model + k,
                            isW ? lay.vh_q + gid : lay.h_q + k};
        float s[6];
        p2p_sums<6>(p, idx, s);
        const float nhd = s[0], nhm = s[1];
        const float gk = sparsity_factor(s[2], nhd, ap.rho, ap.K);
        const float nd = __fmul_rn(strands, nhd), nm = __fmul_rn(strands, nhm);
        float *vel = isW ? ap.vW + gid : ap.vb + k, *par = isW ? ap.W + gid : ap.b + k;
        const float v = next_velocity(*vel, s[3], nd, s[4], nm, __fmul_rn(gk, s[5]), ap.mu, ap.lr, ap.lam);
        *vel = v;
        *par = __fadd_rn(*par, v);
    } else if (gid < nW + ap.K + 4) {
        const int a = gid - nW - ap.K;
        const int idx[6] = {lay.n_vis_data, lay.n_vis_model, lay.v_data + a, lay.v_data + 3 - a, lay.v_model + a, lay.v_model + 3 - a};
        float s[6];
        p2p_sums<6>(p, idx, s);
        const float v = next_velocity(ap.vc[a], __fadd_rn(s[2], s[3]), s[0], __fadd_rn(s[4], s[5]), s[1], 0.f, ap.mu, ap.lr, 0.f);
        ap.vc[a] = v;
        const float cn = __fadd_rn(ap.c[a], v);
        ap.c[a] = cn;
        trace_c(ap, a, cn, ap.step_dev ? *ap.step_dev : ap.step);      // (the cursor advances after every block has finished)
    }
    // the step / minibatch counters are advanced by the last block to finish: every block has read *sequence_dev by then
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(p.ticket, 1u);
        if (t == gridDim.x - 1) {
            *p.ticket = 0u;
            if (ap.step_dev) *ap.step_dev += 1ull;
            if (ap.batch_index_dev) { const int nb = *ap.batch_index_dev + 1; *ap.batch_index_dev = nb >= ap.n_batches ? 0 : nb; }
        }
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
    const size_t cc = (size_t)ceil_div(K, KTILE) * max_gy(K) * (8 * (size_t)MW + 4) * 32;
    const int mtiles = ceil_div(K, 16);
    const size_t mma = (size_t)mtiles * ((2 * num_sms()) / mtiles + 1) * (16 * (size_t)ceil_div(M, 4) + 6) * 32;
    return cc > mma ? cc : mma;
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

static void fill_apply_params(const crbm_train_args *a, ApplyParams *p) {
    p->W = a->W; p->b = a->b; p->c = a->c; p->vW = a->vW; p->vb = a->vb; p->vc = a->vc; p->stats = a->stats;
    fill_layout(a->K, a->M, &p->lay);
    p->K = a->K; p->M = a->M; p->ds = a->doublestranded ? 1 : 0;
    p->lr = a->learning_rate; p->mu = a->momentum; p->rho = a->rho; p->lam = a->lambda_rate;
    p->step_dev = a->step_dev; p->batch_index_dev = a->batch_index_dev; p->n_batches = a->n_batches;
    p->c_trace = a->c_trace_slots > 0 ? a->c_trace : nullptr; p->c_trace_slots = a->c_trace_slots; p->step = a->step;
}
static bool fused_update_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("CRBM_FUSED_UPDATE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
}  // namespace crbm

extern "C" int crbm_train_apply(const crbm_train_args *a, void *stream);
#include "train_tc_impl.cuh"

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
size_t crbm_train_workspace_bytes_for(int K, int M, int batchsize) {
    const size_t mma = 3 * partial_floats(K, M), tc = tt_workspace_floats(K, M, batchsize > 0 ? batchsize : 1);
    // shape-generic path: three probability buffers for a chunk of rows (all particles of the chain; minibatch rows in chunks)
    const size_t gen = M > CRBM_MAX_MOTIF_LEN ? (size_t)3 * K * (batchsize > 0 ? batchsize : 1) * 1024 : 0;
    const size_t m2 = mma > tc ? mma : tc;
    return (m2 > gen ? m2 : gen) * sizeof(float) + 256;
}
size_t crbm_train_workspace_bytes(int K, int M) { return crbm_train_workspace_bytes_for(K, M, 512); }

int crbm_v2h_dense_strided(const uint32_t *packed, int row_words, int N, int L, const float *W, const float *b, int K, int M,
                           int pooling, int strand_mode, int out_kind, float *out, long long out_row_stride, void *stream) {
    CRBM_CHECK_ARG(packed && W && b && out && N > 0, "crbm_v2h_dense: bad arguments");
    const int Lh = L - M + 1;
    CRBM_CHECK_ARG(out_row_stride >= Lh, "crbm_v2h_dense: output row stride %lld < hidden length %d", out_row_stride, Lh);
    int rc = check_shape("crbm_v2h_dense", K, M, pooling, Lh);
    if (rc) return rc;
    CRBM_CHECK_ARG(strand_mode >= 0 && strand_mode <= 2 && (out_kind == 0 || out_kind == 1), "crbm_v2h_dense: bad mode");
    CRBM_CHECK_ARG(row_words >= packed_row_words(L), "crbm_v2h_dense: row_words too small");
    CRBM_CHECK_ARG(((uintptr_t)W & 15) == 0, "crbm_v2h_dense: W must be 16-byte aligned");
    int pool = pooling;
    if (out_kind == CRBM_OUT_PROBABILITY)
        CRBM_CHECK_ARG(Lh % pooling == 0, "crbm_v2h_dense: hidden length %d not divisible by pooling %d", Lh, pooling);
    else pool = 1;
    if (shape_is_generic(M, pool)) {
        GenParams g{};
        g.packed = packed; g.row_words = row_words; g.n_rows = N; g.L = L; g.Lh = Lh; g.W = W; g.b = b; g.K = K; g.M = M; g.P = pool;
        g.mode = GEN_DENSE; g.strand_mode = strand_mode; g.out_kind = out_kind; g.out = out; g.out_row_stride = out_row_stride;
        return launch_gen_v2h(g, (cudaStream_t)stream);
    }
    if (scan_uses_tcgen05() && pool_uses_mma() && rowdense_wanted(K, M, pool, out_row_stride, out)) {
        RowDenseParams r{};
        r.packed = packed; r.row_words = row_words; r.n_rows = N; r.L = L; r.Lh = Lh; r.W = W; r.b = b; r.K = K; r.M = M;
        r.strand_mode = strand_mode; r.out_kind = out_kind; r.out = out; r.row_stride = out_row_stride;
        return launch_rowdense_tc(pool, r, (cudaStream_t)stream);
    }
    Tiling t = make_tiling(K, M, pool, Lh, N);
    DenseParams p{packed, row_words, N, L, Lh, W, b, K, M, pool, strand_mode, out_kind, t.bv, t.nblk, t.n_items, out};
    if ((pool == 1 || pool_uses_mma()) && scan_uses_tcgen05()) return launch_v2h_dense_tc(p, out_row_stride, (cudaStream_t)stream);
    if (out_row_stride != Lh) { set_error("crbm_v2h_dense: a padded output row stride needs pooling == 1"); return CRBM_ERR_UNSUPPORTED; }
    if (pool == 1) {
        // mma.sync tensor-pipe kernel (kept for comparison: CRBM_SCAN_KERNEL=mma): CTAs of one 16-motif m-tile, three per SM: size the grid to one full wave
        const int mtiles = ceil_div(K, 16);
        int gy = ceil_div(ceil_div(t.n_items, 2), 8);             // items of 32 positions, 8 warps per CTA
        const int cap = (3 * num_sms()) / mtiles > 0 ? (3 * num_sms()) / mtiles : 1;
        if (gy > cap) gy = cap;
        return dispatch_dense_mma(ceil_div(M, 4), p, dim3(mtiles, gy), (cudaStream_t)stream);
    }
    return dispatch_dense(t.MW, p, dim3(t.ktiles, t.gy), (cudaStream_t)stream);
}

const char *crbm_v2h_dense_kernel_name(int K, int M, int pooling, int out_kind, long long out_row_stride) {
    const int pool = out_kind == CRBM_OUT_PROBABILITY ? pooling : 1;
    if (K <= 0 || M <= 0 || pool <= 0) return "none";
    if (shape_is_generic(M, pool)) return "gen_v2h_kernel (shape-generic, CUDA cores)";
    if (scan_uses_tcgen05() && pool_uses_mma() && rowdense_wanted(K, M, pool, out_row_stride, nullptr))
        return ((32 / pool) * pool == 32) ? "v2h_rowdense_tc_kernel (tcgen05.mma, motif = TMEM lane, per-warp tiles stored by cp.async.bulk.tensor)"
                                          : "v2h_rowdense_tc_kernel (tcgen05.mma, motif = TMEM lane, rows stored by cp.async.bulk)";
    if ((pool == 1 || pool_uses_mma()) && scan_uses_tcgen05()) return "v2h_dense_tc_kernel (tcgen05.mma, position = TMEM lane, 128-byte row stores)";
    return pool == 1 ? "v2h_dense_mma_kernel (mma.sync)" : "v2h_dense_kernel (CUDA cores)";
}

int crbm_v2h_dense(const uint32_t *packed, int row_words, int N, int L, const float *W, const float *b, int K, int M,
                   int pooling, int strand_mode, int out_kind, float *out, void *stream) {
    return crbm_v2h_dense_strided(packed, row_words, N, L, W, b, K, M, pooling, strand_mode, out_kind, out, (long long)(L - M + 1), stream);
}

int crbm_scan_reduce(const uint32_t *packed, int row_words, int N, int L, const float *W, const float *b, int K, int M,
                     int pooling, int strand_mode, int out_kind, float *out_max, int *out_argmax, float *out_mean, void *stream) {
    CRBM_CHECK_ARG(packed && W && b && N > 0 && (out_max || out_argmax || out_mean), "crbm_scan_reduce: bad arguments");
    const int Lh = L - M + 1;
    int rc = check_shape("crbm_scan_reduce", K, M, pooling, Lh);
    if (rc) return rc;
    CRBM_CHECK_ARG(strand_mode >= 0 && strand_mode <= 2 && (out_kind == 0 || out_kind == 1), "crbm_scan_reduce: bad mode");
    CRBM_CHECK_ARG(row_words >= packed_row_words(L), "crbm_scan_reduce: row_words too small");
    CRBM_CHECK_ARG(((uintptr_t)W & 15) == 0, "crbm_scan_reduce: W must be 16-byte aligned");
    int pool = 1;
    if (out_kind == CRBM_OUT_PROBABILITY) {
        CRBM_CHECK_ARG(Lh % pooling == 0, "crbm_scan_reduce: hidden length %d not divisible by pooling %d", Lh, pooling);
        pool = pooling;
    }
    if (shape_is_generic(M, pool)) {
        GenParams g{};
        g.packed = packed; g.row_words = row_words; g.n_rows = N; g.L = L; g.Lh = Lh; g.W = W; g.b = b; g.K = K; g.M = M; g.P = pool;
        g.mode = GEN_REDUCE; g.strand_mode = strand_mode; g.out_kind = out_kind;
        g.out_max = out_max; g.out_argmax = out_argmax; g.out_mean = out_mean;
        return launch_gen_v2h(g, (cudaStream_t)stream);
    }
    if (rows_uses_tc(M, pool)) {
        // tcgen05 / TMEM kernel with the motif on the TMEM lane: the reductions over positions are register operations (scan_rt.cuh)
        RowsTcParams r{};
        r.packed = packed; r.row_words = row_words; r.n_rows = N; r.L = L; r.Lh = Lh;
        r.W = W; r.b = b; r.K = K; r.M = M; r.strand_mode = strand_mode; r.out_kind = out_kind;
        r.out_max = out_max; r.out_argmax = out_argmax; r.out_mean = out_mean;
        return launch_rows_tc(RT_REDUCE, pool, r, (cudaStream_t)stream);
    }
    Tiling t = make_tiling(K, M, pool, Lh, N);
    ScanReduceParams sp{};
    sp.d = DenseParams{packed, row_words, N, L, Lh, W, b, K, M, pool, strand_mode, out_kind, t.bv, t.nblk, t.n_items, nullptr};
    sp.out_max = out_max; sp.out_argmax = out_argmax; sp.out_mean = out_mean;
    return dispatch_scan_reduce(ceil_div(M, 4), sp, (cudaStream_t)stream);
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
    if (shape_is_generic(M, pooling)) {
        GenParams g{};
        g.packed = packed; g.row_words = row_words; g.n_rows = N; g.L = L; g.Lh = Lh; g.W = W; g.b = b; g.K = K; g.M = M; g.P = pooling;
        g.mode = GEN_FE; g.ds = doublestranded ? 1 : 0; g.fe_motif = fe_motif; g.fe_seq = fe_seq; g.count = hsample_count;
        g.u_fwd = uniforms; g.seed = seed; g.ctr_hi = (uint32_t)offset;
        rc = launch_gen_v2h(g, s);
        if (rc) return rc;
        fe_finish_kernel<<<N, 128, 0, s>>>(packed, row_words, N, L, K, c, fe_seq, fe_motif);
        CRBM_CHECK_LAUNCH();
        return CRBM_OK;
    }
    if (rows_uses_tc(M, pooling)) {
        RowsTcParams r{};
        r.packed = packed; r.row_words = row_words; r.n_rows = N; r.L = L; r.Lh = Lh;
        r.W = W; r.b = b; r.K = K; r.M = M; r.ds = doublestranded ? 1 : 0; r.out_kind = CRBM_OUT_PROBABILITY;
        r.fe_motif = fe_motif; r.fe_seq = fe_seq; r.count = hsample_count; r.uniforms = uniforms;
        r.seed = seed; r.ctr_hi = (uint32_t)offset;
        rc = launch_rows_tc(RT_FE, pooling, r, s);
        if (rc) return rc;
        fe_finish_kernel<<<N, 128, 0, s>>>(packed, row_words, N, L, K, c, fe_seq, fe_motif);
        CRBM_CHECK_LAUNCH();
        return CRBM_OK;
    }
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

// One training step as a chain of launches on ONE stream, in the order
//     h2v -> v2h(chain, sample) [x cd_k]  ->  v2h(minibatch: P, Q statistics)  ->  reduce  ->  update
// every kernel launched as a programmatic dependent of the one before it (common.cuh: launch_chain): the next grid's CTAs move in
// while the previous grid drains, run their prologue, and block in griddepcontrol.wait. The positive phase is independent of the chain
// sweeps and only waits before it exits (V2HParams::pdl_late), so it also overlaps their tail. (An earlier version forked the
// positive phase to a side stream; the two v2h grids each fill the register files, so they never ran side by side: measured
// `accumulate` - `accumulate without positive phase` = the positive kernel's stand-alone time.)
// `phases` selects which part is launched; the partial-sum buffers in the workspace are fixed, so the parts may be issued by
// separate calls (e.g. the chain sweeps while the next minibatch is still being copied in).
// ---- shape-generic training (generic.cuh): probabilities of a chunk of rows -> statistics, chunk by chunk -------------------------
struct GenCountParams {
    const uint32_t *rows[2];
    int row_words[2], n[2], L[2];
    const int *batch_index;
    float *out[2], *scalars;
    float c0, c1, c2, c3;
};
__global__ void __launch_bounds__(256) gen_counts_kernel(const GenCountParams p) {
    const int w = blockIdx.x;
    const uint32_t *rows = p.rows[w];
    if (w == 0 && p.batch_index) rows += (size_t)(*p.batch_index) * p.n[0] * p.row_words[0];
    if (p.n[w] > 0) count_bases_block(rows, p.row_words[w], p.n[w], p.L[w], p.out[w]);
    else if (threadIdx.x < 4) p.out[w][threadIdx.x] = 0.f;
    if (w == 0 && threadIdx.x == 0) { p.scalars[0] = p.c0; p.scalars[1] = p.c1; p.scalars[2] = p.c2; p.scalars[3] = p.c3; }
}
// v2h over `n_rows` rows (the chain: sampling into the bit planes; the minibatch: P and Q), in chunks that fit the workspace, and —
// when `want_stats` — the statistics of every chunk added into vh_p / vh_q / h_* in chunk order.
static int gen_train_rows(const crbm_train_args *a, const uint32_t *rows, int row_words, int n_rows, int L, int Lh, const int *batch_index,
                          int batch_rows, int sample, int iter, bool want_stats, bool want_q, float *vh_p, float *vh_q, float *h_sum,
                          float *h_fwd, float *h_q, cudaStream_t s) {
    const int K = a->K, M = a->M;
    const size_t ws_floats = (a->workspace_bytes - 256) / sizeof(float);
    const size_t per_row = (size_t)K * Lh;
    const size_t max_rows = ws_floats / (3 * per_row);
    if (max_rows < 1) { set_error("shape-generic training: workspace too small for one row of %d x %d probabilities", K, Lh); return CRBM_ERR_ARG; }
    float *pf = reinterpret_cast<float *>(a->workspace);
    float *pr = pf + max_rows * per_row, *q = pr + max_rows * per_row;
    const size_t n_u = (size_t)n_rows * K * (Lh / a->pooling);
    int rc;
    for (int r0 = 0, chunk = 0; r0 < n_rows; r0 += (int)max_rows, ++chunk) {
        const int nr = (int)(n_rows - r0 < (long long)max_rows ? n_rows - r0 : max_rows);
        GenParams g{};
        g.packed = rows; g.row_words = row_words; g.n_rows = nr; g.L = L; g.Lh = Lh; g.row0 = r0;
        g.batch_index = batch_index; g.batch_rows = batch_rows;
        g.W = a->W; g.b = a->b; g.K = K; g.M = M; g.P = a->pooling; g.ds = a->doublestranded ? 1 : 0; g.mode = GEN_TRAIN;
        g.pf = pf; g.pr = g.ds ? pr : nullptr; g.q = want_q ? q : nullptr;
        g.sample = sample; g.bits = a->fantasy_bits; g.chain_rows = a->batchsize; g.ktiles = ceil_div(K, 32);
        g.u_fwd = (sample && a->u_h) ? a->u_h + iter * n_u : nullptr;
        g.u_rc = (sample && a->u_hp) ? a->u_hp + iter * n_u : nullptr;
        g.seed = a->seed; g.step_dev = a->step_dev; g.ctr_hi = a->step_dev ? 0u : (uint32_t)a->step; g.iter = (uint32_t)iter;
        if ((rc = launch_gen_v2h(g, s))) return rc;
        if (!want_stats) continue;
        GenStatsParams t{};
        t.packed = rows; t.row_words = row_words; t.n_rows = nr; t.Lh = Lh; t.row0 = r0; t.batch_index = batch_index; t.batch_rows = batch_rows;
        t.K = K; t.M = M; t.accumulate = chunk > 0 ? 1 : 0;
        t.pf = pf; t.pr = g.pr; t.q = g.q; t.vh_p = vh_p; t.vh_q = vh_q; t.h_sum = h_sum; t.h_fwd = h_fwd; t.h_q = h_q;
        const int warps = K * 4 * M + K;
        gen_stats_kernel<<<ceil_div(warps, 8), 256, 0, s>>>(t);
        CRBM_CHECK_LAUNCH();
    }
    return CRBM_OK;
}

static int train_phases_impl(const crbm_train_args *a, int phases, void *stream, const crbm_p2p_args *p2p = nullptr, const PackJob *pack = nullptr) {
    CRBM_CHECK_ARG(a && a->W && a->b && a->c && a->data && a->fantasy_bits && a->v_model && a->stats && a->workspace,
                   "crbm_train_accumulate: null pointer in args");
    const int K = a->K, M = a->M, P = a->pooling, Lh = a->L - M + 1;
    int rc = check_shape("crbm_train_accumulate", K, M, P, Lh);
    if (rc) return rc;
    CRBM_CHECK_ARG(a->n_data >= 0 && a->batchsize > 0 && a->cd_k > 0 && a->cd_k < (1 << 24), "crbm_train_accumulate: bad batch / cd_k");
    CRBM_CHECK_ARG(Lh % P == 0, "crbm_train_accumulate: hidden length %d not divisible by pooling %d (truncate as fit() does)", Lh, P);
    CRBM_CHECK_ARG(FANTASY_LEN % P == 0, "crbm_train_accumulate: pooling %d does not divide the fantasy length 200 "
                   "(the reference fails on this reshape too)", P);
    CRBM_CHECK_ARG(a->workspace_bytes >= crbm_train_workspace_bytes_for(K, M, a->batchsize), "crbm_train_accumulate: workspace too small (crbm_train_workspace_bytes_for)");
    CRBM_CHECK_ARG(a->data_row_words >= packed_row_words(a->L), "crbm_train_accumulate: data_row_words too small");
    const int Lv = FANTASY_LEN + M - 1;
    CRBM_CHECK_ARG(a->v_row_words >= packed_row_words(Lv), "crbm_train_accumulate: v_row_words too small");
    CRBM_CHECK_ARG(((uintptr_t)a->W & 15) == 0, "crbm_train_accumulate: W must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const bool gen = shape_is_generic(M, P);                                                    // generic.cuh
    if (!gen && train_uses_tc(K, M, P)) return train_phases_tc(a, phases, s, p2p, pack);     // tcgen05 / TMEM path (train_tc.cuh)
    crbm_stats_layout lay;
    fill_layout(K, M, &lay);
    float *st = a->stats;
    const int kw = ceil_div(K, 32);
    const int ds = a->doublestranded ? 1 : 0;
    float *part0 = reinterpret_cast<float *>(a->workspace);
    float *part1 = part0 + partial_floats(K, M);
    float *part2 = part1 + partial_floats(K, M);
    // both contractions on the tensor pipe (v2h_mma.cuh); CRBM_POOL_KERNEL=cuda keeps pooling > 1 on the CUDA-core kernel (v2h.cuh)
    const bool mma = P == 1 || pool_uses_mma();
    const int KS = ceil_div(M, 4);
    auto run_fused = [&](int flags, const V2HParams &p, const Tiling &t) -> int {
        return mma ? dispatch_mma(KS, flags, p, dim3(ceil_div(K, 16), mma_gy(K, t.n_items)), s)
                   : dispatch_fused(t.MW, flags, p, dim3(t.ktiles, t.gy), s);
    };
    ReduceParams red{};
    red.ktiles = mma ? ceil_div(K, 16) : ceil_div(K, KTILE);        // tiles of the partial buffers: m-tiles / k-tiles
    red.MW = 3 * ceil_div(M, 3); red.KS = KS; red.mma = mma ? 1 : 0; red.M = M; red.K = K;
    // two-role launch (v2h_mma_dual_kernel): positive phase and the last chain sweep in one grid
    Tiling tn = make_tiling(K, M, P, FANTASY_LEN, a->batchsize);
    Tiling t = make_tiling(K, M, P, Lh, a->n_data);
    const bool combined = mma ? KS <= 4 : t.MW <= 15;
    const bool dual = !gen && mma && P == 1 && combined && a->n_data > 0 && dual_enabled() &&
                      (phases & CRBM_PHASE_NEGATIVE) && (phases & CRBM_PHASE_POSITIVE);
    int gy_pos = mma_gy(K, t.n_items), gy_neg = mma_gy(K, tn.n_items);
    if (dual) {
        // ONE wave of CTAs shared by the two roles (two per SM in total). The chain role starts later (it waits for h2v) and both
        // roles are bound by the latency of a warp's item chain, so the split follows the item counts with the chain role weighted
        // up. Measured at C2 (42 CTAs per m-tile; us per step): two full waves 42+42: 42.6 (= two launches), 21+21: 46.1,
        // 28+14: 61.6, 18+24: 42.2, 15+27: 38.9, 14+28: 39.7.
        const int G = (2 * num_sms()) / ceil_div(K, 16) > 2 ? (2 * num_sms()) / ceil_div(K, 16) : 2;
        const double wp = (double)t.n_items, wn = 1.6 * (double)tn.n_items;
        int x = (int)(G * wp / (wp + wn) + 0.5);
        x = x < G / 4 ? G / 4 : x > (3 * G) / 4 ? (3 * G) / 4 : x;
        x = x < 1 ? 1 : x;
        if (x < gy_pos) gy_pos = x;
        if (G - x < gy_neg) gy_neg = G - x > 0 ? G - x : 1;
        if (const char *e = getenv("CRBM_DUAL_GY_POS")) { const int v = atoi(e); if (v > 0 && v <= mma_gy(K, t.n_items)) gy_pos = v; }
        if (const char *e = getenv("CRBM_DUAL_GY_NEG")) { const int v = atoi(e); if (v > 0 && v <= mma_gy(K, tn.n_items)) gy_neg = v; }
    }
    auto add_set = [&](const float *partial, const Tiling &t, bool has_p, bool has_q, int gy_mma) -> ReduceSet & {
        ReduceSet &S = red.set[red.nsets];
        const int per = mma ? 8 * KS : 4 * t.MW;
        S.partial = partial; S.gy = mma ? gy_mma : t.gy;
        S.np = has_p ? per : 0; S.nq = has_q ? per : 0; S.nacc = S.np + S.nq + (mma ? 6 : 4);
        red.block_start[red.nsets + 1] = red.block_start[red.nsets] + red.ktiles * S.nacc;
        ++red.nsets;
        return S;
    };

    // ---- host-fed step: the minibatch is packed FIRST (a programmatic dependent of the previous step's update, whose trigger means
    // that the previous positive phase has read the packed rows). Every kernel of the chain triggers its dependents only after its
    // own wait, so the packing grid has completed before the first chain kernel behind it triggers, i.e. before any CTA of the
    // positive phase exists: that phase keeps depending on nothing in flight (pdl_late) and overlaps h2v as in the resident step.
    // (Between h2v and the positive phase, where it was, the packing kernel and its flag wait put the positive phase behind h2v:
    // 43.8 us per host-fed step at C2 against 38.6 resident.)
    bool packed_first = false;
    if (pack && (phases & CRBM_PHASE_POSITIVE) && (phases & CRBM_PHASE_NEGATIVE) && a->n_data > 0 && !gen) {
        if ((rc = launch_pack_job(*pack, a, s))) return rc;
        packed_first = true;
    }
    // ---- negative phase: cd_k Gibbs sweeps of the persistent chain (convRBM.py:391-413) -----------------------------
    V2HParams q_last{};                  // the last chain sweep's v2h launch, deferred into the two-role launch
    if (phases & CRBM_PHASE_NEGATIVE) {
        const size_t n_u_v = (size_t)a->batchsize * Lv, n_u_h = (size_t)a->batchsize * K * (FANTASY_LEN / P);
        const size_t bits_bytes = (size_t)2 * a->batchsize * kw * FANTASY_WORDS * sizeof(uint32_t);
        const size_t w_bytes = (size_t)K * 16 * M;
        const int h2v_items = a->batchsize * ceil_div(Lv, 16);
        // warps per item: at most 4, at least two k-tiles per warp, at most ~64 warps per SM
        // (measured: C2, kw = 4: 1 / 2 / 4 warps -> 47.5 / 46.9 / 49.0 us per step; C4, kw = 16: 1 / 4 warps -> 1515 / 1135 us)
        int split_log2 = 0;
        while (split_log2 < 2 && (4 << split_log2) <= kw && (long long)h2v_items * (2 << split_log2) <= 64ll * num_sms()) ++split_log2;
        if (const char *e = getenv("CRBM_H2V_SPLIT")) split_log2 = atoi(e) <= 1 ? 0 : atoi(e) == 2 ? 1 : 2;
        while ((1 << split_log2) > kw) --split_log2;
        const int h2v_split = 1 << split_log2;
        const int h2v_slots = h2v_items * h2v_split;
        int h2v_nw = ceil_div(h2v_slots, num_sms());
        h2v_nw = h2v_nw < 4 ? 4 : h2v_nw > 16 ? 16 : h2v_nw;
        h2v_nw = ceil_div(h2v_nw, h2v_split) * h2v_split;                // the warps of an item sit in one CTA
        const size_t list_bytes = (size_t)h2v_nw * H2V_LIST * sizeof(uint32_t);
        // W staged in shared memory by the TMA engine only on request (CRBM_H2V_WSMEM=1): measured at C2 the staging latency at the
        // head of the kernel costs more (31.3 us for the chain part) than the L1 misses of the first hits it avoids (30.5 us)
        static const int wsmem_allowed = [] { const char *e = getenv("CRBM_H2V_WSMEM"); return (e && e[0] == '1') ? 1 : 0; }();
        const int w_in_smem = wsmem_allowed && w_bytes + 16 + list_bytes <= 200 * 1024;
        const size_t list_offset = w_in_smem ? w_bytes + 16 : 16;
        const size_t h2v_smem = list_offset + list_bytes;
        static size_t h2v_smem_configured[64];               // per device (the attribute is per device)
        int dev_id = 0;
        CRBM_CHECK_CUDA(cudaGetDevice(&dev_id));
        size_t &configured = h2v_smem_configured[dev_id & 63];
        if (configured < 48 * 1024) configured = 48 * 1024;
        if (h2v_smem > configured) {
            CRBM_CHECK_CUDA(cudaFuncSetAttribute(h2v_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h2v_smem));
            configured = h2v_smem;
        }
        for (int it = 0; it < a->cd_k; ++it) {
            const bool last = it == a->cd_k - 1;
            H2VParams h{};
            h.bits = a->fantasy_bits; h.Bn = a->batchsize; h.K = K; h.kw = kw; h.M = M; h.ds = ds; h.w_in_smem = w_in_smem; h.words_per_row = ceil_div(Lv, 16); h.list_offset = (int)list_offset;
            h.split = h2v_split; h.split_log2 = split_log2;
            h.W = a->W; h.c = a->c; h.v_packed = a->v_model; h.v_row_words = a->v_row_words;
            h.u_v = a->u_v ? a->u_v + it * n_u_v : nullptr;
            h.seed = a->seed; h.step_dev = a->step_dev;
            h.ctr_hi = a->step_dev ? 0u : (uint32_t)a->step; h.iter = (uint32_t)it;
            CRBM_CHECK_CUDA(launch_chain(h2v_bits_kernel, dim3(ceil_div(h2v_slots, h2v_nw)), dim3(32 * h2v_nw), h2v_smem, s, h));
            CRBM_CHECK_LAUNCH();

            V2HParams q{};
            q.packed = a->v_model; q.row_words = a->v_row_words; q.n_rows = a->batchsize; q.L = Lv; q.Lh = FANTASY_LEN;
            q.W = a->W; q.b = a->b; q.K = K; q.M = M; q.pooling = P; q.ds = ds;
            q.bv = tn.bv; q.nblk = tn.nblk; q.n_items = tn.n_items;
            q.partial = part1; q.bits = a->fantasy_bits;
            q.u_fwd = a->u_h ? a->u_h + it * n_u_h : nullptr;
            q.u_rc = a->u_hp ? a->u_hp + it * n_u_h : nullptr;
            q.seed = a->seed; q.step_dev = a->step_dev;
            q.ctr_hi = a->step_dev ? 0u : (uint32_t)a->step; q.iter = (uint32_t)it;
            if (P != 1 || gen) CRBM_CHECK_CUDA(cudaMemsetAsync(a->fantasy_bits, 0, bits_bytes, s));  // the pooled / generic samplers OR their hits in
            if (gen) {
                if ((rc = gen_train_rows(a, a->v_model, a->v_row_words, a->batchsize, Lv, FANTASY_LEN, nullptr, 0, /*sample*/ 1, it,
                                         last, false, st + lay.vh_model, nullptr, st + lay.h_model, nullptr, nullptr, s))) return rc;
                continue;
            }
            if (last && dual) q_last = q;
            else if ((rc = run_fused(last ? (F_STAT_P | F_SAMPLE) : F_SAMPLE, q, tn))) return rc;
        }
    }

    // ---- positive phase on the minibatch (convRBM.py:377-388) and the sparsity statistic (:440-451) ----------
    V2HParams p{};
    p.packed = a->data; p.row_words = a->data_row_words; p.n_rows = a->n_data; p.L = a->L; p.Lh = Lh;
    p.W = a->W; p.b = a->b; p.K = K; p.M = M; p.pooling = P; p.ds = ds;
    p.bv = t.bv; p.nblk = t.nblk; p.n_items = t.n_items;
    p.batch_index = a->batch_index_dev;
    p.pdl_late = 1;
    const bool pos = (phases & CRBM_PHASE_POSITIVE) != 0;
    if (pack && pos && a->n_data > 0 && !packed_first) {
        // host-fed step without a negative phase in this call: the minibatch is packed here; the positive phase then reads what its
        // predecessor wrote and waits for it like every other kernel of the chain
        if ((rc = launch_pack_job(*pack, a, s))) return rc;
        p.pdl_late = 0;
    }
    if (a->n_data == 0) {
        // a data-parallel rank whose shard of a short batch is empty contributes zeros (its statistics are not written below)
        if (gen) {
            // (the generic path writes its statistics directly: only the minibatch's share is zeroed, the model's was written above)
            if (phases & CRBM_PHASE_POSITIVE) {
                const size_t nW = (size_t)K * 4 * M;
                CRBM_CHECK_CUDA(cudaMemsetAsync(st + lay.vh_data, 0, sizeof(float) * nW, s));
                CRBM_CHECK_CUDA(cudaMemsetAsync(st + lay.vh_q, 0, sizeof(float) * nW, s));
                CRBM_CHECK_CUDA(cudaMemsetAsync(st + lay.h_data, 0, sizeof(float) * K, s));
                CRBM_CHECK_CUDA(cudaMemsetAsync(st + lay.h_q, 0, sizeof(float) * K, s));
                CRBM_CHECK_CUDA(cudaMemsetAsync(st + lay.h_fwd, 0, sizeof(float) * K, s));
            }
        } else if (phases & CRBM_PHASE_REDUCE) CRBM_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(float) * lay.total, s));
    } else if (gen) {
        if (pos && (rc = gen_train_rows(a, a->data, a->data_row_words, a->n_data, a->L, Lh, a->batch_index_dev, a->n_data, 0, 0, true, true,
                                        st + lay.vh_data, st + lay.vh_q, st + lay.h_data, st + lay.h_fwd, st + lay.h_q, s))) return rc;
    } else if (combined) {
        p.partial = part0;
        if (dual) {
            V2HDualParams d{};
            d.pos = p; d.neg = q_last; d.gy_pos = gy_pos;
            if ((rc = dispatch_mma_dual(KS, d, dim3(ceil_div(K, 16), gy_pos + gy_neg), s))) return rc;
        } else if (pos && (rc = run_fused(F_STAT_P | F_STAT_Q, p, t))) return rc;
        ReduceSet &S = add_set(part0, t, true, true, gy_pos);
        S.vh_p = st + lay.vh_data; S.vh_q = st + lay.vh_q;
        S.h_sum = st + lay.h_data; S.h_fwd = st + lay.h_fwd; S.h_q = st + lay.h_q;
    } else {
        p.partial = part0;
        if (pos && (rc = run_fused(F_STAT_P, p, t))) return rc;
        ReduceSet &S = add_set(part0, t, true, false, gy_pos);
        S.vh_p = st + lay.vh_data; S.h_sum = st + lay.h_data; S.h_fwd = st + lay.h_fwd;
        p.partial = part2;
        p.ds = 0;                                   // the sparsity statistic uses the forward strand only (convRBM.py:443)
        if (pos && (rc = run_fused(F_STAT_Q, p, t))) return rc;
        ReduceSet &Sq = add_set(part2, t, false, true, gy_pos);
        Sq.vh_q = st + lay.vh_q; Sq.h_q = st + lay.h_q;
    }

    // ---- per-CTA partial sums of both phases -> statistics vector ---------------------------------------------------------
    if (gen) {
        if (phases & CRBM_PHASE_REDUCE) {
            GenCountParams c{};
            c.rows[0] = a->data; c.row_words[0] = a->data_row_words; c.n[0] = a->n_data; c.L[0] = a->L;
            c.rows[1] = a->v_model; c.row_words[1] = a->v_row_words; c.n[1] = a->batchsize; c.L[1] = Lv;
            c.batch_index = a->batch_index_dev;
            c.out[0] = st + lay.v_data; c.out[1] = st + lay.v_model; c.scalars = st + lay.n_hidden_data;
            c.c0 = (float)a->n_data * Lh; c.c1 = (float)a->batchsize * FANTASY_LEN; c.c2 = (float)a->n_data * a->L; c.c3 = (float)a->batchsize * Lv;
            gen_counts_kernel<<<2, 256, 0, s>>>(c);
            CRBM_CHECK_LAUNCH();
        }
        if (phases & CRBM_PHASE_APPLY) return crbm_train_apply(a, stream);
        return CRBM_OK;
    }
    if (phases & CRBM_PHASE_REDUCE) {
        ReduceSet &Sn = add_set(part1, tn, true, false, gy_neg);
        Sn.vh_p = st + lay.vh_model; Sn.h_sum = st + lay.h_model;
        red.counts = st + lay.n_hidden_data;
        red.c0 = (float)a->n_data * Lh; red.c1 = (float)a->batchsize * FANTASY_LEN;
        red.c2 = (float)a->n_data * a->L; red.c3 = (float)a->batchsize * Lv;
        red.cnt_rows[0] = a->data; red.cnt_row_words[0] = a->data_row_words; red.cnt_n[0] = a->n_data; red.cnt_L[0] = a->L;
        red.cnt_rows[1] = a->v_model; red.cnt_row_words[1] = a->v_row_words; red.cnt_n[1] = a->batchsize; red.cnt_L[1] = Lv;
        red.batch_index = a->batch_index_dev;
        red.cnt_out[0] = st + lay.v_data; red.cnt_out[1] = st + lay.v_model;
        // single rank, tensor-pipe layout, update requested in the same call: reduction and update in one launch
        if ((phases & CRBM_PHASE_APPLY) && mma && a->n_data > 0 && fused_update_enabled() && (!p2p || p2p->peer_ll[0])) {
            FusedUpdateParams f{};
            f.red = red;
            fill_apply_params(a, &f.ap);
            f.per = 8 * KS;
            f.sd = 0; f.sq = combined ? 0 : 1; f.sm = red.nsets - 1; f.q_off = combined ? f.per : 0;
            const int n_cta = red.ktiles * (f.per + 2) + 1;
            if (!p2p) {
                CRBM_CHECK_CUDA(launch_chain(reduce_update_kernel, dim3(n_cta), dim3(512), 0, s, f));
            } else {
                // data parallel: reduction, all-reduce over NVLink peer memory and update in one launch (reduce_update_ll_kernel)
                FusedLLParams L{};
                L.f = f;
                fill_ll_view(a, p2p, &L.ll);
                CRBM_CHECK_ARG(p2p->ll_parity_stride >= (size_t)n_cta * CRBM_MAX_PEERS * LL_ROWS * 32 && p2p->error_flag,
                               "crbm_train_step_p2p: receive buffer too small (crbm_p2p_ll_elems)");
                CRBM_CHECK_CUDA(launch_chain(reduce_update_ll_kernel, dim3(n_cta), dim3(512), 0, s, L));
            }
            CRBM_CHECK_LAUNCH();
            return CRBM_OK;
        }
        if (p2p && (phases & CRBM_PHASE_APPLY)) { set_error("crbm_train_step_p2p: internal: the fused exchange is not available for this step"); return CRBM_ERR_UNSUPPORTED; }
        CRBM_CHECK_CUDA(launch_chain(reduce_partials_kernel, dim3(red.block_start[red.nsets] + 2), dim3(128), 0, s, red));
        CRBM_CHECK_LAUNCH();
    }
    if (phases & CRBM_PHASE_APPLY) return crbm_train_apply(a, stream);
    return CRBM_OK;
}

int crbm_train_phases(const crbm_train_args *a, int phases, void *stream) {
    CRBM_CHECK_ARG(phases > 0 && phases < 16, "crbm_train_phases: bad phase mask 0x%x", phases);
    if (phases & CRBM_PHASE_APPLY) CRBM_CHECK_ARG(a && a->vW && a->vb && a->vc, "crbm_train_phases: null velocity pointer in args");
    return train_phases_impl(a, phases, stream);
}

extern "C" int crbm_train_apply_p2p(const crbm_train_args *a, const crbm_p2p_args *p2p, void *stream);
int crbm_train_step_onehot(const crbm_train_args *a, const crbm_p2p_args *p2p, const float *host_onehot, float *staging, int *bad_count,
                           void *copy_stream, void *consumed_event, void *ready_event, void *stream) {
    CRBM_CHECK_ARG(a && a->vW && a->vb && a->vc && host_onehot && staging && bad_count && copy_stream && consumed_event && ready_event,
                   "crbm_train_step_onehot: null argument");
    CRBM_CHECK_ARG(a->n_data > 0 && !a->batch_index_dev, "crbm_train_step_onehot: needs a non-empty minibatch addressed directly");
    cudaStream_t s = (cudaStream_t)stream, cs = (cudaStream_t)copy_stream;
    cudaEvent_t consumed = (cudaEvent_t)consumed_event, ready = (cudaEvent_t)ready_event;
    // chain sweeps first: they do not read the minibatch, so the device is busy while the batch crosses PCIe
    int rc = train_phases_impl(a, CRBM_PHASE_NEGATIVE, stream);
    if (rc) return rc;
    CRBM_CHECK_CUDA(cudaStreamWaitEvent(cs, consumed, 0));      // previous step's positive phase has read the packed rows
    CRBM_CHECK_CUDA(cudaMemcpyAsync(staging, host_onehot, sizeof(float) * 4 * (size_t)a->n_data * a->L, cudaMemcpyDefault, cs));
    rc = crbm_pack_onehot_accumulate(staging, a->n_data, a->L, const_cast<uint32_t *>(a->data), a->data_row_words, bad_count, copy_stream);
    if (rc) return rc;
    CRBM_CHECK_CUDA(cudaEventRecord(ready, cs));
    CRBM_CHECK_CUDA(cudaStreamWaitEvent(s, ready, 0));
    if (p2p) {
        // data parallel: this rank's shard; the reduction publishes the statistics to the peers, the update sums them over NVLink
        CRBM_CHECK_ARG(p2p->n_ranks > 0 && p2p->n_ranks <= CRBM_MAX_PEERS && p2p->rank >= 0 && p2p->rank < p2p->n_ranks && p2p->ticket &&
                       (p2p->parity == 0 || p2p->parity == 1), "crbm_train_step_onehot: bad p2p arguments");
        rc = train_phases_impl(a, CRBM_PHASE_POSITIVE | CRBM_PHASE_REDUCE, stream, p2p);
        if (!rc) rc = crbm_train_apply_p2p(a, p2p, stream);
    } else {
        rc = train_phases_impl(a, CRBM_PHASE_POSITIVE | CRBM_PHASE_REDUCE | CRBM_PHASE_APPLY, stream);
    }
    if (rc) return rc;
    CRBM_CHECK_CUDA(cudaEventRecord(consumed, s));
    return CRBM_OK;
}

static int train_step_packjob(const crbm_train_args *a, const crbm_p2p_args *p2p, const PackJob *job, void *stream);
int crbm_train_step_hostfed(const crbm_train_args *a, const crbm_p2p_args *p2p, const float *onehot, int *bad_count, void *stream) {
    CRBM_CHECK_ARG(a && a->vW && a->vb && a->vc && onehot && bad_count, "crbm_train_step_hostfed: null argument");
    CRBM_CHECK_ARG(a->n_data > 0 && !a->batch_index_dev, "crbm_train_step_hostfed: needs a non-empty minibatch addressed directly");
    PackJob job{onehot, bad_count, nullptr, 0u};
    return train_step_packjob(a, p2p, &job, stream);
}

int crbm_train_step_staged(const crbm_train_args *a, const crbm_p2p_args *p2p, const float *host_onehot, float *staging_dev,
                           unsigned int *ready_flag_dev, const unsigned int *seq_host, int *bad_count, void *copy_stream, void *stream) {
    CRBM_CHECK_ARG(a && a->vW && a->vb && a->vc && host_onehot && staging_dev && ready_flag_dev && seq_host && bad_count && copy_stream,
                   "crbm_train_step_staged: null argument");
    CRBM_CHECK_ARG(a->n_data > 0 && !a->batch_index_dev, "crbm_train_step_staged: needs a non-empty minibatch addressed directly");
    cudaStream_t cs = (cudaStream_t)copy_stream;
    CRBM_CHECK_CUDA(cudaMemcpyAsync(staging_dev, host_onehot, sizeof(float) * 4 * (size_t)a->n_data * a->L, cudaMemcpyHostToDevice, cs));
    CRBM_CHECK_CUDA(cudaMemcpyAsync(ready_flag_dev, seq_host, sizeof(unsigned int), cudaMemcpyHostToDevice, cs));
    PackJob job{staging_dev, bad_count, ready_flag_dev, *seq_host};
    return train_step_packjob(a, p2p, &job, stream);
}

// the fused exchange (reduce_update_ll_kernel) serves the mma.sync partial layout with a non-empty shard on every rank (the caller
// passes the receive buffers only then); everything else exchanges the statistics vector (crbm_train_apply_p2p)
static bool ll_exchange_possible(const crbm_train_args *a, const crbm_p2p_args *p2p) {
    if (!p2p || !p2p->peer_ll[0] || a->n_data <= 0 || !fused_update_enabled()) return false;
    if (shape_is_generic(a->M, a->pooling)) return false;
    return train_uses_tc(a->K, a->M, a->pooling) || a->pooling == 1 || pool_uses_mma();
}
static int train_step_packjob(const crbm_train_args *a, const crbm_p2p_args *p2p, const PackJob *job, void *stream) {
    const int all = CRBM_PHASE_NEGATIVE | CRBM_PHASE_POSITIVE | CRBM_PHASE_REDUCE;
    if (p2p) {
        CRBM_CHECK_ARG(p2p->n_ranks > 0 && p2p->n_ranks <= CRBM_MAX_PEERS && p2p->rank >= 0 && p2p->rank < p2p->n_ranks && p2p->ticket &&
                       (p2p->parity == 0 || p2p->parity == 1), "crbm_train_step_p2p: bad p2p arguments");
        if (ll_exchange_possible(a, p2p)) return train_phases_impl(a, all | CRBM_PHASE_APPLY, stream, p2p, job);
        crbm_p2p_args plain = *p2p;
        for (int r = 0; r < CRBM_MAX_PEERS; ++r) plain.peer_ll[r] = nullptr;
        int rc = train_phases_impl(a, all, stream, &plain, job);
        if (!rc) rc = crbm_train_apply_p2p(a, &plain, stream);
        return rc;
    }
    return train_phases_impl(a, all | CRBM_PHASE_APPLY, stream, nullptr, job);
}

int crbm_train_step_p2p(const crbm_train_args *a, const crbm_p2p_args *p2p, void *stream) {
    CRBM_CHECK_ARG(a && a->vW && a->vb && a->vc && p2p, "crbm_train_step_p2p: null argument");
    for (int r = 0; r < p2p->n_ranks && r < CRBM_MAX_PEERS; ++r) CRBM_CHECK_ARG(p2p->peer_stats[r] && p2p->peer_flags[r], "crbm_train_step_p2p: null peer pointer");
    return train_step_packjob(a, p2p, nullptr, stream);
}

size_t crbm_p2p_ll_elems(int K, int M, int pooling) {
    if (K <= 0 || M <= 0 || pooling <= 0 || shape_is_generic(M, pooling)) return 0;
    size_t n_cta;
    if (train_uses_tc(K, M, pooling)) n_cta = (size_t)ceil_div(K, TT_KC) * (4 * M + 1) + 1;         // reduce_tc_kernel
    else if (pooling == 1 || pool_uses_mma()) n_cta = (size_t)ceil_div(K, 16) * (8 * ceil_div(M, 4) + 2) + 1;   // reduce_update_ll_kernel
    else return 0;
    return n_cta * CRBM_MAX_PEERS * LL_ROWS * 32;
}

int crbm_train_accumulate(const crbm_train_args *a, void *stream) {
    return train_phases_impl(a, CRBM_PHASE_NEGATIVE | CRBM_PHASE_POSITIVE | CRBM_PHASE_REDUCE, stream);
}

int crbm_train_accumulate_p2p(const crbm_train_args *a, const crbm_p2p_args *p2p, void *stream) {
    CRBM_CHECK_ARG(p2p && p2p->n_ranks > 0 && p2p->n_ranks <= CRBM_MAX_PEERS && p2p->rank >= 0 && p2p->rank < p2p->n_ranks &&
                   p2p->ticket && (p2p->parity == 0 || p2p->parity == 1), "crbm_train_accumulate_p2p: bad p2p arguments");
    for (int r = 0; r < p2p->n_ranks; ++r) CRBM_CHECK_ARG(p2p->peer_flags[r], "crbm_train_accumulate_p2p: null peer flag pointer");
    return train_phases_impl(a, CRBM_PHASE_NEGATIVE | CRBM_PHASE_POSITIVE | CRBM_PHASE_REDUCE, stream, p2p);
}

int crbm_train_apply_p2p(const crbm_train_args *a, const crbm_p2p_args *p2p, void *stream) {
    CRBM_CHECK_ARG(a && a->W && a->b && a->c && a->vW && a->vb && a->vc, "crbm_train_apply_p2p: null pointer in args");
    CRBM_CHECK_ARG(p2p && p2p->n_ranks > 0 && p2p->n_ranks <= CRBM_MAX_PEERS && p2p->rank >= 0 && p2p->rank < p2p->n_ranks &&
                   p2p->error_flag && p2p->ticket && (p2p->parity == 0 || p2p->parity == 1), "crbm_train_apply_p2p: bad p2p arguments");
    for (int r = 0; r < p2p->n_ranks; ++r) CRBM_CHECK_ARG(p2p->peer_stats[r] && p2p->peer_flags[r], "crbm_train_apply_p2p: null peer pointer");
    ApplyP2PParams p{};
    ApplyParams &ap = p.ap;
    ap.W = a->W; ap.b = a->b; ap.c = a->c; ap.vW = a->vW; ap.vb = a->vb; ap.vc = a->vc; ap.stats = nullptr;
    fill_layout(a->K, a->M, &ap.lay);
    ap.K = a->K; ap.M = a->M; ap.ds = a->doublestranded ? 1 : 0;
    ap.lr = a->learning_rate; ap.mu = a->momentum; ap.rho = a->rho; ap.lam = a->lambda_rate;
    ap.step_dev = a->step_dev; ap.batch_index_dev = a->batch_index_dev; ap.n_batches = a->n_batches;
    ap.c_trace = a->c_trace_slots > 0 ? a->c_trace : nullptr; ap.c_trace_slots = a->c_trace_slots; ap.step = a->step;
    p.n_ranks = p2p->n_ranks; p.parity = p2p->parity;
    for (int r = 0; r < p2p->n_ranks; ++r) p.peer_stats[r] = p2p->peer_stats[r];
    p.local_flags = p2p->peer_flags[p2p->rank];
    for (int r = 0; r < p2p->n_ranks; ++r) p.peer_flags[r] = p2p->peer_flags[r];
    p.rank = p2p->rank;
    p.stats_stride = p2p->stats_stride; p.sequence = p2p->sequence; p.sequence_dev = a->step_dev; p.error_flag = p2p->error_flag;
    p.ticket = p2p->ticket;
    const int n = a->K * 4 * a->M + a->K + 4;
    CRBM_CHECK_CUDA(launch_chain(apply_update_p2p_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, p));
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

// ---- the persistent chain as the host sees it ------------------------------------------------------------------------------
int crbm_chain_layout(int K, int M, int pooling) { return (K > 0 && M > 0 && pooling > 0 && train_uses_tc(K, M, pooling)) ? 1 : 0; }
size_t crbm_chain_words(int K, int M, int pooling) {
    if (K <= 0 || M <= 0 || pooling <= 0) return 0;
    return train_uses_tc(K, M, pooling) ? tt_chain_words(K, pooling) : (size_t)ceil_div(K, 32) * FANTASY_WORDS;
}
extern "C" int crbm_hidden_bits_to_dense(const uint32_t *bits, int B, int K, float *dense, void *stream);
extern "C" int crbm_hidden_dense_to_bits(const float *dense, int B, int K, uint32_t *bits, void *stream);
int crbm_chain_to_dense(const uint32_t *bits, int B, int K, int M, int pooling, float *dense, void *stream) {
    CRBM_CHECK_ARG(bits && dense && B > 0 && K > 0 && M > 0 && pooling > 0, "crbm_chain_to_dense: bad arguments");
    if (!train_uses_tc(K, M, pooling)) return crbm_hidden_bits_to_dense(bits, B, K, dense, stream);
    const TtGeom g = tt_geom(K, M, pooling);
    const long long n = (long long)B * K * FANTASY_LEN;
    chain_lin_to_dense_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bits, B, K, g.Kpad, g.hw, g.VP, dense);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}
int crbm_chain_from_dense(const float *dense, int B, int K, int M, int pooling, uint32_t *bits, void *stream) {
    CRBM_CHECK_ARG(bits && dense && B > 0 && K > 0 && M > 0 && pooling > 0, "crbm_chain_from_dense: bad arguments");
    if (!train_uses_tc(K, M, pooling)) return crbm_hidden_dense_to_bits(dense, B, K, bits, stream);
    const TtGeom g = tt_geom(K, M, pooling);
    const long long n = (long long)B * g.hw * g.Kpad;
    chain_dense_to_lin_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dense, B, K, g.Kpad, g.hw, g.VP, bits);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

// debugging aid (not part of the ABI in include/crbm_b200.h): clock64() marks of CTA 0 of the last h2v_tc_kernel launch (CRBM_H2V_DEBUG=1)
int crbm_debug_h2v_timeline(long long *host16) {
    CRBM_CHECK_CUDA(cudaMemcpyFromSymbol(host16, g_h2v_timeline, 16 * sizeof(long long)));
    return CRBM_OK;
}

int crbm_train_step(const crbm_train_args *a, void *stream) {
    CRBM_CHECK_ARG(a && a->vW && a->vb && a->vc, "crbm_train_step: null pointer in args");
    return train_phases_impl(a, CRBM_PHASE_NEGATIVE | CRBM_PHASE_POSITIVE | CRBM_PHASE_REDUCE | CRBM_PHASE_APPLY, stream);
}

int crbm_train_apply(const crbm_train_args *a, void *stream) {
    CRBM_CHECK_ARG(a && a->W && a->b && a->c && a->vW && a->vb && a->vc && a->stats, "crbm_train_apply: null pointer in args");
    ApplyParams p{};
    fill_apply_params(a, &p);
    const int n = a->K * 4 * a->M + a->K + 1;
    CRBM_CHECK_CUDA(launch_chain(apply_update_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, p));
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

}  // extern "C"
