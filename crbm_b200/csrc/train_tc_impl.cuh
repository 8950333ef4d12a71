// Host side and small kernels of the tcgen05 training path (included by train.cu after the shared update arithmetic):
//   h2v_chunk_kernel      A5 on the LINEAR chain planes the tcgen05 kernels write (bit = position, word-major, motif contiguous),
//   h2v_finish_kernel     per motif chunk with its weights in shared memory; A6 (softmax, sampling) after the chunks are added
//   reduce_tc_kernel      per-CTA statistics partials -> statistics vector, or -> sparsity factor + momentum update in the same launch
//   chain <-> dense       conversions for tests / inspection
//   train_phases_tc       the launch chain of one step
#pragma once
#include "train_tc.cuh"
#include "scan_rt.cuh"
#include "scan_rd.cuh"
#include "h2v_tc.cuh"

namespace crbm {

#define CRBM_FOR_EACH_TT_P(X) X(1) X(2) X(4) X(5) X(10)
#define CRBM_DECLARE_TT(PV) template <> int launch_v2h_train_tc<PV>(int, const TrainTcParams &, dim3, cudaStream_t);
CRBM_FOR_EACH_TT_P(CRBM_DECLARE_TT)
#undef CRBM_DECLARE_TT

#define CRBM_DECLARE_RT(PV) template <> int launch_v2h_rows_tc<PV>(int, const RowsTcParams &, dim3, cudaStream_t);
CRBM_FOR_EACH_TT_P(CRBM_DECLARE_RT)
#undef CRBM_DECLARE_RT

#define CRBM_DECLARE_RD(PV) template <> int launch_v2h_rowdense_tc<PV>(const RowDenseParams &, const CUtensorMap &, dim3, cudaStream_t);
CRBM_FOR_EACH_TT_P(CRBM_DECLARE_RD)
#undef CRBM_DECLARE_RD

static bool tt_pool_built(int P) { return P == 1 || P == 2 || P == 4 || P == 5 || P == 10; }
// scan reductions / free energy on the tcgen05 rows kernel (scan_rt.cuh): motif_length <= 32, pooling one of the built factors;
// CRBM_ROWS_KERNEL=mma keeps the mma.sync / CUDA-core kernels
static bool rows_uses_tc(int M, int P) {
    static int v = -1;
    if (v < 0) { const char *e = getenv("CRBM_ROWS_KERNEL"); v = (e && e[0] == 'm') ? 0 : 1; }
    return v == 1 && M <= 32 && tt_pool_built(P);
}
static int launch_rows_tc(int mode, int P, RowsTcParams &p, cudaStream_t s) {
    p.Kdim = 16 * ceil_div(4 * p.M, 16);
    p.ksteps = p.Kdim / 16;
    p.n_atoms = ceil_div(p.Kdim, 64);
    const int VP = (32 / P) * P;
    p.nblk32 = ceil_div(p.Lh, VP);
    p.n_quads = ceil_div(p.n_rows, 4);
    p.xstages = TT_MAX_XSTAGES;
    const int n_chunks = ceil_div(p.K, mode == RT_REDUCE ? TT_ROWS : TT_KC);
    int gx = num_sms() / n_chunks > 0 ? num_sms() / n_chunks : 1;
    if (gx > p.n_quads) gx = p.n_quads;
    // CRBM_RT_DEBUG=1 (timing aid, synchronous): clock64() marks of CTA (0,0), printed as cycles since the CTA's first mark
    static const int rt_debug = [] { const char *e = getenv("CRBM_RT_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
    p.dbg = nullptr;
    const size_t dbg_n = (size_t)RT_DBG_ROLES * RT_DBG_UNITS * 4;
    if (rt_debug) {
        CRBM_CHECK_CUDA(cudaMalloc(&p.dbg, dbg_n * sizeof(long long)));
        CRBM_CHECK_CUDA(cudaMemsetAsync(p.dbg, 0, dbg_n * sizeof(long long), s));
    }
    int rc = CRBM_ERR_UNSUPPORTED;
    switch (P) {
#define CRBM_CASE(PV) case PV: rc = launch_v2h_rows_tc<PV>(mode, p, dim3(gx, n_chunks), s); break;
        CRBM_FOR_EACH_TT_P(CRBM_CASE)
#undef CRBM_CASE
        default: set_error("tcgen05 rows kernel: pooling %d not built", P);
    }
    if (rt_debug && rc == CRBM_OK) {
        static int printed = 0;
        long long *h = (long long *)malloc(dbg_n * sizeof(long long));
        CRBM_CHECK_CUDA(cudaStreamSynchronize(s));
        CRBM_CHECK_CUDA(cudaMemcpy(h, p.dbg, dbg_n * sizeof(long long), cudaMemcpyDeviceToHost));
        if (!printed++) {
            long long t0 = 0;
            for (size_t i = 0; i < dbg_n; ++i) if (h[i] && (!t0 || h[i] < t0)) t0 = h[i];
            const char *names[RT_DBG_ROLES] = {"producer0", "producer1", "mma", "epilogue0", "epilogue1"};
            for (int r = 0; r < RT_DBG_ROLES; ++r)
                for (int t = 0; t < RT_DBG_UNITS; ++t) {
                    const long long *e = h + ((size_t)r * RT_DBG_UNITS + t) * 4;
                    if (e[0]) fprintf(stderr, "RT_DBG %s unit %d: %lld %lld %lld %lld\n", names[r], t, e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0);
                }
        }
        free(h);
        cudaFree(p.dbg);
    }
    return rc;
}
// Dense scan with the motif on the TMEM lane and TMA stores (scan_rd.cuh). Pooled probabilities always go there (the position-major
// kernel of scan_tc.cu needs shuffles and writes partial sectors). Un-pooled scans go there when the 128-row MMA tile is well filled:
// measured on B200 (% of the HBM copy peak, row-dense against position-major): K=128 93 / 78, K=384 85 / 80, K=512 86 / 81,
// K=100 77 / 70, but K=200 68 / 80, K=300 72 / 82, K=150 54 / 79 (idle TMEM lanes cost epilogue slots), K=64 52 / 52.
// CRBM_SCAN_KERNEL=rows | pos forces one of them, CRBM_POOL_SCAN=pos keeps the pooled scan on scan_tc.cu. Needs output rows that are
// multiples of 4 floats on a 16-byte base.
static bool rowdense_wanted(int K, int M, int P, long long row_stride, const void *out) {
    static int all = -1, pooled = -1;
    if (all < 0) { const char *e = getenv("CRBM_SCAN_KERNEL"); all = !e ? 2 : e[0] == 'r' ? 1 : e[0] == 'p' ? 0 : 2; }
    if (pooled < 0) { const char *e = getenv("CRBM_POOL_SCAN"); pooled = (e && e[0] == 'p') ? 0 : 1; }
    if (M > 32 || !tt_pool_built(P) || (row_stride & 3) || ((uintptr_t)out & 15)) return false;
    if (P > 1) return pooled == 1;
    if (all != 2) return all == 1;
    const int padded = TT_ROWS * ceil_div(K, TT_ROWS);
    return 10 * K >= 9 * padded || (K <= TT_ROWS && 4 * K >= 3 * padded);
}
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TensorMapEncodeTiledFn>(ptr);
        else
            (void)cudaGetLastError();
    }
    return fn;
}
static int launch_rowdense_tc(int P, RowDenseParams &p, cudaStream_t s) {
    p.Kdim = 16 * ceil_div(4 * p.M, 16);
    p.ksteps = p.Kdim / 16;
    p.n_atoms = ceil_div(p.Kdim, 64);
    const int VP = (32 / P) * P;
    p.nblk32 = ceil_div(p.Lh, VP);
    p.nsteps = ceil_div(p.nblk32, 4);
    p.xstages = RD_XSTAGES;
    const long long len8 = 8 * (long long)ceil_div(p.Lh, 8);
    p.store_len = (int)(p.row_stride < len8 ? p.row_stride : len8);
    const int n_chunks = ceil_div(p.K, TT_ROWS);
    int gx = num_sms() / n_chunks > 0 ? num_sms() / n_chunks : 1;
    if (gx > p.n_rows) gx = p.n_rows;
    // VP = 32: a warp's tile is stored through a tensor map over (position, motif, sequence) of fp32, box 32 x 32 x 1, SWIZZLE_128B;
    // CRBM_RD_STORE=bulk keeps the per-row bulk copies
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    static int want_tensor = -1;
    if (want_tensor < 0) { const char *e = getenv("CRBM_RD_STORE"); want_tensor = (e && e[0] == 'b') ? 0 : 1; }
    p.tensor_store = 0;
    { const char *e = getenv("CRBM_RD_DEBUG"); p.debug = e ? atoi(e) : 0; }
    if (VP == 32 && want_tensor && tensor_map_encoder()) {
        const cuuint64_t dims[3] = {(cuuint64_t)p.store_len, (cuuint64_t)p.K, (cuuint64_t)p.n_rows};
        const cuuint64_t strides[2] = {(cuuint64_t)p.row_stride * 4, (cuuint64_t)p.row_stride * 4 * (cuuint64_t)p.K};
        const cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
        const CUresult r = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.out, dims, strides, box, estr,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) p.tensor_store = 1;
    }
    switch (P) {
#define CRBM_CASE(PV) case PV: return launch_v2h_rowdense_tc<PV>(p, tmap, dim3(gx, n_chunks), s);
        CRBM_FOR_EACH_TT_P(CRBM_CASE)
#undef CRBM_CASE
    }
    set_error("tcgen05 row-dense scan kernel: pooling %d not built", P);
    return CRBM_ERR_UNSUPPORTED;
}
// The tcgen05 path covers motif_length <= 32 (4 M <= 128 contraction columns, one MMA N) and the pooling factors it is instantiated
// for; everything else trains on the mma.sync kernels of v2h_mma.cuh. Small un-pooled banks (K * M < 2048, e.g. 100 x 15) also stay
// there: at that size a step is a chain of four 5-20 us launches, and the mma.sync kernels (two CTAs per SM, prologues overlapping
// the predecessor's tail) are the faster ones (measured at C2: 42.0 us per step against 48.7 us on the tcgen05 kernels, whose CTAs own
// the whole SM). The choice depends on the model shape only, because the layout of the persistent chain follows it.
// CRBM_TRAIN_KERNEL = mma | tc overrides.
static bool train_uses_tc(int K, int M, int P) {
    static int v = -1;
    if (v < 0) { const char *e = getenv("CRBM_TRAIN_KERNEL"); v = !e ? 2 : e[0] == 'm' ? 0 : e[0] == 't' ? 1 : 2; }
    if (M > 32 || !tt_pool_built(P) || v == 0 || shape_is_generic(M, P)) return false;
    return v == 1 || P > 1 || (long long)K * M >= 2048;
}

struct TtGeom { int VP, Kdim, ksteps, n_atoms, n_chunks, Kpad, G, hw, part_floats; };
static TtGeom tt_geom(int K, int M, int P) {
    TtGeom g;
    g.VP = (32 / P) * P;
    g.Kdim = 16 * ceil_div(4 * M, 16);
    g.ksteps = g.Kdim / 16;
    g.n_atoms = ceil_div(g.Kdim, 64);
    g.n_chunks = ceil_div(K, TT_KC);
    g.Kpad = g.n_chunks * TT_KC;
    g.G = num_sms() / g.n_chunks > 0 ? num_sms() / g.n_chunks : 1;
    g.hw = ceil_div(FANTASY_LEN, g.VP);
    g.part_floats = tt_part_floats(g.Kdim);
    return g;
}
static size_t tt_workspace_floats(int K, int M, int batchsize) {
    if (M > 32) return 0;
    const TtGeom g = tt_geom(K, M, 1);
    // two sets of per-CTA statistics partials + the per-chunk visible activities of the chain (h2v_chunk_kernel) + the weight image
    // of the dense h2v GEMM (h2v_tc.cuh: 2 strands x Kpad / 64 k-chunks of 32 KB)
    return 2 * (size_t)g.n_chunks * g.G * g.part_floats + (size_t)g.n_chunks * batchsize * (32 * ceil_div(FANTASY_LEN + M - 1, 32)) * 4 +
           (size_t)2 * g.n_chunks * (2 * HT_A_BYTES / 4);
}
// h2v of the tcgen05 path: dense GEMM on the tensor cores (h2v_tc.cuh) or the sparse walk over the set bits (h2v_chunk_kernel)
static bool h2v_uses_tc(int kk) {
    static int v = -1;
    if (v < 0) { const char *e = getenv("CRBM_H2V_KERNEL"); v = !e ? 2 : e[0] == 'c' ? 0 : e[0] == 't' ? 1 : 2; }
    return v == 1 || (v == 2 && kk >= 512);
}
static size_t tt_chain_words(int K, int P) {           // per (strand, particle)
    const int VP = (32 / P) * P;
    return (size_t)ceil_div(FANTASY_LEN, VP) * ceil_div(K, TT_KC) * TT_KC;
}

static int dispatch_train_tc(int P, int flags, const TrainTcParams &p, dim3 grid, cudaStream_t s) {
    switch (P) {
#define CRBM_CASE(PV) case PV: return launch_v2h_train_tc<PV>(flags, p, grid, s);
        CRBM_FOR_EACH_TT_P(CRBM_CASE)
#undef CRBM_CASE
    }
    set_error("tcgen05 training kernel: pooling %d not built", P);
    return CRBM_ERR_UNSUPPORTED;
}

// ---- A5 + A6 on the linear chain planes: motif chunks with the weight slice in shared memory -----------------------------------
// The planes are uint32 H[strand][n][w][Kpad]: bit b of word w of motif k is hidden position w * VP + b (32 consecutive motifs of one
// word index = one 128-byte line). The first version walked 16-position items over ALL motifs with the weights in L1/L2 (the scheme
// of h2v_bits_kernel, train.cu): ncu at K = 512 (profiles/r02_train_tc.md) showed 40.7 M warp instructions for 0.39 M hits, issue
// slots 68 % busy — bound by the instructions of its per-round bookkeeping (80 rounds per item) and by four scalar weight loads per
// hit, not by memory. Here a CTA owns a chunk of 64 motifs whose weights sit in shared memory transposed to
// W4[kl][m] = (W[k,A,m], W[k,C,m], W[k,G,m], W[k,T,m]): a hit costs ONE 16-byte shared-memory load per lane. A work item is 32
// visible positions of one particle (lane = position, four base accumulators); the lane first acts as "motif lane" and gathers the
// set bits of all words of both strands and both 32-motif tiles that reach the item in ONE compaction round. The chunk's partial
// activities go to act_part[chunk][n][i] (float4); h2v_finish_kernel adds the chunks in a fixed order, takes the 4-way softmax and
// draws the base. Per hit and item: ~12 instructions instead of ~100.
struct H2VChunkParams {
    const uint32_t *bits;
    int Bn, K, Kpad, hw, VP, M, ds, items_per_row, lv_pad;
    const float *W;
    float4 *act_part;          // [n_chunks][Bn][lv_pad]
};
constexpr int H2C_WARPS = 8;
constexpr int H2C_LIST = 1024 + 4;

__global__ void __launch_bounds__(32 * H2C_WARPS) h2v_chunk_kernel(const H2VChunkParams p) {
    extern __shared__ __align__(128) unsigned char h2c_smem[];
    float4 *W4 = reinterpret_cast<float4 *>(h2c_smem);                                   // [64][M]
    uint32_t *lists = reinterpret_cast<uint32_t *>(h2c_smem + (size_t)TT_KC * p.M * 16);
    const int M = p.M, chunk = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_wait();
    pdl_launch_dependents();
    for (int e = threadIdx.x; e < TT_KC * M; e += blockDim.x) {
        const int kl = e / M, m = e - kl * M, k = chunk * TT_KC + kl;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < p.K) {
            const float *src = p.W + (size_t)k * 4 * M + m;
            w = make_float4(__ldg(src), __ldg(src + M), __ldg(src + 2 * M), __ldg(src + 3 * M));
        }
        W4[e] = w;
    }
    __syncthreads();
    uint32_t *list = lists + warp * H2C_LIST;
    const int strands = p.ds ? 2 : 1;
    const int n_items = p.Bn * p.items_per_row;
    for (int item = blockIdx.x * H2C_WARPS + warp; item < n_items; item += gridDim.x * H2C_WARPS) {
        const int n = item / p.items_per_row, w32 = item - n * p.items_per_row;
        const int i = 32 * w32 + lane;
        const int jlo = max(0, 32 * w32 - (M - 1)), jhi = min(FANTASY_LEN - 1, 32 * w32 + 31);
        const int wlo = jlo / p.VP, whi = jhi / p.VP;
        // this lane as motif lane: its words (strand, tile, word) masked to the hidden positions that reach the item
        uint32_t wd[12];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int k = chunk * TT_KC + 32 * t + lane;
                const uint32_t *plane = p.bits + ((size_t)s * p.Bn + n) * p.hw * p.Kpad + k;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int wq = wlo + q;
                    uint32_t word = 0u;
                    if (s < strands && wq <= whi && k < p.K) {
                        const int b0 = max(0, jlo - wq * p.VP), b1 = min(31, jhi - wq * p.VP);
                        const uint32_t mask = (b1 >= 31 ? 0xffffffffu : ((1u << (b1 + 1)) - 1u)) & ~((1u << b0) - 1u);
                        word = __ldg(plane + (size_t)wq * p.Kpad) & mask;
                    }
                    wd[(s * 2 + t) * 3 + q] = word;
                }
            }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (;;) {
            // up to 32 hits per lane and pass, so that a pass never exceeds the list (dense chains take several passes)
            int mine = 0;
#pragma unroll
            for (int u = 0; u < 12; ++u) mine += __popc(wd[u]);
            if (__ballot_sync(0xffffffffu, mine != 0) == 0u) break;
            const int take = min(mine, 32);
            int incl = take;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            int pos = incl - take, left = take;
#pragma unroll
            for (int u = 0; u < 12; ++u) {
                const int s = u / 6, t = (u / 3) & 1, q = u % 3;
                const uint32_t tagk = ((uint32_t)s << 15) | ((uint32_t)(32 * t + lane) << 8);
                const int jbase = (wlo + q) * p.VP;
                while (wd[u] && left > 0) {
                    const int b = __ffs(wd[u]) - 1;
                    wd[u] &= wd[u] - 1;
                    list[pos++] = tagk | (uint32_t)(jbase + b);
                    --left;
                }
            }
            if (lane < 4) list[total + lane] = 0xFFu;          // padding entries: position 255 never matches
            __syncwarp();
            for (int h0 = 0; h0 < total; h0 += 4) {
                const uint4 d = *reinterpret_cast<const uint4 *>(list + h0);
                const uint32_t dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int m = i - (int)(dd[u] & 0xFFu);
                    const int kl = (dd[u] >> 8) & 63;
                    const bool rc = (dd[u] >> 15) & 1u;        // warp-uniform
                    if (m >= 0 && m < M) {
                        const float4 w = W4[kl * M + (rc ? M - 1 - m : m)];
                        if (rc) { acc.x += w.w; acc.y += w.z; acc.z += w.y; acc.w += w.x; }      // W[k, 3-a, M-1-m] (convRBM.py:285)
                        else { acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w; }
                    }
                }
            }
            __syncwarp();
        }
        p.act_part[((size_t)chunk * p.Bn + n) * p.lv_pad + i] = acc;
    }
}

struct H2VFinishParams {
    const float4 *act_part;
    int Bn, n_chunks, lv_pad, Lv;
    const float *c;
    uint32_t *v_packed;
    int v_row_words;
    const float *u_v;
    uint64_t seed;
    uint32_t ctr_hi, iter;
    const unsigned long long *step_dev;
};
__global__ void __launch_bounds__(256) h2v_finish_kernel(const H2VFinishParams p) {
    pdl_wait();
    pdl_launch_dependents();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = gid / p.lv_pad, i = gid - n * p.lv_pad, lane = threadIdx.x & 31;
    int pick = 0;
    const bool live = n < p.Bn;
    if (live && i < p.Lv) {
        float4 a = make_float4(__ldg(p.c), __ldg(p.c + 1), __ldg(p.c + 2), __ldg(p.c + 3));
        for (int ch = 0; ch < p.n_chunks; ++ch) {              // fixed order: the sample does not depend on timing
            const float4 v = p.act_part[((size_t)ch * p.Bn + n) * p.lv_pad + i];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        const float mx = fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w));
        const float e0 = __expf(a.x - mx), e1 = __expf(a.y - mx), e2 = __expf(a.z - mx), e3 = __expf(a.w - mx);
        const float tot = e0 + e1 + e2 + e3;
        float u;
        if (p.u_v) u = p.u_v[(size_t)n * p.Lv + i];
        else {
            const uint32_t hi = p.ctr_hi + (p.step_dev ? (uint32_t)(*p.step_dev) : 0u);
            u = u01(philox4x32(make_uint4((uint32_t)(n * p.Lv + i), 0u, hi, RNG_VISIBLE | (p.iter << 8)),
                               make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32))).x);
        }
        const float ut = u * tot;
        pick = (e0 > ut) ? 0 : (e0 + e1 > ut) ? 1 : (e0 + e1 + e2 > ut) ? 2 : 3;
    }
    uint32_t v = (uint32_t)pick << (2 * (lane & 15));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    if (live && (lane & 15) == 0) p.v_packed[(size_t)n * p.v_row_words + FRONT_PAD_WORDS + (i >> 4)] = v;
}

// ---- statistics partials -> statistics vector / update -----------------------------------------------------------------------
// partial of CTA x of chunk ch (train_tc.cuh): P[c][r] (Kdim x 128), Q[c][kl] (Kdim x 64), sum_p[128], sum_q[64].
struct TcReduceParams {
    const float *pos, *neg;        // [n_chunks][G][part_floats]; pos == null: empty minibatch shard (contributes zeros)
    int G_pos, G_neg, part_floats, Kdim, n_chunks, K, M, ds;
    float *stats;                  // destination when !fused
    crbm_stats_layout lay;
    float c0, c1, c2, c3;          // N*Lh, B*200, N*L, B*(199+M)
    const uint32_t *cnt_rows[2];
    int cnt_row_words[2], cnt_n[2], cnt_L[2];
    const int *batch_index;
    ApplyParams ap;
    int fused;
    LLView ll;                     // ll.n_ranks > 0 (data parallel, fused): the sums are exchanged over peer memory inside this launch
};
// One CTA of 512 threads per (chunk, contraction column c): thread (kl, slice) adds the partials of the CTAs x = slice, slice + 8, ...
// of its motif for the five sums the update of W[k, c] needs (six independent load streams, L2 resident), the eight slices are
// combined through shared memory in a fixed order — the summation order is part of the contract between the fused and the two-launch
// path (bit-identical parameters) — and the threads of slice 0 apply the update or write the statistics.
constexpr int TCR_SLICES = 8;
constexpr int TCR_BATCH = 5;
__device__ __forceinline__ float tcr_combine(const float (*sh)[TT_KC], int kl) {
    return ((sh[0][kl] + sh[1][kl]) + (sh[2][kl] + sh[3][kl])) + ((sh[4][kl] + sh[5][kl]) + (sh[6][kl] + sh[7][kl]));
}

__global__ void __launch_bounds__(512) reduce_tc_kernel(const TcReduceParams p) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float sh[6][TCR_SLICES][TT_KC];
    __shared__ float loc[5][TT_KC];
    __shared__ float gs[LL_ROWS][32];
    __shared__ float counts[8];
    __shared__ int timed_out_s;
    const ApplyParams &ap = p.ap;
    const int kl = threadIdx.x & (TT_KC - 1), slice = threadIdx.x >> 6;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fourM = 4 * p.M;
    const int n_w = p.n_chunks * fourM;
    const size_t pf = (size_t)p.part_floats;
    const size_t q_off = (size_t)p.Kdim * TT_ROWS, s_off = (size_t)p.Kdim * 192;
    float nhd = p.c0, nhm = p.c1;
    const float strands = p.ds ? 2.f : 1.f;
    // data parallel, fused (crbm_train_step_p2p): the CTA's sums go through the exchange of reduce_update_ll_kernel (train.cu) before the
    // update: rows of 32 (value, sequence) pairs pushed to every rank, gathered and added in rank order
    const bool ll = p.ll.n_ranks > 0;
    const uint32_t seq = ll ? (uint32_t)(p.ll.sequence_dev ? *p.ll.sequence_dev + 1ull : p.ll.sequence) : 0u;
    const int cta = blockIdx.x;
    int timed_out = 0;
    if (threadIdx.x == 0) timed_out_s = 0;
    const float nan = __int_as_float(0x7fc00000);
    if (cta < n_w) {
        const int ch = cta / fourM, c = cta - ch * fourM, cr = fourM - 1 - c;
        const int k = ch * TT_KC + kl;
        float df = 0.f, dr = 0.f, qq = 0.f, mf = 0.f, mr = 0.f, hf = 0.f;
        // The partials of a slice are added in the order x = slice, slice + 8, ... (part of the bit-identity contract); the loads of
        // TCR_BATCH of them are issued together (a rolled loop paid one L2 round trip per partial: ~10 of them in a row at 74 CTAs per
        // chunk). A partial beyond the grid contributes +0.f, which leaves every sum as it is.
        const size_t o_df = (size_t)c * TT_ROWS + kl, o_dr = (size_t)cr * TT_ROWS + TT_KC + kl, o_q = q_off + (size_t)c * TT_KC + kl, o_h = s_off + kl;
        if (p.pos) {
            const float *pos = p.pos + (size_t)ch * p.G_pos * pf;
            for (int x0 = slice; x0 < p.G_pos; x0 += TCR_SLICES * TCR_BATCH) {
                float a0[TCR_BATCH], a1[TCR_BATCH], a2[TCR_BATCH], a3[TCR_BATCH];
#pragma unroll
                for (int u = 0; u < TCR_BATCH; ++u) {
                    const int x = x0 + TCR_SLICES * u;
                    const bool ok = x < p.G_pos;
                    const float *b = pos + (size_t)(ok ? x : slice) * pf;
                    a0[u] = ok ? b[o_df] : 0.f;
                    a1[u] = (ok && p.ds) ? b[o_dr] : 0.f;
                    a2[u] = ok ? b[o_q] : 0.f;
                    a3[u] = ok ? b[o_h] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < TCR_BATCH; ++u) { df += a0[u]; dr += a1[u]; qq += a2[u]; hf += a3[u]; }
            }
        }
        {
            const float *neg = p.neg + (size_t)ch * p.G_neg * pf;
            for (int x0 = slice; x0 < p.G_neg; x0 += TCR_SLICES * TCR_BATCH) {
                float a0[TCR_BATCH], a1[TCR_BATCH];
#pragma unroll
                for (int u = 0; u < TCR_BATCH; ++u) {
                    const int x = x0 + TCR_SLICES * u;
                    const bool ok = x < p.G_neg;
                    const float *b = neg + (size_t)(ok ? x : slice) * pf;
                    a0[u] = ok ? b[o_df] : 0.f;
                    a1[u] = (ok && p.ds) ? b[o_dr] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < TCR_BATCH; ++u) { mf += a0[u]; mr += a1[u]; }
            }
        }
        sh[0][slice][kl] = df; sh[1][slice][kl] = dr; sh[2][slice][kl] = qq;
        sh[3][slice][kl] = mf; sh[4][slice][kl] = mr; sh[5][slice][kl] = hf;
        __syncthreads();
        float vd = 0.f, vq = 0.f, vm = 0.f, hfs = 0.f;
        if (slice == 0) {
            vd = tcr_combine(sh[0], kl) + tcr_combine(sh[1], kl); vq = tcr_combine(sh[2], kl);
            vm = tcr_combine(sh[3], kl) + tcr_combine(sh[4], kl); hfs = tcr_combine(sh[5], kl);
        }
        if (ll) {
            // rows 0-7: data P, Q, model P, forward hidden sum (two rows of 32 motifs each); row 8: N*Lh, B*200 in lanes 0, 1
            if (slice == 0) { loc[0][kl] = vd; loc[1][kl] = vq; loc[2][kl] = vm; loc[3][kl] = hfs; }
            __syncthreads();
            if (warp < 9) {
                const float v = warp < 8 ? loc[warp >> 1][32 * (warp & 1) + lane] : (lane == 0 ? p.c0 : lane == 1 ? p.c1 : 0.f);
                gs[warp][lane] = ll_allreduce_row(p.ll, cta, warp, lane, v, seq, timed_out);
                if (timed_out) timed_out_s = 1;
            }
            __syncthreads();
            const int hrow = kl >> 5, hl = kl & 31;
            vd = gs[hrow][hl]; vq = gs[2 + hrow][hl]; vm = gs[4 + hrow][hl]; hfs = gs[6 + hrow][hl];
            nhd = gs[8][0]; nhm = gs[8][1];
        }
        if (slice == 0 && k < p.K) {
            const float nd = __fmul_rn(strands, nhd), nm = __fmul_rn(strands, nhm);
            const int m = c >> 2, a = c & 3;
            const int gid = (k * 4 + a) * p.M + m;
            if (p.fused) {
                if (timed_out_s) { ap.vW[gid] = nan; ap.W[gid] = nan; }
                else {
                    const float gk = sparsity_factor(hfs, nhd, ap.rho, ap.K);
                    const float vel = next_velocity(ap.vW[gid], vd, nd, vm, nm, __fmul_rn(gk, vq), ap.mu, ap.lr, ap.lam);
                    ap.vW[gid] = vel;
                    ap.W[gid] = __fadd_rn(ap.W[gid], vel);
                }
            } else {
                p.stats[p.lay.vh_data + gid] = vd;
                p.stats[p.lay.vh_q + gid] = vq;
                p.stats[p.lay.vh_model + gid] = vm;
            }
        }
    } else if (cta < n_w + p.n_chunks) {
        const int ch = cta - n_w;
        const int k = ch * TT_KC + kl;
        float hf = 0.f, hr = 0.f, hq = 0.f, mf = 0.f, mr = 0.f;
        if (p.pos) {
            const float *pos = p.pos + (size_t)ch * p.G_pos * pf + s_off;
            for (int x0 = slice; x0 < p.G_pos; x0 += TCR_SLICES * TCR_BATCH) {
                float a0[TCR_BATCH], a1[TCR_BATCH], a2[TCR_BATCH];
#pragma unroll
                for (int u = 0; u < TCR_BATCH; ++u) {
                    const int x = x0 + TCR_SLICES * u;
                    const bool ok = x < p.G_pos;
                    const float *b = pos + (size_t)(ok ? x : slice) * pf;
                    a0[u] = ok ? b[kl] : 0.f; a1[u] = ok ? b[TT_KC + kl] : 0.f; a2[u] = ok ? b[TT_ROWS + kl] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < TCR_BATCH; ++u) { hf += a0[u]; hr += a1[u]; hq += a2[u]; }
            }
        }
        {
            const float *neg = p.neg + (size_t)ch * p.G_neg * pf + s_off;
            for (int x0 = slice; x0 < p.G_neg; x0 += TCR_SLICES * TCR_BATCH) {
                float a0[TCR_BATCH], a1[TCR_BATCH];
#pragma unroll
                for (int u = 0; u < TCR_BATCH; ++u) {
                    const int x = x0 + TCR_SLICES * u;
                    const bool ok = x < p.G_neg;
                    const float *b = neg + (size_t)(ok ? x : slice) * pf;
                    a0[u] = ok ? b[kl] : 0.f; a1[u] = ok ? b[TT_KC + kl] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < TCR_BATCH; ++u) { mf += a0[u]; mr += a1[u]; }
            }
        }
        sh[0][slice][kl] = hf; sh[1][slice][kl] = hr; sh[2][slice][kl] = hq; sh[3][slice][kl] = mf; sh[4][slice][kl] = mr;
        __syncthreads();
        float hfs = 0.f, hrs = 0.f, hqs = 0.f, mfs = 0.f, mrs = 0.f;
        if (slice == 0) {
            hfs = tcr_combine(sh[0], kl); hrs = tcr_combine(sh[1], kl); hqs = tcr_combine(sh[2], kl);
            mfs = tcr_combine(sh[3], kl); mrs = tcr_combine(sh[4], kl);
        }
        if (ll) {
            // rows 0-9: sum_p, sum_pr (data), sum_q, sum_p, sum_pr (model), two rows each; row 10: N*Lh, B*200
            if (slice == 0) { loc[0][kl] = hfs; loc[1][kl] = hrs; loc[2][kl] = hqs; loc[3][kl] = mfs; loc[4][kl] = mrs; }
            __syncthreads();
            if (warp < 11) {
                const float v = warp < 10 ? loc[warp >> 1][32 * (warp & 1) + lane] : (lane == 0 ? p.c0 : lane == 1 ? p.c1 : 0.f);
                gs[warp][lane] = ll_allreduce_row(p.ll, cta, warp, lane, v, seq, timed_out);
                if (timed_out) timed_out_s = 1;
            }
            __syncthreads();
            const int hrow = kl >> 5, hl = kl & 31;
            hfs = gs[hrow][hl]; hrs = gs[2 + hrow][hl]; hqs = gs[4 + hrow][hl]; mfs = gs[6 + hrow][hl]; mrs = gs[8 + hrow][hl];
            nhd = gs[10][0]; nhm = gs[10][1];
        }
        if (slice == 0 && k < p.K) {
            const float nd = __fmul_rn(strands, nhd), nm = __fmul_rn(strands, nhm);
            const float hdata = hfs + hrs, hmodel = mfs + mrs;
            if (p.fused) {
                if (timed_out_s) { ap.vb[k] = nan; ap.b[k] = nan; }
                else {
                    const float gk = sparsity_factor(hfs, nhd, ap.rho, ap.K);
                    const float vel = next_velocity(ap.vb[k], hdata, nd, hmodel, nm, __fmul_rn(gk, hqs), ap.mu, ap.lr, ap.lam);
                    ap.vb[k] = vel;
                    ap.b[k] = __fadd_rn(ap.b[k], vel);
                }
            } else {
                p.stats[p.lay.h_fwd + k] = hfs;
                p.stats[p.lay.h_data + k] = hdata;
                p.stats[p.lay.h_model + k] = hmodel;
                p.stats[p.lay.h_q + k] = hqs;
            }
        }
    } else {
        // base counts of the minibatch and of the sampled visible layer (A9); fused: the update of c and the cursors of the next step
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const uint32_t *rows = p.cnt_rows[w];
            if (w == 0 && p.batch_index) rows += (size_t)(*p.batch_index) * p.cnt_n[0] * p.cnt_row_words[0];
            if (p.cnt_n[w] > 0) count_bases_block(rows, p.cnt_row_words[w], p.cnt_n[w], p.cnt_L[w], counts + 4 * w);
            else if (threadIdx.x < 4) counts[4 * w + threadIdx.x] = 0.f;
            __syncthreads();
        }
        float nvd = p.c2, nvm = p.c3;
        if (ll) {
            // row 0: the eight base counts in lanes 0-7; row 1: N*L and B*(199+M) in lanes 0, 1
            if (warp < 2) {
                const float v = warp == 0 ? (lane < 8 ? counts[lane] : 0.f) : (lane == 0 ? p.c2 : lane == 1 ? p.c3 : 0.f);
                gs[warp][lane] = ll_allreduce_row(p.ll, cta, warp, lane, v, seq, timed_out);
                if (timed_out) timed_out_s = 1;
            }
            __syncthreads();
            if (threadIdx.x < 8) counts[threadIdx.x] = gs[0][threadIdx.x];
            nvd = gs[1][0]; nvm = gs[1][1];
            __syncthreads();
        }
        if (p.fused) {
            if (threadIdx.x < 4) {
                const int a = threadIdx.x;
                if (timed_out_s) { ap.vc[a] = nan; ap.c[a] = nan; }
                else {
                    const float v = next_velocity(ap.vc[a], __fadd_rn(counts[a], counts[3 - a]), nvd, __fadd_rn(counts[4 + a], counts[7 - a]), nvm,
                                                  0.f, ap.mu, ap.lr, 0.f);
                    ap.vc[a] = v;
                    const float cn = __fadd_rn(ap.c[a], v);
                    ap.c[a] = cn;
                    trace_c(ap, a, cn, ap.step_dev ? *ap.step_dev : ap.step);
                }
            }
            __syncthreads();
            if (!ll && threadIdx.x == 0) {
                if (ap.step_dev) *ap.step_dev += 1ull;
                if (ap.batch_index_dev) { const int nb = *ap.batch_index_dev + 1; *ap.batch_index_dev = nb >= ap.n_batches ? 0 : nb; }
            }
        } else {
            if (threadIdx.x < 4) { p.stats[p.lay.v_data + threadIdx.x] = counts[threadIdx.x]; p.stats[p.lay.v_model + threadIdx.x] = counts[4 + threadIdx.x]; }
            if (threadIdx.x == 0) {
                p.stats[p.lay.n_hidden_data] = p.c0; p.stats[p.lay.n_hidden_model] = p.c1;
                p.stats[p.lay.n_vis_data] = p.c2; p.stats[p.lay.n_vis_model] = p.c3;
            }
        }
    }
    if (ll) {
        // every CTA reads *sequence_dev at its start: the cursors are advanced by the last CTA to finish
        __syncthreads();
        if (threadIdx.x == 0) {
            if (timed_out_s) *p.ll.error_flag = 1;
            const unsigned int t = atomicAdd(p.ll.ticket, 1u);
            if (t == gridDim.x - 1) {
                *p.ll.ticket = 0u;
                if (ap.step_dev) *ap.step_dev += 1ull;
                if (ap.batch_index_dev) { const int nb = *ap.batch_index_dev + 1; *ap.batch_index_dev = nb >= ap.n_batches ? 0 : nb; }
            }
        }
    }
}

// ---- linear chain planes <-> the reference's dense (B,K,1,200) tensors ------------------------------------------------------
__global__ void chain_lin_to_dense_kernel(const uint32_t *__restrict__ bits, int Bn, int K, int Kpad, int hw, int VP, float *__restrict__ dense) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)Bn * K * FANTASY_LEN) return;
    const int j = (int)(gid % FANTASY_LEN);
    const int k = (int)((gid / FANTASY_LEN) % K);
    const int n = (int)(gid / ((long long)FANTASY_LEN * K));
    const uint32_t w = bits[((size_t)n * hw + j / VP) * Kpad + k];
    dense[gid] = (float)((w >> (j % VP)) & 1u);
}
__global__ void chain_dense_to_lin_kernel(const float *__restrict__ dense, int Bn, int K, int Kpad, int hw, int VP, uint32_t *__restrict__ bits) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)Bn * hw * Kpad) return;
    const int k = (int)(gid % Kpad);
    const int wq = (int)((gid / Kpad) % hw);
    const int n = (int)(gid / ((long long)Kpad * hw));
    uint32_t w = 0u;
    if (k < K)
        for (int b = 0; b < VP; ++b) {
            const int j = wq * VP + b;
            if (j < FANTASY_LEN && dense[((size_t)n * K + k) * FANTASY_LEN + j] != 0.f) w |= 1u << b;
        }
    bits[gid] = w;
}

// ---- one training step on the tcgen05 kernels -----------------------------------------------------------------------------------
//     h2v_lin -> v2h_train_tc<SAMPLE> [x cd_k, the last one with the model statistics] -> v2h_train_tc<P|Q> (minibatch) -> reduce_tc
// the same chain of programmatic dependents as the mma.sync path (train.cu), two partial buffers in the workspace.
static int train_phases_tc(const crbm_train_args *a, int phases, cudaStream_t s, const crbm_p2p_args *p2p, const PackJob *pack = nullptr) {
    const int K = a->K, M = a->M, P = a->pooling, Lh = a->L - M + 1;
    const TtGeom g = tt_geom(K, M, P);
    const int ds = a->doublestranded ? 1 : 0;
    const int Lv = FANTASY_LEN + M - 1;
    float *part_pos = reinterpret_cast<float *>(a->workspace);
    float *part_neg = part_pos + (size_t)g.n_chunks * g.G * g.part_floats;
    int rc;
    auto base_params = [&](TrainTcParams &t) {
        t.W = a->W; t.b = a->b; t.K = K; t.M = M; t.ds = ds;
        t.ksteps = g.ksteps; t.n_atoms = g.n_atoms; t.Kdim = g.Kdim; t.part_floats = g.part_floats;
        t.seed = a->seed; t.step_dev = a->step_dev;
        // shared memory: W (hi, lo) + ring of one-hot tiles + P/Q operand sets. Two sets when they fit beside a ring of at least
        // 6 tiles (a tile lives from its GEMM 1 until the GEMM 2 of the same unit, i.e. across the whole epilogue of the unit).
        const size_t cap = 227 * 1024;                              // opt-in maximum of dynamic shared memory per CTA on sm_100
        t.npsets = tt_smem_bytes(g.n_atoms, 4, 2) <= cap ? 2 : 1;
        t.xstages = TT_MAX_XSTAGES;
        while (t.xstages > 3 && tt_smem_bytes(g.n_atoms, t.xstages, t.npsets) > cap) --t.xstages;
        if (const char *e = getenv("CRBM_TT_NPSETS")) t.npsets = atoi(e) == 2 ? 2 : 1;          // experiments
        if (const char *e = getenv("CRBM_TT_XSTAGES")) t.xstages = atoi(e) < 2 ? 2 : atoi(e) > TT_MAX_XSTAGES ? TT_MAX_XSTAGES : atoi(e);
    };
    // ---- negative phase ----------------------------------------------------------------------------------------------------
    const int nb_neg = ceil_div(FANTASY_LEN, g.VP);
    const int units_neg = ceil_div(a->batchsize * nb_neg, 2);
    const int gx_neg = units_neg < g.G ? units_neg : g.G;
    if (phases & CRBM_PHASE_NEGATIVE) {
        const size_t n_u_v = (size_t)a->batchsize * Lv, n_u_h = (size_t)a->batchsize * K * (FANTASY_LEN / P);
        const int items_per_row = ceil_div(Lv, 32), lv_pad = 32 * items_per_row;
        float4 *act_part = reinterpret_cast<float4 *>(part_neg + (size_t)g.n_chunks * g.G * g.part_floats);
        CRBM_CHECK_ARG(a->workspace_bytes >= tt_workspace_floats(K, M, a->batchsize) * sizeof(float),
                       "crbm_train_accumulate: workspace too small for batchsize %d (crbm_train_workspace_bytes_for)", a->batchsize);
        const size_t h2c_smem = (size_t)TT_KC * M * 16 + (size_t)H2C_WARPS * H2C_LIST * sizeof(uint32_t);
        static size_t h2c_configured[64];
        int dev_id = 0;
        CRBM_CHECK_CUDA(cudaGetDevice(&dev_id));
        if (h2c_smem > h2c_configured[dev_id & 63]) {
            CRBM_CHECK_CUDA(cudaFuncSetAttribute(h2v_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h2c_smem));
            h2c_configured[dev_id & 63] = h2c_smem;
        }
        const int h2c_items = a->batchsize * items_per_row;
        int h2c_gx = ceil_div(h2c_items, H2C_WARPS);
        const int h2c_cap = (3 * num_sms()) / g.n_chunks > 0 ? (3 * num_sms()) / g.n_chunks : 1;
        if (h2c_gx > h2c_cap) h2c_gx = h2c_cap;
        // dense above 512 contraction columns (strands x Kpad): measured C4 (1024 columns) 886 -> 551 us per step; below, a particle is
        // a handful of MMAs and the sparse walk is faster (C3, 256 columns: 98 us per step against 107). CRBM_H2V_KERNEL=chunk|tc overrides.
        const bool h2v_dense = h2v_uses_tc((ds ? 2 : 1) * g.Kpad);
        unsigned char *a_img = reinterpret_cast<unsigned char *>(act_part + (size_t)g.n_chunks * a->batchsize * lv_pad);
        const int nkc = (ds ? 2 : 1) * g.n_chunks;
        if (h2v_dense) {
            // the weight image of the step's h2v GEMMs (W does not change between the sweeps of a step)
            static size_t ht_configured[64];
            if (ht_smem_bytes() > ht_configured[dev_id & 63]) {
                CRBM_CHECK_CUDA(cudaFuncSetAttribute(h2v_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ht_smem_bytes()));
                ht_configured[dev_id & 63] = ht_smem_bytes();
            }
            H2VPrepParams pp{a->W, K, g.Kpad, M, nkc, a_img};
            CRBM_CHECK_CUDA(launch_chain(h2v_prep_kernel, dim3(ceil_div(nkc * 1024, 256)), dim3(256), 0, s, pp));
            CRBM_CHECK_LAUNCH();
        }
        for (int it = 0; it < a->cd_k; ++it) {
            const bool last = it == a->cd_k - 1;
            if (h2v_dense) {
                H2VTcParams h{};
                h.a_img = a_img; h.bits = a->fantasy_bits; h.Bn = a->batchsize; h.K = K; h.Kpad = g.Kpad; h.hw = g.hw; h.VP = g.VP;
                h.M = M; h.nkc = nkc; h.c = a->c; h.v_packed = a->v_model; h.v_row_words = a->v_row_words; h.Lv = Lv;
                h.u_v = a->u_v ? a->u_v + it * n_u_v : nullptr;
                h.seed = a->seed; h.step_dev = a->step_dev;
                h.ctr_hi = a->step_dev ? 0u : (uint32_t)a->step; h.iter = (uint32_t)it;
                static const int cl_env = [] { const char *e = getenv("CRBM_H2V_CLUSTER"); const int v = e ? atoi(e) : 4; return v == 1 || v == 2 || v == 4 ? v : 4; }();
                h.cl = cl_env;
                static const int dbg_env = [] { const char *e = getenv("CRBM_H2V_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
                h.dbg = dbg_env;
                const int groups = ceil_div(a->batchsize, h.cl), max_clusters = num_sms() / h.cl;
                const int gx = (groups < max_clusters ? groups : max_clusters) * h.cl;
                CRBM_CHECK_CUDA(launch_h2v_tc(h, gx, s));
                CRBM_CHECK_LAUNCH();
            } else {
            H2VChunkParams h{};
            h.bits = a->fantasy_bits; h.Bn = a->batchsize; h.K = K; h.Kpad = g.Kpad; h.hw = g.hw; h.VP = g.VP; h.M = M; h.ds = ds;
            h.items_per_row = items_per_row; h.lv_pad = lv_pad; h.W = a->W; h.act_part = act_part;
            CRBM_CHECK_CUDA(launch_chain(h2v_chunk_kernel, dim3(h2c_gx, g.n_chunks), dim3(32 * H2C_WARPS), h2c_smem, s, h));
            CRBM_CHECK_LAUNCH();
            H2VFinishParams f{};
            f.act_part = act_part; f.Bn = a->batchsize; f.n_chunks = g.n_chunks; f.lv_pad = lv_pad; f.Lv = Lv;
            f.c = a->c; f.v_packed = a->v_model; f.v_row_words = a->v_row_words;
            f.u_v = a->u_v ? a->u_v + it * n_u_v : nullptr;
            f.seed = a->seed; f.step_dev = a->step_dev;
            f.ctr_hi = a->step_dev ? 0u : (uint32_t)a->step; f.iter = (uint32_t)it;
            CRBM_CHECK_CUDA(launch_chain(h2v_finish_kernel, dim3(ceil_div(a->batchsize * lv_pad, 256)), dim3(256), 0, s, f));
            CRBM_CHECK_LAUNCH();
            }

            TrainTcParams q{};
            base_params(q);
            q.packed = a->v_model; q.row_words = a->v_row_words; q.n_rows = a->batchsize; q.L = Lv; q.Lh = FANTASY_LEN;
            q.nblk32 = nb_neg; q.n_units = units_neg;
            q.partial = part_neg; q.bits = a->fantasy_bits; q.hw = g.hw; q.Kpad = g.Kpad;
            q.u_fwd = a->u_h ? a->u_h + it * n_u_h : nullptr;
            q.u_rc = a->u_hp ? a->u_hp + it * n_u_h : nullptr;
            q.ctr_hi = a->step_dev ? 0u : (uint32_t)a->step; q.iter = (uint32_t)it;
            if ((rc = dispatch_train_tc(P, last ? (F_STAT_P | F_SAMPLE) : F_SAMPLE, q, dim3(gx_neg, g.n_chunks), s))) return rc;
        }
    }
    // ---- positive phase + sparsity statistic ---------------------------------------------------------------------------------------
    const int nb_pos = ceil_div(Lh, g.VP);
    const int units_pos = ceil_div(a->n_data * nb_pos, 2);
    const int gx_pos = units_pos < g.G ? (units_pos > 0 ? units_pos : 1) : g.G;
    if ((phases & CRBM_PHASE_POSITIVE) && a->n_data > 0) {
        TrainTcParams t{};
        base_params(t);
        t.packed = a->data; t.row_words = a->data_row_words; t.n_rows = a->n_data; t.L = a->L; t.Lh = Lh;
        t.batch_index = a->batch_index_dev;
        t.nblk32 = nb_pos; t.n_units = units_pos;
        t.partial = part_pos;
        t.pdl_late = 1;
        if (pack) {             // host-fed step: see train_phases_impl
            if ((rc = launch_pack_job(*pack, a, s))) return rc;
            t.pdl_late = 0;
        }
        if ((rc = dispatch_train_tc(P, F_STAT_P | F_STAT_Q, t, dim3(gx_pos, g.n_chunks), s))) return rc;
    }
    // ---- reduction (+ update) -----------------------------------------------------------------------------------------------------
    if (phases & CRBM_PHASE_REDUCE) {
        TcReduceParams r{};
        r.pos = a->n_data > 0 ? part_pos : nullptr; r.neg = part_neg;
        r.G_pos = gx_pos; r.G_neg = gx_neg; r.part_floats = g.part_floats; r.Kdim = g.Kdim; r.n_chunks = g.n_chunks;
        r.K = K; r.M = M; r.ds = ds;
        r.stats = a->stats;
        fill_layout(K, M, &r.lay);
        r.c0 = (float)a->n_data * Lh; r.c1 = (float)a->batchsize * FANTASY_LEN;
        r.c2 = (float)a->n_data * a->L; r.c3 = (float)a->batchsize * Lv;
        r.cnt_rows[0] = a->data; r.cnt_row_words[0] = a->data_row_words; r.cnt_n[0] = a->n_data; r.cnt_L[0] = a->L;
        r.cnt_rows[1] = a->v_model; r.cnt_row_words[1] = a->v_row_words; r.cnt_n[1] = a->batchsize; r.cnt_L[1] = Lv;
        r.batch_index = a->batch_index_dev;
        const bool ll = p2p && p2p->peer_ll[0];
        r.fused = ((phases & CRBM_PHASE_APPLY) && (!p2p || ll) && a->n_data > 0 && fused_update_enabled()) ? 1 : 0;
        if (r.fused) fill_apply_params(a, &r.ap);
        if (r.fused && ll) {
            const size_t n_cta = (size_t)g.n_chunks * 4 * M + g.n_chunks + 1;
            CRBM_CHECK_ARG(p2p->ll_parity_stride >= n_cta * CRBM_MAX_PEERS * LL_ROWS * 32 && p2p->error_flag,
                           "crbm_train_step_p2p: receive buffer too small (crbm_p2p_ll_elems)");
            fill_ll_view(a, p2p, &r.ll);
        }
        if (p2p && (phases & CRBM_PHASE_APPLY) && !r.fused) { set_error("crbm_train_step_p2p: internal: the fused exchange is not available for this step"); return CRBM_ERR_UNSUPPORTED; }
        CRBM_CHECK_CUDA(launch_chain(reduce_tc_kernel, dim3(g.n_chunks * 4 * M + g.n_chunks + 1), dim3(512), 0, s, r));
        CRBM_CHECK_LAUNCH();
        if (r.fused) return CRBM_OK;
    }
    if (phases & CRBM_PHASE_APPLY) return crbm_train_apply(a, (void *)s);
    return CRBM_OK;
}

}  // namespace crbm
