// A5 + A6 of the chain sweeps as a dense GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulator in tensor memory).
//
// Reference semantics (secomo/convRBM.py): A5 `_topDownActivity` :277-292 (both strands, the reverse one through
// W[k, 3-a, M-1-m], :285), A6 `_topDownProbability` / `_topDownSample` :294-315 (4-way softmax over the bases, one multinomial draw
// per position) — the same quantities as h2v_bits_kernel / h2v_chunk_kernel + h2v_finish_kernel, which walk the SET bits of the
// chain. Those kernels are bound by the instructions of their compaction bookkeeping (ncu, profiles/r02_train_tc.md: 27-41 M warp
// instructions per launch at K = 512 for 0.4 M hits) and slow down as the chain fills; here the cost does not depend on the chain:
//
//   GEMM     D[r, v] = sum_kk  Wt[r, kk] . H[v, kk]^T        r = 32 a + m  (base a, motif offset m; 128 TMEM lanes),
//                                                            kk = s Kpad + k (strand s, motif k), v = virtual hidden position
//            A = Wt: W[k, a, m] for s = 0 and W[k, 3-a, M-1-m] for s = 1, fp16 hi + lo (22 significand bits), laid out once per step by
//                h2v_prep_kernel as a K-major SWIZZLE_128B image in global memory (L2 resident, shared by all CTAs) and streamed
//                through shared memory in chunks of 64 kk by the TMA engine (cp.async.bulk, mbarrier completion);
//            B = H: the chain planes uint32 H[s][n][w][Kpad] (bit b of word w = hidden position w VP + b) expanded to fp16 {0, 1}
//                in shared memory by eight producer warps (one 16-byte chunk = 8 motifs of one position per lane); never in HBM.
//   epilogue visible activity  x[i, a] = sum_m D[32 a + m, i - m]: the accumulator is read 32 columns at a time, staged in shared
//            memory, and warp a of the epilogue adds the diagonals of base a (bank-conflict free: a row's stride is 33 floats);
//            then c, the 4-way softmax, the multinomial draw and the packed 2-bit row of the sampled visible layer.
// One tile = one particle of the persistent chain (N = 7 words x 32 bit positions = 224 columns, K = strands x Kpad). At K = 512,
// M = 25, both strands: 128 MMAs of 128 x 224 x 16 per particle.
//
// Warp roles (768 threads, one CTA per SM, persistent over particles blockIdx.x, blockIdx.x + gridDim.x, ...):
//   warp 0 streams A, warp 1 issues the MMAs, warp 2 owns the TMEM allocation, warps 4-7 epilogue (TMEM lane = thread),
//   warp 3 streams the raw chain words, warps 8-23 expand them into the B tiles. Two accumulators of 224 columns.
#pragma once
#include "train_tc.cuh"

namespace crbm {

constexpr int HT_THREADS = 768;
constexpr int HT_BUILD_WARPS = 16;
constexpr int HT_ASTAGES = 4, HT_BSTAGES = 2, HT_HSTAGES = 3;
constexpr int HT_N = 224;                         // columns of a tile: 7 words x 32 bit positions
constexpr int HT_WORDS = 7;
constexpr int HT_A_BYTES = 128 * 128;             // one 64-kk atom of 128 rows
constexpr int HT_B_BYTES = HT_N * 128;
constexpr int HT_CHUNK_STRIDE = 64;            // staging row of the epilogue: 32 columns shifted by the row's motif offset m (<= 31)

struct H2VTcParams {
    const unsigned char *a_img;    // [nkc][2][HT_A_BYTES]: hi, lo atoms of the k-chunks (h2v_prep_kernel)
    const uint32_t *bits;          // chain planes [2][Bn][hw][Kpad]
    int Bn, K, Kpad, hw, VP, M, nkc;
    int cl;                        // CTAs per cluster (1, 2 or 4): the weight image is fetched once per cluster and multicast
    const float *c;
    uint32_t *v_packed;
    int v_row_words, Lv;
    const float *u_v;              // injected uniforms [n * Lv + i] or null
    uint64_t seed;
    uint32_t ctr_hi, iter;
    const unsigned long long *step_dev;
    int dbg;                       // CRBM_H2V_DEBUG=1: CTA 0 records clock64() at the stages of its first tile (crbm_debug_h2v_timeline)
};
__device__ long long g_h2v_timeline[16];
#define HT_MARK(i) do { if (p.dbg && blockIdx.x == 0) g_h2v_timeline[i] = clock64(); } while (0)

constexpr int HT_HBITS_BYTES = HT_WORDS * 256;    // raw chain words of one stage: 7 words x 64 motifs
__host__ __device__ inline size_t ht_smem_bytes() {
    return 1024 + (size_t)HT_ASTAGES * 2 * HT_A_BYTES + (size_t)HT_BSTAGES * HT_B_BYTES + 128 * HT_CHUNK_STRIDE * 4 + 4 * 256 * 4 +
           HT_HSTAGES * HT_HBITS_BYTES + 256;
}

// W -> the A image: thread = one 16-byte chunk (8 consecutive kk of one row)
struct H2VPrepParams {
    const float *W;
    int K, Kpad, M, nkc;
    unsigned char *a_img;
};
__global__ void __launch_bounds__(256) h2v_prep_kernel(const H2VPrepParams p) {
    pdl_wait();
    pdl_launch_dependents();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.nkc * 1024) return;
    const int kc = gid >> 10, r = (gid >> 3) & 127, c8 = gid & 7;
    const int a = r >> 5, m = r & 31;
    float w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int kk = 64 * kc + 8 * c8 + e;
        const int s = kk / p.Kpad, k = kk - s * p.Kpad;
        float v = 0.f;
        if (m < p.M && k < p.K)
            v = s == 0 ? __ldg(p.W + ((size_t)k * 4 + a) * p.M + m) : __ldg(p.W + ((size_t)k * 4 + (3 - a)) * p.M + (p.M - 1 - m));
        w[e] = v;
    }
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 h2 = __floats2half2_rn(w[2 * e], w[2 * e + 1]);
        const float2 back = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(w[2 * e] - back.x, w[2 * e + 1] - back.y);
        hh[e] = *reinterpret_cast<const uint32_t *>(&h2);
        ll[e] = *reinterpret_cast<const uint32_t *>(&l2);
    }
    unsigned char *dst = p.a_img + (size_t)kc * 2 * HT_A_BYTES + sw128_off(r, c8);
    *reinterpret_cast<uint4 *>(dst) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    *reinterpret_cast<uint4 *>(dst + HT_A_BYTES) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
}

// ---- cluster helpers: every CTA of a cluster fetches 1/cl of each weight chunk and multicasts it into the same shared-memory
// offset of all its peers (the chunk is the same for every particle: without this, 128 CTAs pull the same 32 KB from L2 at once,
// 64 MB per launch at K = 512, and the L2 -> SM path, not the tensor pipe, bounds the kernel)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d_multicast(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __launch_bounds__(HT_THREADS, 1) h2v_tc_kernel(const H2VTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    if (threadIdx.x == 0) HT_MARK(0);
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    // three rings with their own barriers, sized for what each has to hide (ncu on the first version, one ring of three (A, B) stages:
    // 1.25 us per k-chunk against 0.48 us of MMAs — the 32 KB weight chunk has ~1.2 us of L2 latency + transfer, and the chain bits,
    // the expansion and the MMAs of a stage waited for it in turn):
    //   A   weight chunks (hi, lo), 4 x 32 KB, filled by the TMA engine as soon as the MMAs of four chunks ago have completed
    //   H   raw chain words, 3 x 1.8 KB, fetched ahead (they do not depend on anything in this kernel)
    //   B   expanded fp16 tiles, 2 x 28 KB: the expansion of chunk i+1 overlaps the MMAs of chunk i
    unsigned char *ring_a = smem;
    unsigned char *ring_b = ring_a + (size_t)HT_ASTAGES * 2 * HT_A_BYTES;
    float *chunk_s = reinterpret_cast<float *>(ring_b + (size_t)HT_BSTAGES * HT_B_BYTES);           // [128][33]
    float *act_s = chunk_s + 128 * HT_CHUNK_STRIDE;                                                 // [4][256]
    uint32_t *hbits = reinterpret_cast<uint32_t *>(act_s + 4 * 256);                                // [HT_HSTAGES][7][64]
    uint64_t *full_a = reinterpret_cast<uint64_t *>(hbits + HT_HSTAGES * (HT_HBITS_BYTES / 4));
    uint64_t *empty_a = full_a + HT_ASTAGES;
    uint64_t *full_h = empty_a + HT_ASTAGES;
    uint64_t *empty_h = full_h + HT_HSTAGES;
    uint64_t *full_b = empty_h + HT_HSTAGES;
    uint64_t *empty_b = full_b + HT_BSTAGES;
    uint64_t *acc_full = empty_b + HT_BSTAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int nkc = p.nkc, cl = p.cl;
    const int rank = cl > 1 ? (int)cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << cl) - 1u);
    // tiles: cluster `cid` works on the particle groups cid, cid + n_clusters, ...; CTA `rank` of the cluster on particle group * cl + rank.
    // All CTAs of a cluster run the same number of stages (they feed each other); a CTA whose particle does not exist works on the
    // last one and writes nothing.
    const int n_clusters = gridDim.x / cl, cid = blockIdx.x / cl, n_groups = ceil_div(p.Bn, cl);

    if (warp == 2) tmem_alloc(tmem_slot, 512u);
    for (int i = tid; i < 128 * HT_CHUNK_STRIDE; i += HT_THREADS) chunk_s[i] = 0.f;     // the pads of the skewed rows stay zero
    if (tid == 0) {
        for (int i = 0; i < HT_ASTAGES; ++i) { mbar_init(full_a + i, 1); mbar_init(empty_a + i, cl); }
        for (int i = 0; i < HT_HSTAGES; ++i) { mbar_init(full_h + i, 1); mbar_init(empty_h + i, HT_BUILD_WARPS); }
        for (int i = 0; i < HT_BSTAGES; ++i) { mbar_init(full_b + i, HT_BUILD_WARPS); mbar_init(empty_b + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (cl > 1) cluster_sync_all();                 // the peers' barriers exist before anything is sent to them
    if (threadIdx.x == 0) HT_MARK(1);
    const uint32_t tmem_base = *tmem_slot;
    const int my_tiles = cid < n_groups ? (n_groups - 1 - cid) / n_clusters + 1 : 0;
    const int total = my_tiles * nkc;
    auto particle = [&](int t) { return (cid + t * n_clusters) * cl + rank; };
    // programmatic dependent launch: the A image (h2v_prep_kernel) and the chain planes (the previous sweep) belong to predecessors
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0) HT_MARK(2);

    if (warp == 0) {
        // ================================================= TMA stream of the weight image ====================================
        if (lane == 0) {
            for (int it = 0; it < total; ++it) {
                const int stage = it % HT_ASTAGES, use = it / HT_ASTAGES, kc = it % nkc;
                if (use > 0) mbar_wait(empty_a + stage, (uint32_t)((use - 1) & 1));   // every CTA of the cluster has released the stage
                unsigned char *dst = ring_a + (size_t)stage * 2 * HT_A_BYTES;
                mbar_expect_tx(full_a + stage, 2 * HT_A_BYTES);
                if (cl == 1) tma_load_1d(dst, p.a_img + (size_t)kc * 2 * HT_A_BYTES, 2 * HT_A_BYTES, full_a + stage);
                else {
                    const uint32_t slice = 2 * HT_A_BYTES / (uint32_t)cl;
                    tma_load_1d_multicast(dst + rank * slice, p.a_img + (size_t)kc * 2 * HT_A_BYTES + rank * slice, slice, full_a + stage, cmask);
                }
            }
        }
    } else if (warp == 3) {
        // ================================================= TMA stream of the raw chain words ================================
        if (lane == 0) {
            for (int it = 0; it < total; ++it) {
                const int stage = it % HT_HSTAGES, use = it / HT_HSTAGES, t = it / nkc, kc = it - t * nkc;
                const int n = min(particle(t), p.Bn - 1);
                const int kk0 = 64 * kc, s = kk0 / p.Kpad, k0 = kk0 - s * p.Kpad;
                if (use > 0) mbar_wait(empty_h + stage, (uint32_t)((use - 1) & 1));
                mbar_expect_tx(full_h + stage, (uint32_t)p.hw * 256u);
                const uint32_t *src = p.bits + (((size_t)s * p.Bn + n) * p.hw) * p.Kpad + k0;
                for (int w = 0; w < p.hw; ++w)
                    tma_load_1d(hbits + stage * (HT_HBITS_BYTES / 4) + w * 64, src + (size_t)w * p.Kpad, 256u, full_h + stage);
            }
        }
    } else if (warp == 1) {
        // ================================================= MMA issue (one thread) ==========================================
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(HT_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int it = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int buf = t & 1;
                if (t >= 2) mbar_wait(acc_empty + buf, (uint32_t)(((t >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t d = tmem_base + 256 * buf;
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int sa_i = it % HT_ASTAGES, sb_i = it % HT_BSTAGES;
                    mbar_wait(full_a + sa_i, (uint32_t)((it / HT_ASTAGES) & 1));
                    if (it == 0) HT_MARK(3);
                    if (it == 8) HT_MARK(10);
                    mbar_wait(full_b + sb_i, (uint32_t)((it / HT_BSTAGES) & 1));
                    if (it == 0) HT_MARK(4);
                    if (it == 8) HT_MARK(11);
                    tc_fence_after();
                    const uint32_t sa = umma_desc_lo(smem_u32(ring_a + (size_t)sa_i * 2 * HT_A_BYTES)), sb = umma_desc_lo(smem_u32(ring_b + (size_t)sb_i * HT_B_BYTES));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        umma_f16_lo(d, sa + 2 * ks, sb + 2 * ks, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                        umma_f16_lo(d, sa + (HT_A_BYTES >> 4) + 2 * ks, sb + 2 * ks, idesc, 1u);
                    }
                    umma_commit(empty_b + sb_i);
                    if (cl == 1) umma_commit(empty_a + sa_i);
                    else umma_commit_multicast(empty_a + sa_i, cmask);
                }
                umma_commit(acc_full + buf);
                if (t == 0) HT_MARK(5);
            }
        }
    } else if (warp >= 8) {
        // ================================================= B: chain bits -> fp16 tiles ======================================
        // warp (bw, half) owns the 16-byte chunk bw of the rows of every second word (motifs 8 bw .. 8 bw + 7 of the k-chunk; words
        // half, half + 2, ...); lane = bit position.
        const int bw = (warp - 8) & 7, half = (warp - 8) >> 3;
        const uint32_t sel = 1u << lane;
        const bool live = lane < p.VP;
        for (int it = 0; it < total; ++it) {
            const int sh_i = it % HT_HSTAGES, sb_i = it % HT_BSTAGES;
            mbar_wait(full_h + sh_i, (uint32_t)((it / HT_HSTAGES) & 1));
            const uint4 *hsrc = reinterpret_cast<const uint4 *>(hbits + sh_i * (HT_HBITS_BYTES / 4)) + 2 * bw;
            uint4 w0[(HT_WORDS + 1) / 2], w1[(HT_WORDS + 1) / 2];
#pragma unroll
            for (int wi = 0; wi < (HT_WORDS + 1) / 2; ++wi) {
                const int w = 2 * wi + half;
                w0[wi] = w < p.hw ? hsrc[w * 16] : make_uint4(0u, 0u, 0u, 0u);
                w1[wi] = w < p.hw ? hsrc[w * 16 + 1] : make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_h + sh_i);           // the words are in registers
            if (it >= HT_BSTAGES) mbar_wait(empty_b + sb_i, (uint32_t)(((it / HT_BSTAGES) - 1) & 1));
            unsigned char *B = ring_b + (size_t)sb_i * HT_B_BYTES;
#pragma unroll
            for (int wi = 0; wi < (HT_WORDS + 1) / 2; ++wi) {
                const int w = 2 * wi + half;
                if (w < HT_WORDS) {
                    auto h = [&](uint32_t lo, uint32_t hi) -> uint32_t {        // fp16 pair ([bit of lo], [bit of hi])
                        return ((lo & sel) ? 0x3C00u : 0u) | ((hi & sel) ? 0x3C000000u : 0u);
                    };
                    uint4 q = make_uint4(h(w0[wi].x, w0[wi].y), h(w0[wi].z, w0[wi].w), h(w1[wi].x, w1[wi].y), h(w1[wi].z, w1[wi].w));
                    if (!live) q = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4 *>(B + sw128_off(32 * w + lane, bw)) = q;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_b + sb_i);
        }
    } else if (warp >= 4) {
        // ================================================= epilogue =========================================================
        const int a = warp - 4, et = tid - 128;                  // base of this warp's diagonals; thread index among the 128
        const int r = 32 * a + lane;                             // TMEM lane = row (a, m = lane)
        const int span = 32 + p.M - 1;                           // visible positions one word of hidden positions reaches
        const float c0 = __ldg(p.c), c1 = __ldg(p.c + 1), c2 = __ldg(p.c + 2), c3 = __ldg(p.c + 3);
        for (int t = 0; t < my_tiles; ++t) {
            const int buf = t & 1, n_raw = particle(t), n = min(n_raw, p.Bn - 1);
            const bool real = n_raw < p.Bn;
            for (int i = et; i < 4 * 256; i += 128) act_s[i] = 0.f;
            mbar_wait_backoff(acc_full + buf, (uint32_t)((t >> 1) & 1));
            if (t == 0 && tid == 128) HT_MARK(6);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * a) << 16) + 256 * buf;
            for (int w = 0; w < p.hw; ++w) {
                uint32_t va[16], vb[16];
                tmem_ld16(taddr + 32 * w, va);
                tmem_ld16(taddr + 32 * w + 16, vb);
                tmem_ld_wait();
                tmem_ld_fence16(va);
                tmem_ld_fence16(vb);
                if (t == 0 && tid == 128 && w == 1) HT_MARK(12);
                if (w == p.hw - 1) {                             // the accumulator is in registers / shared memory from here on
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + buf);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");    // the previous word's diagonals have been read
                // skewed staging: row (a, m) is written m columns to the right, so that visible position il of the word is COLUMN il of
                // every row and the diagonal sum is a plain column sum with uniform bounds. (The first version kept the rows aligned
                // and predicated every term on 0 <= il - m < 32: 64 predicated loads per thread and word, serialised through the
                // predicate registers — 1.1 us per word, 8.5 us of exposed epilogue per particle.) Writes: lane m -> stride 65 floats;
                // reads: lane il -> stride 1: both conflict free.
                float *row = chunk_s + r * HT_CHUNK_STRIDE + lane;
#pragma unroll
                for (int i = 0; i < 16; ++i) { row[i] = __uint_as_float(va[i]); row[16 + i] = __uint_as_float(vb[i]); }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (t == 0 && tid == 128 && w == 1) HT_MARK(13);
                // x[w VP + il, a] += sum_m D[32 a + m, il - m]; columns >= VP of a word are zero (their B rows are)
                {
                    const float *src = chunk_s + (32 * a) * HT_CHUNK_STRIDE + lane;
                    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
                    int m = 0;
                    for (; m + 2 <= p.M; m += 2) {
                        a0 += src[m * HT_CHUNK_STRIDE];
                        b0 += src[m * HT_CHUNK_STRIDE + 32];
                        a1 += src[(m + 1) * HT_CHUNK_STRIDE];
                        b1 += src[(m + 1) * HT_CHUNK_STRIDE + 32];
                    }
                    if (m < p.M) { a0 += src[m * HT_CHUNK_STRIDE]; b0 += src[m * HT_CHUNK_STRIDE + 32]; }
                    act_s[a * 256 + w * p.VP + lane] += a0 + a1;
                    if (lane + 32 < span) act_s[a * 256 + w * p.VP + lane + 32] += b0 + b1;
                }
                if (t == 0 && tid == 128 && w == 1) HT_MARK(14);
                if (t == 0 && tid == 128 && w == 0) HT_MARK(15);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (t == 0 && tid == 128) HT_MARK(7);
            // A6: c, softmax over the four bases, one draw per position, packed 2-bit codes (16 positions per word)
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
                const int i = et + 128 * rep;
                int pick = 0;
                if (i < p.Lv) {
                    const float v0 = act_s[i] + c0, v1 = act_s[256 + i] + c1, v2 = act_s[512 + i] + c2, v3 = act_s[768 + i] + c3;
                    const float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
                    const float e0 = __expf(v0 - mx), e1 = __expf(v1 - mx), e2 = __expf(v2 - mx), e3 = __expf(v3 - mx);
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
                if (real && (lane & 15) == 0 && i < 16 * ceil_div(p.Lv, 16)) p.v_packed[(size_t)n * p.v_row_words + FRONT_PAD_WORDS + (i >> 4)] = v;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");        // act_s is zeroed for the next particle after this point
            if (t == 0 && tid == 128) HT_MARK(8);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512u);
    if (cl > 1) cluster_sync_all();                 // no CTA leaves while a peer may still signal its barriers
    if (threadIdx.x == 0) HT_MARK(9);
}

// programmatic dependent launch + thread-block cluster
static inline cudaError_t launch_h2v_tc(const H2VTcParams &p, int grid, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(HT_THREADS); cfg.dynamicSmemBytes = ht_smem_bytes(); cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (p.cl > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = p.cl; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, h2v_tc_kernel, p);
}

}  // namespace crbm
