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
// Warp roles (512 threads, one CTA per SM, persistent over particles blockIdx.x, blockIdx.x + gridDim.x, ...):
//   warp 0 streams A, warp 1 issues the MMAs, warp 2 owns the TMEM allocation, warps 4-7 epilogue (TMEM lane = thread),
//   warps 8-15 expand the chain bits into the B tiles. Three stages of (A hi, A lo, B); two accumulators of 224 columns.
#pragma once
#include "train_tc.cuh"

namespace crbm {

constexpr int HT_THREADS = 512;
constexpr int HT_STAGES = 3;
constexpr int HT_N = 224;                         // columns of a tile: 7 words x 32 bit positions
constexpr int HT_WORDS = 7;
constexpr int HT_A_BYTES = 128 * 128;             // one 64-kk atom of 128 rows
constexpr int HT_B_BYTES = HT_N * 128;
constexpr int HT_STAGE_BYTES = 2 * HT_A_BYTES + HT_B_BYTES;
constexpr int HT_CHUNK_STRIDE = 33;

struct H2VTcParams {
    const unsigned char *a_img;    // [nkc][2][HT_A_BYTES]: hi, lo atoms of the k-chunks (h2v_prep_kernel)
    const uint32_t *bits;          // chain planes [2][Bn][hw][Kpad]
    int Bn, K, Kpad, hw, VP, M, nkc;
    const float *c;
    uint32_t *v_packed;
    int v_row_words, Lv;
    const float *u_v;              // injected uniforms [n * Lv + i] or null
    uint64_t seed;
    uint32_t ctr_hi, iter;
    const unsigned long long *step_dev;
};

constexpr int HT_HBITS_BYTES = HT_WORDS * 256;    // raw chain words of one stage: 7 words x 64 motifs
__host__ __device__ inline size_t ht_smem_bytes() {
    return 1024 + (size_t)HT_STAGES * HT_STAGE_BYTES + 128 * HT_CHUNK_STRIDE * 4 + 4 * 256 * 4 + HT_STAGES * HT_HBITS_BYTES + 256;
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

__global__ void __launch_bounds__(HT_THREADS, 1) h2v_tc_kernel(const H2VTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char *stages = smem;
    float *chunk_s = reinterpret_cast<float *>(stages + (size_t)HT_STAGES * HT_STAGE_BYTES);        // [128][33]
    float *act_s = chunk_s + 128 * HT_CHUNK_STRIDE;                                                 // [4][256]
    uint32_t *hbits = reinterpret_cast<uint32_t *>(act_s + 4 * 256);                                // [stage][7][64] raw chain words
    uint64_t *full_a = reinterpret_cast<uint64_t *>(hbits + HT_STAGES * (HT_HBITS_BYTES / 4));
    uint64_t *full_b = full_a + HT_STAGES;
    uint64_t *empty = full_b + HT_STAGES;
    uint64_t *acc_full = empty + HT_STAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int G = gridDim.x, nkc = p.nkc;

    if (warp == 2) tmem_alloc(tmem_slot, 512u);
    if (tid == 0) {
        for (int i = 0; i < HT_STAGES; ++i) { mbar_init(full_a + i, 1); mbar_init(full_b + i, 8); mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_tiles = ((int)blockIdx.x < p.Bn) ? (p.Bn - 1 - (int)blockIdx.x) / G + 1 : 0;
    // programmatic dependent launch: the A image (h2v_prep_kernel) and the chain planes (the previous sweep) belong to predecessors
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ================================================= TMA stream: weight image + raw chain words ======================
        // (the chain words go through shared memory too: fetched by the producer warps themselves, the L2 latency of every
        // stage was exposed — 38 us per launch at K = 512 with the tensor pipe 24 % busy)
        if (lane == 0) {
            for (int it = 0; it < my_tiles * nkc; ++it) {
                const int stage = it % HT_STAGES, use = it / HT_STAGES, t = it / nkc, kc = it - t * nkc;
                const int n = (int)blockIdx.x + t * G;
                const int kk0 = 64 * kc, s = kk0 / p.Kpad, k0 = kk0 - s * p.Kpad;
                if (use > 0) mbar_wait(empty + stage, (uint32_t)((use - 1) & 1));
                unsigned char *dst = stages + (size_t)stage * HT_STAGE_BYTES;
                mbar_expect_tx(full_a + stage, 2 * HT_A_BYTES + (uint32_t)p.hw * 256u);
                tma_load_1d(dst, p.a_img + (size_t)kc * 2 * HT_A_BYTES, 2 * HT_A_BYTES, full_a + stage);
                const uint32_t *src = p.bits + (((size_t)s * p.Bn + n) * p.hw) * p.Kpad + k0;
                for (int w = 0; w < p.hw; ++w)
                    tma_load_1d(hbits + stage * (HT_HBITS_BYTES / 4) + w * 64, src + (size_t)w * p.Kpad, 256u, full_a + stage);
            }
        }
    } else if (warp == 1) {
        // ================================================= MMA issue (one thread) ==========================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(HT_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int it = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int buf = t & 1;
                if (t >= 2) mbar_wait(acc_empty + buf, (uint32_t)(((t >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t d = tmem_base + 256 * buf;
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int stage = it % HT_STAGES, use = it / HT_STAGES;
                    mbar_wait(full_a + stage, (uint32_t)(use & 1));
                    mbar_wait(full_b + stage, (uint32_t)(use & 1));
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stages + (size_t)stage * HT_STAGE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t bd = umma_desc_sw128(sa + 2 * HT_A_BYTES + 32 * ks);
                        umma_f16(d, umma_desc_sw128(sa + 32 * ks), bd, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                        umma_f16(d, umma_desc_sw128(sa + HT_A_BYTES + 32 * ks), bd, idesc, 1u);
                    }
                    umma_commit(empty + stage);
                }
                umma_commit(acc_full + buf);
            }
        }
    } else if (warp >= 8) {
        // ================================================= B: chain bits -> fp16 tiles ======================================
        // warp bw owns the 16-byte chunk bw of every row (motifs 8 bw .. 8 bw + 7 of the k-chunk); lane = bit position
        const int bw = warp - 8;
        for (int it = 0; it < my_tiles * nkc; ++it) {
            const int stage = it % HT_STAGES, use = it / HT_STAGES;
            mbar_wait(full_a + stage, (uint32_t)(use & 1));       // the stage is free (the loader saw `empty`) and its chain words are in
            const uint4 *hsrc = reinterpret_cast<const uint4 *>(hbits + stage * (HT_HBITS_BYTES / 4)) + 2 * bw;
            uint4 w0[HT_WORDS], w1[HT_WORDS];
#pragma unroll
            for (int w = 0; w < HT_WORDS; ++w) {
                if (w < p.hw) {
                    w0[w] = hsrc[w * 16];
                    w1[w] = hsrc[w * 16 + 1];
                } else {
                    w0[w] = make_uint4(0u, 0u, 0u, 0u);
                    w1[w] = w0[w];
                }
            }
            unsigned char *B = stages + (size_t)stage * HT_STAGE_BYTES + 2 * HT_A_BYTES;
            const bool live = lane < p.VP;
#pragma unroll
            for (int w = 0; w < HT_WORDS; ++w) {
                auto h = [&](uint32_t lo, uint32_t hi) -> uint32_t {        // fp16 pair ([bit of lo], [bit of hi])
                    return (((lo >> lane) & 1u) * 0x3C00u) | (((hi >> lane) & 1u) * 0x3C000000u);
                };
                uint4 q = make_uint4(h(w0[w].x, w0[w].y), h(w0[w].z, w0[w].w), h(w1[w].x, w1[w].y), h(w1[w].z, w1[w].w));
                if (!live) q = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4 *>(B + sw128_off(32 * w + lane, bw)) = q;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_b + stage);
        }
    } else if (warp >= 4) {
        // ================================================= epilogue =========================================================
        const int a = warp - 4, et = tid - 128;                  // base of this warp's diagonals; thread index among the 128
        const int r = 32 * a + lane;                             // TMEM lane = row (a, m = lane)
        const int span = 32 + p.M - 1;                           // visible positions one word of hidden positions reaches
        const float c0 = __ldg(p.c), c1 = __ldg(p.c + 1), c2 = __ldg(p.c + 2), c3 = __ldg(p.c + 3);
        for (int t = 0; t < my_tiles; ++t) {
            const int buf = t & 1, n = (int)blockIdx.x + t * G;
            for (int i = et; i < 4 * 256; i += 128) act_s[i] = 0.f;
            mbar_wait_backoff(acc_full + buf, (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * a) << 16) + 256 * buf;
            for (int w = 0; w < p.hw; ++w) {
                uint32_t va[16], vb[16];
                tmem_ld16(taddr + 32 * w, va);
                tmem_ld16(taddr + 32 * w + 16, vb);
                tmem_ld_wait();
                tmem_ld_fence16(va);
                tmem_ld_fence16(vb);
                if (w == p.hw - 1) {                             // the accumulator is in registers / shared memory from here on
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + buf);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");    // the previous word's diagonals have been read
                float *row = chunk_s + r * HT_CHUNK_STRIDE;
#pragma unroll
                for (int i = 0; i < 16; ++i) { row[i] = __uint_as_float(va[i]); row[16 + i] = __uint_as_float(vb[i]); }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // x[w VP + il, a] += sum_m D[32 a + m, il - m]; columns >= VP of a word are zero (their B rows are)
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    const int il = lane + 32 * rep;
                    if (il < span) {
                        float acc = 0.f;
                        const int m_lo = il > 31 ? il - 31 : 0, m_hi = il < p.M - 1 ? il : p.M - 1;
                        for (int m = m_lo; m <= m_hi; ++m) acc += chunk_s[(32 * a + m) * HT_CHUNK_STRIDE + il - m];
                        act_s[a * 256 + w * p.VP + il] += acc;
                    }
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
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
                if ((lane & 15) == 0 && i < 16 * ceil_div(p.Lv, 16)) p.v_packed[(size_t)n * p.v_row_words + FRONT_PAD_WORDS + (i >> 4)] = v;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");        // act_s is zeroed for the next particle after this point
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512u);
}

}  // namespace crbm
