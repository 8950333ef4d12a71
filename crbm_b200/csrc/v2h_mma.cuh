// Fused visible->hidden + statistics kernel for pooling == 1, with both contractions on the tensor pipe.
//
// Why this kernel exists (ncu, profiles/r01_v2h_fused_cudacore.md): in the CUDA-core kernel of v2h.cuh the
// statistics outer product acc[a][m] += p[s-m] needs 4*M register accumulators per motif behind a 4-way switch on the
// base; fully unrolled that is ~7k SASS instructions, the warps stall on instruction fetch ("no_instruction" 2.3
// per issue, 28% issue utilisation) and 255 registers/thread leave 8 warps per SM. Here the two contractions
//     x[k, j]          = sum_c  W[k, c] * X[j, c]        (c = 4*m + a;  X[j, c] = [base(j+m) == a], one-hot)
//     VH[k, c]        += sum_j  p[k, j] * X[j, c]        (+ reverse-strand fold, + Q for the sparsity term)
// are issued as mma.sync.m16n8k16 (bf16 inputs, fp32 accumulate):
//   * X is one-hot, hence exact in bf16; it is never materialised — every lane builds its B fragments straight from
//     the packed 2-bit bases with a few shifts and compares;
//   * W is split once per CTA into three bf16 terms (hi+mid+lo = 24 mantissa bits) so that x is fp32-accurate, and
//     kept in shared memory in fragment order (one conflict-free LDS.128 per A fragment), staged by cp.async.bulk;
//   * the accumulator fragment of the first MMA (16 motifs x 8 positions) has exactly the register layout of the A
//     fragment of the second one, so probabilities go from the sigmoid epilogue into the statistics MMA without
//     touching shared memory; p is split into two bf16 terms (error <= 2^-17 per term, unbiased);
//   * the VH accumulators live in MMA fragments (16*KS registers per statistic), there is no branch in the loop and
//     the whole kernel is ~1.5k instructions.
// Work decomposition is the one of v2h.cuh: grid (k-tiles of 32 motifs, CTAs), every warp walks blocks of 16 hidden
// positions of one sequence and keeps its accumulators for the whole kernel; one tree reduction per CTA at the end.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "v2h.cuh"

namespace crbm {

constexpr int MTILE = 16;                                       // motifs per CTA of the tensor-pipe kernels (one MMA m-tile)
constexpr int NSPLIT = 2;                                       // W = hi + lo in fp16 (11 + 11 significand bits; three bf16 terms in the first version)

template <int KS> struct MmaShape {
    static constexpr int NT2 = 2 * KS;                          // 8-column tiles over c = 4*m + a  (16*KS columns)
    static constexpr int FRAGS = 2 * NSPLIT * KS;               // [strand][split][ks]
    static constexpr size_t wimg_bytes = (size_t)FRAGS * 512;   // 32 lanes x 16 bytes per fragment
    static constexpr size_t wtile_bytes = (size_t)MTILE * 4 * CRBM_MAX_MOTIF_LEN * sizeof(float);
    static constexpr size_t smem_bytes = wimg_bytes + wtile_bytes + 16;
};
template <int KS, int FLAGS> constexpr size_t mma_smem_bytes() {
    // W fragments + staging tile, or the scratch of the final 8-warp tree reduction, whichever is larger
    const size_t a = MmaShape<KS>::smem_bytes;
    const size_t nacc = ((FLAGS & F_STAT_P) ? 8 * KS : 0) + ((FLAGS & F_STAT_Q) ? 8 * KS : 0) + 6;
    const size_t b = 4 * nacc * 32 * sizeof(float);
    return a > b ? a : b;
}

template <int KS, int FLAGS> struct MmaAcc {
    static constexpr int NP = (FLAGS & F_STAT_P) ? 8 * KS : 0;       // registers per thread
    static constexpr int NQ = (FLAGS & F_STAT_Q) ? 8 * KS : 0;
    static constexpr int NACC = NP + NQ + 6;                         // + sumP, sumPr, sumQ for the thread's 2 motif rows
};

__device__ __forceinline__ void hmma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void hmma_bf16(float (&d)[4], const uint4 &a, uint32_t b0, uint32_t b1) {
    hmma_bf16(d, a.x, a.y, a.z, a.w, b0, b1);
}
__device__ __forceinline__ void hmma_f16(float (&d)[4], const uint4 &a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ void hmma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    hmma_bf16(d, a[0], a[1], a[2], a[3], b0, b1);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
// (v0, v1) -> hi = bf16x2(v0, v1), mid = bf16x2 of the rounding residuals
__device__ __forceinline__ void split2(float v0, float v1, uint32_t &hi, uint32_t &mid) {
    hi = pack_bf16x2(v0, v1);
    mid = pack_bf16x2(v0 - __uint_as_float(hi << 16), v1 - __uint_as_float(hi & 0xffff0000u));
}

// W tile (16 motifs x 4 x M, fp32, staged by the TMA engine) -> two fp16 images (hi, lo) in MMA A-fragment order, both strands.
// Element (row r, column c) of the ks-th fragment is Wx[r][16*ks + c]; forward strand Wx[k][4m+a] = W[k][a][m],
// reverse strand Wx[k][4m+a] = W[k][3-a][M-1-m] (convRBM.py:241 with filter_flip=True); zero beyond M.
template <int KS>
__device__ __forceinline__ void build_w_fragments(uint32_t *wimg, float *wtile, uint64_t *bar,
                                                  const float *__restrict__ W, int K, int M, int mtile) {
    const int k0 = mtile * MTILE;
    const int nk = min(MTILE, K - k0);
    const uint32_t bytes = (uint32_t)nk * 4u * (uint32_t)M * 4u;
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        tma_load_1d(wtile, W + (size_t)k0 * 4 * M, bytes, bar);
    }
    mbar_wait(bar, 0);
    // one work unit = one 32-bit register of one fragment: (strand, ks, lane, reg)
    const int units = 2 * KS * 32 * 4;
    for (int u = threadIdx.x; u < units; u += blockDim.x) {
        const int reg = u & 3, lane = (u >> 2) & 31, rest = u >> 7;
        const int ks = rest % KS, strand = rest / KS;
        const int g = lane >> 2, q = lane & 3;
        const int row = g + 8 * (reg & 1);
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = 16 * ks + 2 * q + e + 8 * (reg >> 1);
            const int m = c >> 2, a = c & 3;
            float w = 0.f;
            if (m < M && row < nk) w = strand == 0 ? wtile[(row * 4 + a) * M + m] : wtile[(row * 4 + 3 - a) * M + (M - 1 - m)];
            v[e] = w;
        }
        // W = hi + lo, both fp16: |W - hi - lo| <= 2^-22 |W| (3e-8 absolute below the fp16 normal range), as in scan_tc.cu
        const uint32_t hi = pack_half2(v[0], v[1]);
        const __half2 hh = *reinterpret_cast<const __half2 *>(&hi);
        const uint32_t lo = pack_half2(v[0] - __low2float(hh), v[1] - __high2float(hh));
        const int f0 = (strand * NSPLIT + 0) * KS + ks;              // fragment index [strand][split][ks]
        wimg[(f0 * 32 + lane) * 4 + reg] = hi;
        wimg[((f0 + KS) * 32 + lane) * 4 + reg] = lo;
    }
    __syncthreads();
}

// ---- one-hot B fragments from packed bases, bit-parallel ---------------------------------------------------------------
// A B-fragment register is a bf16x2 whose halves are 1.0 (0x3F80) or 0. For a 16-base window word w the comparisons are
// done for all 16 bases at once (a few LOP3), the two match bits of a register are brought to bits 0 and 16 of an
// 8-base "group" word, and a register then costs SHF + LOP3 + IMAD:  ((G >> 2i) & 0x00010001) * 0x3F80.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t onehot_reg(uint32_t group, int i /*0,2,4,6: first base inside its 8-base group*/) {
    return ((group >> (2 * i)) & 0x00010001u) * 0x3F80u;
}
// same with fp16 ones (0x3C00): B operand of the up-pass, whose A operand (W) is fp16
__device__ __forceinline__ uint32_t onehot_reg_h(uint32_t group, int i) {
    return ((group >> (2 * i)) & 0x00010001u) * 0x3C00u;
}
// Up-pass operand: for base b at window position i, low half = [b == 2*qodd], high half = [b == 2*qodd + 1].
// Fills grp[2w], grp[2w+1] for the NW window words.
template <int NW>
__device__ __forceinline__ void upass_groups(const uint32_t (&w)[NW], uint32_t qodd_mask /*qodd ? 0 : 0x55555555*/,
                                             uint32_t (&grp)[2 * NW]) {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const uint32_t hi = (w[i] >> 1) & 0x55555555u, lo = w[i] & 0x55555555u;
        const uint32_t m = hi ^ qodd_mask;                   // bit 2i: (base >> 1) == qodd
        const uint32_t a = m & ~lo, b = m & lo;              // base & 1 == 0 / == 1
        grp[2 * i] = prmt(a, b, 0x5410u);                    // bases 0-7: a in the low 16 bits, b in the high 16 bits
        grp[2 * i + 1] = prmt(a, b, 0x7632u);                // bases 8-15
    }
}
// Statistics operand: for bases (i, i+1) against one code: low half = [base_i == code], high half = [base_{i+1} == code].
template <int NW>
__device__ __forceinline__ void pair_groups(const uint32_t (&w)[NW], uint32_t code_rep /*code * 0x55555555*/,
                                            uint32_t (&grp)[2 * NW]) {
    uint32_t e[NW + 1];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const uint32_t eq = ~(w[i] ^ code_rep);
        e[i] = eq & (eq >> 1) & 0x55555555u;                 // bit 2i: base i == code
    }
    e[NW] = 0u;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const uint32_t nxt = __funnelshift_r(e[i], e[i + 1], 2);   // bit 2i: base i+1 == code
        grp[2 * i] = prmt(e[i], nxt, 0x5410u);
        grp[2 * i + 1] = prmt(e[i], nxt, 0x7632u);
    }
}
// out-of-line copy of the generator for call sites inside heavily unrolled code
static __device__ __noinline__ uint4 philox4x32_call(uint4 ctr, uint2 key) { return philox4x32(ctr, key); }

// 1 / (1 + exp(-x)) with the SFU approximations only (ex2.approx: 2 ulp, rcp.approx: 1 ulp; saturates cleanly to 0 / 1)
__device__ __forceinline__ float fast_sigmoid(float x) {
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
    return r;
}

// grid (m-tiles of 16 motifs, CTAs), 256 threads, two CTAs per SM: every warp owns the CTA's m-tile for a stream of
// 16-position blocks. (ncu on the first version with 32-motif warps: 254 registers, 8 warps/SM, one wave of 4.05 blocks
// per warp on average but 5 at most; 16-motif warps need <= 128 registers, and m-tiles beyond K are not launched.)
// The body is shared by the single-role kernel and by the two-role kernel below: `mtile` is the CTA's m-tile, `by` / `gy` its index
// and the number of CTAs among those that walk the same item stream (blockIdx.y / gridDim.y in the single-role kernel).
template <int KS, int FLAGS, bool POOL>
__device__ __forceinline__ void v2h_mma_body(const V2HParams &p, unsigned char *smem, const int mtile, const int by, const int gy) {
    using Sh = MmaShape<KS>;
    using Ac = MmaAcc<KS, FLAGS>;
    constexpr int NT2 = Sh::NT2;
    constexpr bool STAT_P = FLAGS & F_STAT_P, STAT_Q = FLAGS & F_STAT_Q, SAMPLE = FLAGS & F_SAMPLE;
    uint32_t *wimg = reinterpret_cast<uint32_t *>(smem);
    float *wtile = reinterpret_cast<float *>(smem + Sh::wimg_bytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + Sh::wimg_bytes + Sh::wtile_bytes);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    build_w_fragments<KS>(wimg, wtile, bar, p.W, p.K, p.M, mtile);
    const uint4 *wfrag = reinterpret_cast<const uint4 *>(wimg) + lane;

    // the thread's two motif rows: h = 0, 1  ->  k = 16*mtile + 8*h + g
    float bias[2];
    bool kval[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = mtile * MTILE + 8 * h + g;
        kval[h] = k < p.K;
        bias[h] = kval[h] ? __ldg(p.b + k) : 0.f;
    }
    const bool all_k = (mtile + 1) * MTILE <= p.K;

    float accp[NT2][4], accq[NT2][4];
#pragma unroll
    for (int t = 0; t < NT2; ++t)
#pragma unroll
        for (int r = 0; r < 4; ++r) { accp[t][r] = 0.f; accq[t][r] = 0.f; }
    float sum_p[2] = {0.f, 0.f}, sum_pr[2] = {0.f, 0.f}, sum_q[2] = {0.f, 0.f};

    // lane-dependent parts of the base offsets (see the derivations in the header comment of each stage)
    const int lo1 = g + (q >> 1);                       // up-pass B: base j0 + 8nt + g + 4ks + q/2 (+2)
    const int lo2 = 2 * q + (g >> 2);                   // forward statistics B: base j0 + 2q + m_c (+1) (+8), m_c = 2nt2 + g/4
    const int mshift = p.M - 4 * KS;                    // in [-3, 0]
    const int sh3 = 16 + 2 * q - (g >> 2) + 1 + mshift; // reverse statistics: window starts at base j0 - 16 + sh3
    const uint32_t qmask = (q & 1) ? 0u : 0x55555555u;
    const uint32_t code_f = (uint32_t)(g & 3) * 0x55555555u, code_r = (uint32_t)(3 - (g & 3)) * 0x55555555u;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const uint32_t ctr_hi = p.ctr_hi + (p.step_dev ? (uint32_t)(*p.step_dev) : 0u);
    const int ktiles = (p.K + KTILE - 1) / KTILE;
    const int njb = (p.Lh + 15) >> 4;

    // pooling > 1 (POOL): blocks of bv = (16 / P) * P positions, so that pool groups never straddle blocks; the MMA tile still
    // spans 16 positions from j0 = blk * bv and the tail is masked. The thread's four positions 8nt + 2q + e of a motif row
    // belong to the groups gidx[2nt + e]; group sums are quad reductions (the four lanes of a quad hold one motif row).
    const float fpool = (float)p.pooling;
    const int ng = POOL ? p.bv / p.pooling : 0;
    const int groups_per_seq = p.Lh / p.pooling;
    int gidx[4] = {0, 0, 0, 0};
    PoolMasks mk{0u, 0u, 0u};
    if (POOL) {
#pragma unroll
        for (int i = 0; i < 4; ++i) gidx[i] = (8 * (i >> 1) + 2 * q + (i & 1)) / p.pooling;
        mk = make_pool_masks(p.pooling);
    }
    const uint32_t *packed = p.packed + (p.batch_index ? (size_t)(*p.batch_index) * p.n_rows * p.row_words : 0);
    // Programmatic dependent launch: everything above reads only W, b and the step / minibatch cursors, which the update
    // kernel of the previous step wrote and which were complete before the predecessor grid could let this one start
    // (every kernel of the chain triggers its dependents only after its own wait). The sequence rows and the bit planes
    // belong to the predecessor (h2v) for the chain sweeps; the positive phase depends on nothing in flight and only
    // waits before it exits, so that "this grid has completed" still implies "everything before it has".
    if (!p.pdl_late) pdl_wait();
    pdl_launch_dependents();
    for (int item = by * nwarps + warp; item < p.n_items; item += gy * nwarps) {
        const int n = item / p.nblk, jb = item - n * p.nblk, j0 = POOL ? jb * p.bv : jb * 16;
        const uint32_t *row = packed + (size_t)n * p.row_words + (j0 >> 4);     // word holding bases j0-16 .. j0-1 (j0 % 16 == 0)
        uint32_t wd[5];                                                        // wd[i]: bases j0-16+16i .. j0-1+16i
        if (POOL) {
            uint32_t raw[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) raw[i] = __ldg(row + i);
            const int sh = 2 * (j0 & 15);                                      // warp-uniform
#pragma unroll
            for (int i = 0; i < 5; ++i) wd[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
        } else {
#pragma unroll
            for (int i = 0; i < 5; ++i) wd[i] = __ldg(row + i);
        }

        // ---- x = W (x) X on the tensor pipe, both strands ------------------------------------------------------
        float d1[2][2][4];                                     // [strand][nt][reg]
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r) d1[s][nt][r] = 0.f;
        {
            uint32_t v[3], vg[6];
#pragma unroll
            for (int i = 0; i < 3; ++i) v[i] = __funnelshift_r(wd[1 + i], wd[2 + i], 2 * lo1);
            upass_groups<3>(v, qmask, vg);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t bx[2][2];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        const int off = 8 * nt + 4 * ks + 2 * hb;
                        bx[nt][hb] = onehot_reg_h(vg[off >> 3], off & 7);
                    }
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (s == 1 && !p.ds) continue;
#pragma unroll
                    for (int sp = 0; sp < NSPLIT; ++sp) {
                        const uint4 a = wfrag[((s * NSPLIT + sp) * KS + ks) * 32];
                        hmma_f16(d1[s][0], a, bx[0][0], bx[0][1]);
                        hmma_f16(d1[s][1], a, bx[1][0], bx[1][1]);
                    }
                }
            }
        }
        // ---- sigmoid epilogue, sampling -------------------------------------------------------------------------
        // d1[s][nt][r]: motif row h = r>>1, position j0 + 8*nt + 2*q + (r&1)
        const int jrem = p.Lh - j0 - 2 * q;                   // position valid iff 8*nt + (r&1) < jrem
        const bool full = all_k && (j0 + 16 <= p.Lh);         // warp-uniform: no masking needed
        float qv[2][4];                                       // POOL: Q = p (1 - sum_group p) of the forward strand
        if (!POOL) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 1 && !p.ds) continue;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float x = d1[s][nt][r] + bias[r >> 1];
                        d1[s][nt][r] = fast_sigmoid(x);
                    }
                if (!full) {
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (!(kval[r >> 1] && (8 * nt + (r & 1) < jrem))) d1[s][nt][r] = 0.f;
                }
            }
        } else {
            // p = exp(x) / (P + sum_group exp(x)) (convRBM.py:252). x is clamped at 60 instead of subtracting the group maximum
            // (identical unless a weight bank has diverged; e^60 * 16 is finite in fp32).
            const uint32_t validbits = valid_mask(p.bv, j0, p.Lh);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 1 && !p.ds) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float e[4], inv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int pos = 8 * (i >> 1) + 2 * q + (i & 1);
                        const float x = fminf(d1[s][i >> 1][2 * h + (i & 1)] + bias[h], 60.f);
                        float t;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 1.4426950408889634f));
                        e[i] = (kval[h] && ((validbits >> pos) & 1u)) ? t : 0.f;
                        inv[i] = 0.f;
                    }
#pragma unroll
                    for (int gI = 0; gI < BLOCK_POS / 2; ++gI) {
                        if (gI < ng) {                         // warp-uniform
                            float part = 0.f;
#pragma unroll
                            for (int i = 0; i < 4; ++i) part += gidx[i] == gI ? e[i] : 0.f;
                            part += __shfl_xor_sync(0xffffffffu, part, 1);
                            part += __shfl_xor_sync(0xffffffffu, part, 2);
                            float r;
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fpool + part));
#pragma unroll
                            for (int i = 0; i < 4; ++i) inv[i] = gidx[i] == gI ? r : inv[i];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float pr = e[i] * inv[i];
                        d1[s][i >> 1][2 * h + (i & 1)] = pr;
                        if (s == 0) qv[i >> 1][2 * h + (i & 1)] = pr * (fpool * inv[i]);
                    }
                }
            }
        }
        if (SAMPLE && POOL) {
            // one uniform per pool group (convRBM.py:259-267): every lane of a quad gathers the 16 probabilities of its motif
            // row and walks them in position order exactly like the CUDA-core kernel (v2h.cuh sample_block); lane q == 0
            // sets the bits (the planes were zeroed before the launch; hits are rare).
            mk.valid = valid_mask(p.bv, j0, p.Lh);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 1 && !p.ds) continue;
                const float *inj = s == 0 ? p.u_fwd : p.u_rc;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = mtile * MTILE + 8 * h + g;
                    float pv[BLOCK_POS];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq)
                            pv[8 * (i >> 1) + 2 * qq + (i & 1)] = __shfl_sync(0xffffffffu, d1[s][i >> 1][2 * h + (i & 1)], (lane & 28) | qq);
                    // the block's group uniforms: injected, or one / two Philox draws per (motif row, block) — taken before the
                    // position loop so that the unrolled loop holds no copy of the generator (the first version inlined one per
                    // position: 126 us per launch at C3, the instruction cache thrashing)
                    const int g0 = j0 / p.pooling;
                    float ug[BLOCK_POS / 2];
                    if (inj) {
                        const float *ur = inj + (size_t)(n * p.K + (kval[h] ? k : 0)) * groups_per_seq + g0;
#pragma unroll
                        for (int i = 0; i < BLOCK_POS / 2; ++i) ug[i] = (i < ng && g0 + i < groups_per_seq) ? ur[i] : 1.f;
                    } else {
                        const uint32_t tag = (s == 0 ? RNG_HIDDEN : RNG_HIDDEN_RC) | (p.iter << 8);
                        const uint4 r0 = philox4x32_call(make_uint4((uint32_t)(n * p.K + k), (uint32_t)(2 * jb), ctr_hi, tag), key);
                        ug[0] = u01(r0.x); ug[1] = u01(r0.y); ug[2] = u01(r0.z); ug[3] = u01(r0.w);
                        ug[4] = ug[5] = ug[6] = ug[7] = 1.f;
                        if (ng > 4) {                          // warp-uniform
                            const uint4 r1 = philox4x32_call(make_uint4((uint32_t)(n * p.K + k), (uint32_t)(2 * jb + 1), ctr_hi, tag), key);
                            ug[4] = u01(r1.x); ug[5] = u01(r1.y); ug[6] = u01(r1.z); ug[7] = u01(r1.w);
                        }
                    }
                    uint32_t bits = 0u;
                    bool done = false;
                    float u = 0.f, run = 0.f;
                    int gi = -1;
#pragma unroll
                    for (int r = 0; r < BLOCK_POS; ++r) {
                        if ((mk.start >> r) & 1u) {            // warp-uniform
                            ++gi; done = false; run = 0.f;
                            u = ug[0];
#pragma unroll
                            for (int i = 1; i < BLOCK_POS / 2; ++i) u = gi == i ? ug[i] : u;
                        }
                        run += pv[r];
                        const bool hit = ((mk.valid >> r) & 1u) && !done && (run > u);
                        done = done || hit;
                        if (hit) bits |= 1u << r;
                    }
                    if (q == 0 && kval[h])
                        scatter_sample_bits(p.bits + ((size_t)(s * p.n_rows + n) * ktiles + (mtile >> 1)) * FANTASY_WORDS, bits, j0,
                                            (mtile & 1) * 16 + 8 * h + g);
                }
            }
        }
        if (SAMPLE && !POOL) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 1 && !p.ds) continue;
                const float *inj = s == 0 ? p.u_fwd : p.u_rc;
                uint32_t mine = 0u;
                bool hit[8];                                     // index 4*h + 2*nt + e
                if (inj) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int k = mtile * MTILE + 8 * h + g;
                        const float *up = inj + (size_t)(n * p.K + (kval[h] ? k : 0)) * p.Lh + j0 + 2 * q;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float u = (8 * (i >> 1) + (i & 1) < jrem) ? up[8 * (i >> 1) + (i & 1)] : 1.f;
                            hit[4 * h + i] = d1[s][i >> 1][2 * h + (i & 1)] > u;
                        }
                    }
                } else {
                    // One Philox draw per lane and strand = eight 16-bit uniforms for its 2 motifs x 4 positions. A sample is
                    // decided by floor(65536 p) against its 16 bits; only when the two are equal (probability 2^-16 per
                    // sample) does the warp take a second draw and compare the fraction: exact Bernoulli(p) to 2^-32, at
                    // half the generator work of one 24-bit uniform per sample (the generator was a third of this kernel).
                    const uint32_t tag = (s == 0 ? RNG_HIDDEN : RNG_HIDDEN_RC) | (p.iter << 8);
                    const uint4 ctr = make_uint4((uint32_t)(n * p.K + mtile * MTILE + g), (uint32_t)(jb * 4 + q), ctr_hi, tag);
                    const uint4 rnd = philox4x32(ctr, key);
                    const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
                    uint32_t ambm = 0u;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        // floor(65536 p) against the 16 bits, both carried in the pattern of a float near 2^23 (floor_bits, common.cuh)
                        const uint32_t tb = floor_bits(d1[s][(i >> 1) & 1][2 * (i >> 2) + (i & 1)], 65536.f);
                        const uint32_t ub = 0x4B000000u | ((rw[i >> 1] >> (16 * (i & 1))) & 0xffffu);
                        hit[i] = tb > ub;
                        ambm |= (tb == ub ? 1u : 0u) << i;
                    }
                    if (__any_sync(0xffffffffu, ambm != 0u)) {          // rare: ~1 % of the blocks
                        const uint4 r2 = philox4x32_call(make_uint4(ctr.x, ctr.y, ctr.z, tag + 16u), key);
                        const uint32_t r2w[4] = {r2.x, r2.y, r2.z, r2.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if ((ambm >> i) & 1u) {
                                const float t = d1[s][(i >> 1) & 1][2 * (i >> 2) + (i & 1)] * 65536.f;
                                hit[i] = (t - floorf(t)) * 65536.f > (float)((r2w[i >> 1] >> (16 * (i & 1))) & 0xffffu);
                            }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const uint32_t word = __ballot_sync(0xffffffffu, hit[4 * h + 2 * nt + e]);
                            if (lane == ((h * 2 + nt) * 2 + e)) mine = word;
                        }
                // the 8 words of this m-tile: word index ((mt*2+h)*2+nt)*2+e with mt = mtile & 1, k-tile = mtile >> 1
                if (lane < 8)
                    p.bits[(((size_t)s * p.n_rows + n) * ktiles + (mtile >> 1)) * (njb * 16) + jb * 16 + (mtile & 1) * 8 + lane] = mine;
            }
        }
        // ---- statistics on the tensor pipe ----------------------------------------------------------------------
        if (STAT_P || STAT_Q) {
            // A fragments (16 motifs x 16 positions) straight from the accumulator layout of the first MMA
            uint32_t ap[2][4], apr[2][4], aq[2][4];           // [split][reg]
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int reg = 2 * nt + h;
                    const float p0 = d1[0][nt][2 * h], p1 = d1[0][nt][2 * h + 1];
                    if (STAT_P) {
                        split2(p0, p1, ap[0][reg], ap[1][reg]);
                        sum_p[h] += p0 + p1;
                        if (p.ds) {
                            const float s0 = d1[1][nt][2 * h], s1 = d1[1][nt][2 * h + 1];
                            split2(s0, s1, apr[0][reg], apr[1][reg]);
                            sum_pr[h] += s0 + s1;
                        }
                    }
                    if (STAT_Q) {
                        const float q0 = POOL ? qv[nt][2 * h] : p0 * (1.f - p0), q1 = POOL ? qv[nt][2 * h + 1] : p1 * (1.f - p1);
                        split2(q0, q1, aq[0][reg], aq[1][reg]);
                        sum_q[h] += q0 + q1;
                        if (!STAT_P) sum_p[h] += p0 + p1;
                    }
                }
            uint32_t uf[3], ur[3], gf[6], gr[6];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                uf[i] = __funnelshift_r(wd[1 + i], wd[2 + i], 2 * lo2);
                const uint32_t lo = sh3 >= 16 ? wd[1 + i] : wd[i], hi = sh3 >= 16 ? wd[2 + i] : wd[1 + i];
                ur[i] = __funnelshift_r(lo, hi, 2 * (sh3 & 15));
            }
            pair_groups<3>(uf, code_f, gf);
            if (STAT_P && p.ds) pair_groups<3>(ur, code_r, gr);
#pragma unroll
            for (int t = 0; t < NT2; ++t) {
                // column c = 8t + g -> (a_c, m_c) = (g&3, 2t + g/4); K rows = positions 2q, 2q+1 | 2q+8, 2q+9
                const uint32_t bf0 = onehot_reg(gf[(2 * t) >> 3], (2 * t) & 7);
                const uint32_t bf1 = onehot_reg(gf[(2 * t + 8) >> 3], (2 * t + 8) & 7);
                if (STAT_P) {
                    hmma_bf16(accp[t], ap[0], bf0, bf1);
                    hmma_bf16(accp[t], ap[1], bf0, bf1);
                }
                if (STAT_Q) {
                    hmma_bf16(accq[t], aq[0], bf0, bf1);
                    hmma_bf16(accq[t], aq[1], bf0, bf1);
                }
                if (STAT_P && p.ds) {
                    // strand fold (convRBM.py:367): column (a_c, m_c) also collects p'[j] [base(j + M-1-m_c) == 3-a_c];
                    // with t' = NT2-1-t the base offset is static: 2*t' (+8) from a window starting at j0 + lo3 + M - 4*KS
                    const uint32_t br0 = onehot_reg(gr[(2 * (NT2 - 1 - t)) >> 3], (2 * (NT2 - 1 - t)) & 7);
                    const uint32_t br1 = onehot_reg(gr[(2 * (NT2 - 1 - t) + 8) >> 3], (2 * (NT2 - 1 - t) + 8) & 7);
                    hmma_bf16(accp[t], apr[0], br0, br1);
                    hmma_bf16(accp[t], apr[1], br0, br1);
                }
            }
        }
    }

    // ---- reduce over the warps of the CTA, write the CTA partial ------------------------------------------------------
    // Partial layout (per CTA): [NACC][32]: slot t*4 + r of statistic P, then of Q, then sum_p[2], sum_pr[2], sum_q[2].
    if (STAT_P || STAT_Q) {
        constexpr int NACC = Ac::NACC;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                sum_p[h] += __shfl_xor_sync(0xffffffffu, sum_p[h], o);
                sum_pr[h] += __shfl_xor_sync(0xffffffffu, sum_pr[h], o);
                sum_q[h] += __shfl_xor_sync(0xffffffffu, sum_q[h], o);
            }
            if (q != 0) { sum_p[h] = 0.f; sum_pr[h] = 0.f; sum_q[h] = 0.f; }
        }
        __syncthreads();
        float *scr = reinterpret_cast<float *>(smem);
        auto visit = [&](auto &&f) {
            if (STAT_P) {
#pragma unroll
                for (int t = 0; t < NT2; ++t)
#pragma unroll
                    for (int r = 0; r < 4; ++r) f(t * 4 + r, accp[t][r]);
            }
            if (STAT_Q) {
#pragma unroll
                for (int t = 0; t < NT2; ++t)
#pragma unroll
                    for (int r = 0; r < 4; ++r) f(Ac::NP + t * 4 + r, accq[t][r]);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                f(Ac::NP + Ac::NQ + h, sum_p[h]);
                f(Ac::NP + Ac::NQ + 2 + h, sum_pr[h]);
                f(Ac::NP + Ac::NQ + 4 + h, sum_q[h]);
            }
        };
        for (int half = nwarps >> 1; half >= 1; half >>= 1) {
            if (warp >= half && warp < 2 * half) {
                float *dst = scr + (size_t)(warp - half) * NACC * 32 + lane;
                visit([&](int o, float &v) { dst[o * 32] = v; });
            }
            __syncthreads();
            if (warp < half) {
                const float *src = scr + (size_t)warp * NACC * 32 + lane;
                visit([&](int o, float &v) { v += src[o * 32]; });
            }
            __syncthreads();
        }
        if (warp == 0) {
            float *dst = p.partial + ((size_t)(mtile * gy + by) * NACC) * 32 + lane;
            visit([&](int o, float &v) { dst[o * 32] = v; });
        }
    }
    if (p.pdl_late) pdl_wait();
}

template <int KS, int FLAGS, bool POOL>
__global__ void __launch_bounds__(256, 2) v2h_mma_kernel(const V2HParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    v2h_mma_body<KS, FLAGS, POOL>(p, smem, blockIdx.x, blockIdx.y, gridDim.y);
}

// Two roles in ONE grid (un-pooled, motif_length <= 16, the last chain sweep of a step): the CTAs with blockIdx.y < gy_pos run the
// positive phase on the minibatch (P and Q statistics), the others the chain sweep (sampling + model statistics). The positive
// phase depends on nothing in flight, so its CTAs — dispatched first — start while the predecessor (h2v of the chain) is still
// running and share the SMs with it; the chain CTAs block in griddepcontrol.wait until h2v has completed. One launch boundary,
// one prologue / ramp / tail per step less than two back-to-back single-role launches (C2: 42 -> ~35 us per step), and the h2v
// kernel (latency bound, few registers) no longer has the GPU to itself.
struct V2HDualParams {
    V2HParams pos, neg;
    int gy_pos;
};
template <int KS>
__global__ void __launch_bounds__(256, 2) v2h_mma_dual_kernel(const V2HDualParams d) {
    extern __shared__ __align__(128) unsigned char smem[];
    if ((int)blockIdx.y < d.gy_pos) v2h_mma_body<KS, F_STAT_P | F_STAT_Q, false>(d.pos, smem, blockIdx.x, blockIdx.y, d.gy_pos);
    else v2h_mma_body<KS, F_STAT_P | F_SAMPLE, false>(d.neg, smem, blockIdx.x, blockIdx.y - d.gy_pos, gridDim.y - d.gy_pos);
}

// ----------------------------------------------------------------------------------------------------------
// Dense output (motif scan `getHitProbs`, eager `_bottomUpActivity` / `_bottomUpProbability`) for pooling == 1:
// the same tensor-pipe up-pass, accumulator fragments stored straight to out[n][k][j] (each quad writes 32 contiguous
// bytes of one motif row; the two n-tiles of a block complete a 64-byte run). HBM-write bound by design: 4*K bytes
// per nucleotide against 0.25 bytes read.
// ----------------------------------------------------------------------------------------------------------
// Store shape (scratch/store_pattern.cu, no arithmetic, same out[n][k][j] layout): accumulator fragments written
// directly = 32-byte pieces of 8 rows per instruction -> 1.4 TB/s; 128 contiguous bytes of one row per instruction
// -> 3.4 TB/s. So a warp computes 16 motifs x 32 positions, transposes the tile through a private 2 KB shared-memory
// buffer and writes one row piece (32 floats) per instruction.
template <int KS>
__global__ void __launch_bounds__(256, 3) v2h_dense_mma_kernel(const DenseParams p) {
    using Sh = MmaShape<KS>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *wimg = reinterpret_cast<uint32_t *>(smem);
    float *wtile = reinterpret_cast<float *>(smem + Sh::wimg_bytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + Sh::wimg_bytes + Sh::wtile_bytes);
    float *stage_all = reinterpret_cast<float *>(smem + Sh::wimg_bytes + Sh::wtile_bytes + 16);    // [warp][16][34]
    const int mtile = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    build_w_fragments<KS>(wimg, wtile, bar, p.W, p.K, p.M, mtile);
    const uint4 *wfrag = reinterpret_cast<const uint4 *>(wimg) + lane;
    const bool want_f = p.strand_mode != CRBM_STRAND_RC, want_r = p.strand_mode != CRBM_STRAND_FWD;
    const float bscale = p.strand_mode == CRBM_STRAND_SUM ? 2.f : 1.f;        // both activities carry the bias (convRBM.py:513)
    float bias[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = mtile * MTILE + 8 * h + g;
        bias[h] = k < p.K ? bscale * __ldg(p.b + k) : 0.f;
    }
    const int nrows = min(MTILE, p.K - mtile * MTILE);
    const int lo1 = g + (q >> 1);
    const uint32_t qmask = (q & 1) ? 0u : 0x55555555u;
    const bool prob = p.out_kind == CRBM_OUT_PROBABILITY;
    constexpr int SS = 34;                                   // stage row stride (floats): conflict-free float2 writes
    float *stage = stage_all + warp * (MTILE * SS);
    const int nblk2 = (p.nblk + 1) >> 1, n_items2 = p.n_rows * nblk2;

    for (int item = blockIdx.y * nwarps + warp; item < n_items2; item += gridDim.y * nwarps) {
        const int n = item / nblk2, jb2 = item - n * nblk2, j0 = jb2 * 32;
        const uint32_t *row = p.packed + (size_t)n * p.row_words + (j0 >> 4);
        uint32_t wd[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) wd[i] = __ldg(row + i);
        float d1[4][4];                                      // [nt][reg], nt = 8-position tile of the 32-position item
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) d1[nt][r] = 0.f;
        uint32_t vg[2][6];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) v[i] = __funnelshift_r(wd[1 + half + i], wd[2 + half + i], 2 * lo1);
            upass_groups<3>(v, qmask, vg[half]);
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t bx[4][2];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    const int off = 8 * (nt & 1) + 4 * ks + 2 * hb;
                    bx[nt][hb] = onehot_reg_h(vg[nt >> 1][off >> 3], off & 7);
                }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 0 ? !want_f : !want_r) continue;
#pragma unroll
                for (int sp = 0; sp < NSPLIT; ++sp) {
                    const uint4 a = wfrag[((s * NSPLIT + sp) * KS + ks) * 32];
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) hmma_f16(d1[nt], a, bx[nt][0], bx[nt][1]);
                }
            }
        }
        // epilogue -> stage[row][position]
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                float x0 = d1[nt][2 * h] + bias[h], x1 = d1[nt][2 * h + 1] + bias[h];
                if (prob) {
                    x0 = fast_sigmoid(x0);
                    x1 = fast_sigmoid(x1);
                }
                *reinterpret_cast<float2 *>(stage + (8 * h + g) * SS + 8 * nt + 2 * q) = make_float2(x0, x1);
            }
        __syncwarp();
        const int jn = min(32, p.Lh - j0);                   // valid positions of this item
        float *obase = p.out + ((size_t)n * p.K + mtile * MTILE) * p.Lh + j0 + lane;
#pragma unroll 4
        for (int r = 0; r < nrows; ++r)
            if (lane < jn) __stcs(obase + (size_t)r * p.Lh, stage[r * SS + lane]);
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------------------------
// Motif scan with the reductions its callers apply fused in (reference secomo/utils.py:155,245-247 take
// the max over positions of motifHitProbs, :306 the mean): per sequence and motif the maximum hit probability, its
// position (first maximum, like numpy.argmax) and the mean over positions. The 4*K bytes/nt of output never reach HBM,
// which turns the scan from HBM-write bound into compute bound and makes genome-scale inputs fit.
// One warp walks all position blocks of one sequence for the CTA's m-tile and keeps running (max, argmax, sum) in
// registers; the four lanes of a quad hold different positions of the same motif rows and are combined by shuffles.
// ----------------------------------------------------------------------------------------------------------
struct ScanReduceParams {
    DenseParams d;
    float *out_max, *out_mean;     // (N,K) each, may be null
    int *out_argmax;               // (N,K), may be null
};

template <int KS>
__global__ void __launch_bounds__(256, 3) v2h_scan_reduce_kernel(const ScanReduceParams sp) {
    using Sh = MmaShape<KS>;
    const DenseParams &p = sp.d;
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *wimg = reinterpret_cast<uint32_t *>(smem);
    float *wtile = reinterpret_cast<float *>(smem + Sh::wimg_bytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + Sh::wimg_bytes + Sh::wtile_bytes);
    const int mtile = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    build_w_fragments<KS>(wimg, wtile, bar, p.W, p.K, p.M, mtile);
    const uint4 *wfrag = reinterpret_cast<const uint4 *>(wimg) + lane;
    const bool want_f = p.strand_mode != CRBM_STRAND_RC, want_r = p.strand_mode != CRBM_STRAND_FWD;
    const float bscale = p.strand_mode == CRBM_STRAND_SUM ? 2.f : 1.f;
    float bias[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = mtile * MTILE + 8 * h + g;
        bias[h] = k < p.K ? bscale * __ldg(p.b + k) : 0.f;
    }
    const int lo1 = g + (q >> 1);
    const uint32_t qmask = (q & 1) ? 0u : 0x55555555u;
    const bool prob = p.out_kind == CRBM_OUT_PROBABILITY;
    // pooled probabilities: blocks of bv = (16 / P) * P positions and quad-reduced group sums, as in v2h_mma_kernel<POOL>
    const bool pooled = prob && p.pooling > 1;
    const int bstep = pooled ? p.bv : 16, ng = pooled ? p.bv / p.pooling : 0;
    const float fpool = (float)p.pooling;
    int gidx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gidx[i] = (8 * (i >> 1) + 2 * q + (i & 1)) / p.pooling;

    for (int n = blockIdx.y * nwarps + warp; n < p.n_rows; n += gridDim.y * nwarps) {
        float best[2] = {-INFINITY, -INFINITY}, sum[2] = {0.f, 0.f};
        int bpos[2] = {0, 0};
        const uint32_t *row0 = p.packed + (size_t)n * p.row_words;
        for (int jb = 0; jb < p.nblk; ++jb) {
            const int j0 = jb * bstep;
            uint32_t wd[5];
            {
                uint32_t rawd[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) rawd[i] = __ldg(row0 + (j0 >> 4) + i);
                const int shw = 2 * (j0 & 15);                                  // warp-uniform; 0 unless pooled
#pragma unroll
                for (int i = 0; i < 5; ++i) wd[i] = __funnelshift_r(rawd[i], rawd[i + 1], shw);
            }
            float d1[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r) d1[nt][r] = 0.f;
            uint32_t v[3], vg[6];
#pragma unroll
            for (int i = 0; i < 3; ++i) v[i] = __funnelshift_r(wd[1 + i], wd[2 + i], 2 * lo1);
            upass_groups<3>(v, qmask, vg);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t bx[2][2];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        const int off = 8 * nt + 4 * ks + 2 * hb;
                        bx[nt][hb] = onehot_reg_h(vg[off >> 3], off & 7);
                    }
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (s == 0 ? !want_f : !want_r) continue;
#pragma unroll
                    for (int spl = 0; spl < NSPLIT; ++spl) {
                        const uint4 a = wfrag[((s * NSPLIT + spl) * KS + ks) * 32];
                        hmma_f16(d1[0], a, bx[0][0], bx[0][1]);
                        hmma_f16(d1[1], a, bx[1][0], bx[1][1]);
                    }
                }
            }
            const int jrem = min(p.Lh - j0, bstep) - 2 * q;      // position 8nt + 2q + e of the block is real iff 8nt + e < jrem
            if (pooled) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float ev[4], inv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float x = fminf(d1[i >> 1][2 * h + (i & 1)] + bias[h], 60.f);
                        float t;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 1.4426950408889634f));
                        ev[i] = (8 * (i >> 1) + (i & 1) < jrem) ? t : 0.f;
                        inv[i] = 0.f;
                    }
#pragma unroll
                    for (int gI = 0; gI < BLOCK_POS / 2; ++gI) {
                        if (gI < ng) {                         // warp-uniform
                            float part = 0.f;
#pragma unroll
                            for (int i = 0; i < 4; ++i) part += gidx[i] == gI ? ev[i] : 0.f;
                            part += __shfl_xor_sync(0xffffffffu, part, 1);
                            part += __shfl_xor_sync(0xffffffffu, part, 2);
                            float r;
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fpool + part));
#pragma unroll
                            for (int i = 0; i < 4; ++i) inv[i] = gidx[i] == gI ? r : inv[i];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) d1[i >> 1][2 * h + (i & 1)] = ev[i] * inv[i];
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (8 * nt + e < jrem) {
                            float x = d1[nt][2 * h + e];
                            if (!pooled) {
                                x += bias[h];
                                if (prob) x = fast_sigmoid(x);
                            }
                            sum[h] += x;
                            if (x > best[h]) { best[h] = x; bpos[h] = j0 + 8 * nt + 2 * q + e; }   // ascending positions: first max kept
                        }
                    }
        }
        // combine the four lanes of the quad (same motif rows, interleaved positions): max with smallest position on ties
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best[h], o);
                const int op = __shfl_xor_sync(0xffffffffu, bpos[h], o);
                sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], o);
                if (ob > best[h] || (ob == best[h] && op < bpos[h])) { best[h] = ob; bpos[h] = op; }
            }
            const int k = mtile * MTILE + 8 * h + g;
            if (q == 0 && k < p.K) {
                const size_t o = (size_t)n * p.K + k;
                if (sp.out_max) sp.out_max[o] = best[h];
                if (sp.out_argmax) sp.out_argmax[o] = bpos[h];
                if (sp.out_mean) sp.out_mean[o] = sum[h] / (float)p.Lh;
            }
        }
    }
}

template <int KS> int launch_v2h_mma(int flags, const V2HParams &p, dim3 grid, cudaStream_t stream);
template <int KS> int launch_v2h_mma_dual(const V2HDualParams &d, dim3 grid, cudaStream_t stream);
template <int KS> int launch_v2h_scan_reduce(const ScanReduceParams &p, cudaStream_t stream);
template <int KS> int launch_v2h_dense_mma(const DenseParams &p, dim3 grid, cudaStream_t stream);

}  // namespace crbm
