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
#include "v2h.cuh"

namespace crbm {

template <int KS> struct MmaShape {
    static constexpr int NT2 = 2 * KS;                          // 8-column tiles over c = 4*m + a  (16*KS columns)
    static constexpr int FRAGS = 2 * 3 * 2 * KS;                // [strand][split][mt][ks]
    static constexpr size_t wimg_bytes = (size_t)FRAGS * 512;   // 32 lanes x 16 bytes per fragment
    static constexpr size_t wtile_bytes = (size_t)KTILE * 4 * CRBM_MAX_MOTIF_LEN * sizeof(float);
    static constexpr size_t smem_bytes = wimg_bytes + wtile_bytes + 16;
};
template <int KS, int FLAGS> constexpr size_t mma_smem_bytes() {
    // W fragments + staging tile, or the scratch of the final 8-warp tree reduction, whichever is larger
    const size_t a = MmaShape<KS>::smem_bytes;
    const size_t nacc = ((FLAGS & F_STAT_P) ? 16 * KS : 0) + ((FLAGS & F_STAT_Q) ? 16 * KS : 0) + 12;
    const size_t b = 4 * nacc * 32 * sizeof(float);
    return a > b ? a : b;
}

template <int KS, int FLAGS> struct MmaAcc {
    static constexpr int NP = (FLAGS & F_STAT_P) ? 16 * KS : 0;      // registers per thread
    static constexpr int NQ = (FLAGS & F_STAT_Q) ? 16 * KS : 0;
    static constexpr int NACC = NP + NQ + 12;                        // + sumP, sumPr, sumQ for the thread's 4 motif rows
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

// W tile (32 motifs x 4 x M, fp32, staged by the TMA engine) -> three bf16 images in MMA A-fragment order, both strands.
// Element (row r, column c) of the (mt, ks) fragment is Wx[16*mt + r][16*ks + c]; forward strand Wx[k][4m+a] = W[k][a][m],
// reverse strand Wx[k][4m+a] = W[k][3-a][M-1-m] (convRBM.py:241 with filter_flip=True); zero beyond M.
template <int KS>
__device__ __forceinline__ void build_w_fragments(uint32_t *wimg, float *wtile, uint64_t *bar,
                                                  const float *__restrict__ W, int K, int M, int kt) {
    const int k0 = kt * KTILE;
    const int nk = min(KTILE, K - k0);
    const uint32_t bytes = (uint32_t)nk * 4u * (uint32_t)M * 4u;
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        tma_load_1d(wtile, W + (size_t)k0 * 4 * M, bytes, bar);
    }
    mbar_wait(bar, 0);
    // one work unit = one 32-bit register of one fragment: (strand, mt, ks, lane, reg)
    const int units = 2 * 2 * KS * 32 * 4;
    for (int u = threadIdx.x; u < units; u += blockDim.x) {
        const int reg = u & 3, lane = (u >> 2) & 31, rest = u >> 7;
        const int ks = rest % KS, mt = (rest / KS) & 1, strand = rest / (2 * KS);
        const int g = lane >> 2, q = lane & 3;
        const int row = 16 * mt + g + 8 * (reg & 1);
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = 16 * ks + 2 * q + e + 8 * (reg >> 1);
            const int m = c >> 2, a = c & 3;
            float w = 0.f;
            if (m < M && row < nk) w = strand == 0 ? wtile[(row * 4 + a) * M + m] : wtile[(row * 4 + 3 - a) * M + (M - 1 - m)];
            v[e] = w;
        }
        uint32_t hi, mid, lo;
        split2(v[0], v[1], hi, mid);
        const float r0 = (v[0] - __uint_as_float(hi << 16)) - __uint_as_float(mid << 16);
        const float r1 = (v[1] - __uint_as_float(hi & 0xffff0000u)) - __uint_as_float(mid & 0xffff0000u);
        lo = pack_bf16x2(r0, r1);
        const int f0 = ((strand * 3 + 0) * 2 + mt) * KS + ks;
        const int stride = 2 * KS;                                   // fragments between consecutive splits
        wimg[(f0 * 32 + lane) * 4 + reg] = hi;
        wimg[((f0 + stride) * 32 + lane) * 4 + reg] = mid;
        wimg[((f0 + 2 * stride) * 32 + lane) * 4 + reg] = lo;
    }
    __syncthreads();
}

// bf16x2 B-fragment register for two consecutive sequence positions against one base code: 1.0 where equal.
__device__ __forceinline__ uint32_t onehot_pair(uint32_t two_bases /*4 bits*/, uint32_t code_x5) {
    const uint32_t t = two_bases ^ code_x5;
    return ((t & 3u) ? 0u : 0x3F80u) | ((t & 12u) ? 0u : 0x3F800000u);
}

template <int KS, int FLAGS>
__global__ void __launch_bounds__(256, 1) v2h_mma_kernel(const V2HParams p) {
    using Sh = MmaShape<KS>;
    using Ac = MmaAcc<KS, FLAGS>;
    constexpr int NT2 = Sh::NT2;
    constexpr bool STAT_P = FLAGS & F_STAT_P, STAT_Q = FLAGS & F_STAT_Q, SAMPLE = FLAGS & F_SAMPLE;
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *wimg = reinterpret_cast<uint32_t *>(smem);
    float *wtile = reinterpret_cast<float *>(smem + Sh::wimg_bytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + Sh::wimg_bytes + Sh::wtile_bytes);

    const int kt = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    build_w_fragments<KS>(wimg, wtile, bar, p.W, p.K, p.M, kt);
    const uint4 *wfrag = reinterpret_cast<const uint4 *>(wimg) + lane;

    // the thread's four motif rows: r4 = 2*mt + h  ->  k = 32*kt + 16*mt + 8*h + g
    float bias[4];
    bool kval[4];
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
        const int k = kt * KTILE + 16 * (r4 >> 1) + 8 * (r4 & 1) + g;
        kval[r4] = k < p.K;
        bias[r4] = kval[r4] ? __ldg(p.b + k) : 0.f;
    }

    float accp[2][NT2][4], accq[2][NT2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int t = 0; t < NT2; ++t)
#pragma unroll
            for (int r = 0; r < 4; ++r) { accp[mt][t][r] = 0.f; accq[mt][t][r] = 0.f; }
    float sum_p[4] = {0.f, 0.f, 0.f, 0.f}, sum_pr[4] = {0.f, 0.f, 0.f, 0.f}, sum_q[4] = {0.f, 0.f, 0.f, 0.f};

    // lane-dependent parts of the base offsets (see the derivations in the header comment of each stage)
    const int lo1 = g + (q >> 1);                       // up-pass B: base j0 + 8nt + g + 4ks + q/2 (+2)
    const int lo2 = 2 * q + (g >> 2);                   // forward statistics B: base j0 + 2q + m_c (+1) (+8), m_c = 2nt2 + g/4
    const int mshift = p.M - 4 * KS;                    // in [-3, 0]
    const int sh3 = 16 + 2 * q - (g >> 2) + 1 + mshift; // reverse statistics: window starts at base j0 - 16 + sh3
    const uint32_t qodd = q & 1;
    const uint32_t code_f = (uint32_t)(g & 3) * 5u, code_r = (uint32_t)(3 - (g & 3)) * 5u;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const uint32_t ctr_hi = p.ctr_hi + (p.step_dev ? (uint32_t)(*p.step_dev) * 64u : 0u);
    const int ktiles = gridDim.x;
    const int njb = (p.Lh + 15) >> 4;

    const uint32_t *packed = p.packed + (p.batch_index ? (size_t)(*p.batch_index) * p.n_rows * p.row_words : 0);
    for (int item = blockIdx.y * nwarps + warp; item < p.n_items; item += gridDim.y * nwarps) {
        const int n = item / p.nblk, jb = item - n * p.nblk, j0 = jb * 16;
        const uint32_t *row = packed + (size_t)n * p.row_words + (j0 >> 4);     // word holding bases j0-16 .. j0-1
        uint32_t wd[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) wd[i] = __ldg(row + i);

        // ---- x = W (x) X on the tensor pipe, both strands ------------------------------------------------------
        float d1[2][2][2][4];                                  // [strand][mt][nt][reg]
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) d1[s][mt][nt][r] = 0.f;
        {
            uint32_t v[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) v[i] = __funnelshift_r(wd[1 + i], wd[2 + i], 2 * lo1);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t bx[2][2];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        const uint32_t bb = window_bits<1, 3>(v, 8 * nt + 4 * ks + 2 * hb);
                        bx[nt][hb] = ((bb >> 1) == qodd) ? (0x3F80u << ((bb & 1u) << 4)) : 0u;
                    }
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (s == 1 && !p.ds) continue;
#pragma unroll
                    for (int sp = 0; sp < 3; ++sp)
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const uint4 a = wfrag[(((s * 3 + sp) * 2 + mt) * KS + ks) * 32];
                            hmma_bf16(d1[s][mt][0], a, bx[0][0], bx[0][1]);
                            hmma_bf16(d1[s][mt][1], a, bx[1][0], bx[1][1]);
                        }
                }
            }
        }
        // ---- sigmoid epilogue, sampling -------------------------------------------------------------------------
        // d1[s][mt][nt][r]: motif row r4 = 2*mt + (r>>1), position j0 + 8*nt + 2*q + (r&1)
        const int jrem = p.Lh - j0 - 2 * q;                   // position valid iff 8*nt + (r&1) < jrem
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (s == 1 && !p.ds) continue;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int r4 = 2 * mt + (r >> 1);
                        const bool ok = kval[r4] && (8 * nt + (r & 1) < jrem);
                        const float x = d1[s][mt][nt][r] + bias[r4];
                        d1[s][mt][nt][r] = ok ? __fdividef(1.0f, 1.0f + __expf(-x)) : 0.f;
                    }
        }
        if (SAMPLE) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 1 && !p.ds) continue;
                const float *inj = s == 0 ? p.u_fwd : p.u_rc;
                uint32_t mine = 0u;
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4) {
                    const int mt = r4 >> 1, h = r4 & 1;
                    const int k = kt * KTILE + 16 * mt + 8 * h + g;
                    float u[4];                                  // positions (nt, e): index 2*nt + e
                    if (inj) {
                        const float *up = inj + (size_t)(n * p.K + (kval[r4] ? k : 0)) * p.Lh + j0 + 2 * q;
#pragma unroll
                        for (int i = 0; i < 4; ++i) u[i] = (8 * (i >> 1) + (i & 1) < jrem) ? up[8 * (i >> 1) + (i & 1)] : 1.f;
                    } else {
                        const uint4 rnd = philox4x32(make_uint4((uint32_t)(n * p.K + k), (uint32_t)(jb * 4 + q), ctr_hi,
                                                                s == 0 ? RNG_HIDDEN : RNG_HIDDEN_RC), key);
                        u[0] = u01(rnd.x); u[1] = u01(rnd.y); u[2] = u01(rnd.z); u[3] = u01(rnd.w);
                    }
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const bool hit = d1[s][mt][nt][2 * h + e] > u[2 * nt + e];
                            const uint32_t word = __ballot_sync(0xffffffffu, hit);
                            if (lane == ((r4 * 2 + nt) * 2 + e)) mine = word;
                        }
                }
                if (lane < 16)
                    p.bits[(((size_t)s * p.n_rows + n) * ktiles + kt) * (njb * 16) + jb * 16 + lane] = mine;
            }
        }
        // ---- statistics on the tensor pipe ----------------------------------------------------------------------
        if (STAT_P || STAT_Q) {
            // A fragments (16 motifs x 16 positions) straight from the accumulator layout of the first MMA
            uint32_t ap[2][2][4], apr[2][2][4], aq[2][2][4];           // [split][mt][reg]
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int reg = 2 * nt + h, r4 = 2 * mt + h;
                        const float p0 = d1[0][mt][nt][2 * h], p1 = d1[0][mt][nt][2 * h + 1];
                        if (STAT_P) {
                            split2(p0, p1, ap[0][mt][reg], ap[1][mt][reg]);
                            sum_p[r4] += p0 + p1;
                            if (p.ds) {
                                const float s0 = d1[1][mt][nt][2 * h], s1 = d1[1][mt][nt][2 * h + 1];
                                split2(s0, s1, apr[0][mt][reg], apr[1][mt][reg]);
                                sum_pr[r4] += s0 + s1;
                            }
                        }
                        if (STAT_Q) {
                            const float q0 = p0 * (1.f - p0), q1 = p1 * (1.f - p1);
                            split2(q0, q1, aq[0][mt][reg], aq[1][mt][reg]);
                            sum_q[r4] += q0 + q1;
                            if (!STAT_P) sum_p[r4] += p0 + p1;
                        }
                    }
            uint32_t uf[3], ur[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                uf[i] = __funnelshift_r(wd[1 + i], wd[2 + i], 2 * lo2);
                const uint32_t lo = sh3 >= 16 ? wd[1 + i] : wd[i], hi = sh3 >= 16 ? wd[2 + i] : wd[1 + i];
                ur[i] = __funnelshift_r(lo, hi, 2 * (sh3 & 15));
            }
#pragma unroll
            for (int t = 0; t < NT2; ++t) {
                // column c = 8t + g -> (a_c, m_c) = (g&3, 2t + g/4); K rows = positions 2q, 2q+1 | 2q+8, 2q+9
                const uint32_t bf0 = onehot_pair(window_bits<2, 3>(uf, 2 * t), code_f);
                const uint32_t bf1 = onehot_pair(window_bits<2, 3>(uf, 2 * t + 8), code_f);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    if (STAT_P) {
                        hmma_bf16(accp[mt][t], ap[0][mt], bf0, bf1);
                        hmma_bf16(accp[mt][t], ap[1][mt], bf0, bf1);
                    }
                    if (STAT_Q) {
                        hmma_bf16(accq[mt][t], aq[0][mt], bf0, bf1);
                        hmma_bf16(accq[mt][t], aq[1][mt], bf0, bf1);
                    }
                }
                if (STAT_P && p.ds) {
                    // strand fold (convRBM.py:367): column (a_c, m_c) also collects p'[j] [base(j + M-1-m_c) == 3-a_c];
                    // with t' = NT2-1-t the base offset is static: 2*t' (+8) from a window starting at j0 + lo3 + M - 4*KS
                    const uint32_t br0 = onehot_pair(window_bits<2, 3>(ur, 2 * (NT2 - 1 - t)), code_r);
                    const uint32_t br1 = onehot_pair(window_bits<2, 3>(ur, 2 * (NT2 - 1 - t) + 8), code_r);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        hmma_bf16(accp[mt][t], apr[0][mt], br0, br1);
                        hmma_bf16(accp[mt][t], apr[1][mt], br0, br1);
                    }
                }
            }
        }
    }

    // ---- reduce over the warps of the CTA, write the CTA partial ------------------------------------------------------
    if (STAT_P || STAT_Q) {
        constexpr int NACC = Ac::NACC;
        // per-motif sums: add the four lanes of a quad (they hold different positions of the same motif rows)
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                sum_p[r4] += __shfl_xor_sync(0xffffffffu, sum_p[r4], o);
                sum_pr[r4] += __shfl_xor_sync(0xffffffffu, sum_pr[r4], o);
                sum_q[r4] += __shfl_xor_sync(0xffffffffu, sum_q[r4], o);
            }
            if (q != 0) { sum_p[r4] = 0.f; sum_pr[r4] = 0.f; sum_q[r4] = 0.f; }
        }
        __syncthreads();
        float *scr = reinterpret_cast<float *>(smem);
        auto put = [&](float *dst) {
            int o = 0;
            if (STAT_P) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int t = 0; t < NT2; ++t)
#pragma unroll
                        for (int r = 0; r < 4; ++r) dst[(o++) * 32] = accp[mt][t][r];
            }
            if (STAT_Q) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int t = 0; t < NT2; ++t)
#pragma unroll
                        for (int r = 0; r < 4; ++r) dst[(o++) * 32] = accq[mt][t][r];
            }
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) dst[(o++) * 32] = sum_p[r4];
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) dst[(o++) * 32] = sum_pr[r4];
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) dst[(o++) * 32] = sum_q[r4];
        };
        auto add = [&](const float *src) {
            int o = 0;
            if (STAT_P) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int t = 0; t < NT2; ++t)
#pragma unroll
                        for (int r = 0; r < 4; ++r) accp[mt][t][r] += src[(o++) * 32];
            }
            if (STAT_Q) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int t = 0; t < NT2; ++t)
#pragma unroll
                        for (int r = 0; r < 4; ++r) accq[mt][t][r] += src[(o++) * 32];
            }
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) sum_p[r4] += src[(o++) * 32];
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) sum_pr[r4] += src[(o++) * 32];
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) sum_q[r4] += src[(o++) * 32];
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
// Dense output (motif scan `getHitProbs`, eager `_bottomUpActivity` / `_bottomUpProbability`) for pooling == 1:
// the same tensor-pipe up-pass, accumulator fragments stored straight to out[n][k][j] (each quad writes 32 contiguous
// bytes of one motif row; the two n-tiles of a block complete a 64-byte run). HBM-write bound by design: 4*K bytes
// per nucleotide against 0.25 bytes read.
// ----------------------------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(256, 3) v2h_dense_mma_kernel(const DenseParams p) {
    using Sh = MmaShape<KS>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *wimg = reinterpret_cast<uint32_t *>(smem);
    float *wtile = reinterpret_cast<float *>(smem + Sh::wimg_bytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + Sh::wimg_bytes + Sh::wtile_bytes);
    const int kt = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    build_w_fragments<KS>(wimg, wtile, bar, p.W, p.K, p.M, kt);
    const uint4 *wfrag = reinterpret_cast<const uint4 *>(wimg) + lane;
    const bool want_f = p.strand_mode != CRBM_STRAND_RC, want_r = p.strand_mode != CRBM_STRAND_FWD;
    const float bscale = p.strand_mode == CRBM_STRAND_SUM ? 2.f : 1.f;        // both activities carry the bias (convRBM.py:513)
    float bias[4];
    bool kval[4];
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
        const int k = kt * KTILE + 16 * (r4 >> 1) + 8 * (r4 & 1) + g;
        kval[r4] = k < p.K;
        bias[r4] = kval[r4] ? bscale * __ldg(p.b + k) : 0.f;
    }
    const bool mt_on[2] = {kt * KTILE < p.K, kt * KTILE + 16 < p.K};
    const int lo1 = g + (q >> 1);
    const uint32_t qodd = q & 1;
    const bool prob = p.out_kind == CRBM_OUT_PROBABILITY;
    // Lh even and 8-byte aligned rows allow float2 stores
    const bool vec2 = (p.Lh & 1) == 0;

    for (int item = blockIdx.y * nwarps + warp; item < p.n_items; item += gridDim.y * nwarps) {
        const int n = item / p.nblk, jb = item - n * p.nblk, j0 = jb * 16;
        const uint32_t *row = p.packed + (size_t)n * p.row_words + (j0 >> 4);
        uint32_t wd[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) wd[i] = __ldg(row + i);
        float d1[2][2][4];                                  // [mt][nt][reg]
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r) d1[mt][nt][r] = 0.f;
        uint32_t v[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) v[i] = __funnelshift_r(wd[1 + i], wd[2 + i], 2 * lo1);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t bx[2][2];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    const uint32_t bb = window_bits<1, 3>(v, 8 * nt + 4 * ks + 2 * hb);
                    bx[nt][hb] = ((bb >> 1) == qodd) ? (0x3F80u << ((bb & 1u) << 4)) : 0u;
                }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 0 ? !want_f : !want_r) continue;
#pragma unroll
                for (int sp = 0; sp < 3; ++sp)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        if (!mt_on[mt]) continue;
                        const uint4 a = wfrag[(((s * 3 + sp) * 2 + mt) * KS + ks) * 32];
                        hmma_bf16(d1[mt][0], a, bx[0][0], bx[0][1]);
                        hmma_bf16(d1[mt][1], a, bx[1][0], bx[1][1]);
                    }
            }
        }
        const int jrem = p.Lh - j0 - 2 * q;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            if (!mt_on[mt]) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r4 = 2 * mt + h;
                if (!kval[r4]) continue;
                float *orow = p.out + ((size_t)n * p.K + kt * KTILE + 16 * mt + 8 * h + g) * p.Lh + j0 + 2 * q;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    float x0 = d1[mt][nt][2 * h] + bias[r4], x1 = d1[mt][nt][2 * h + 1] + bias[r4];
                    if (prob) {
                        x0 = __fdividef(1.0f, 1.0f + __expf(-x0));
                        x1 = __fdividef(1.0f, 1.0f + __expf(-x1));
                    }
                    if (vec2 && 8 * nt + 1 < jrem) {
                        __stcs(reinterpret_cast<float2 *>(orow + 8 * nt), make_float2(x0, x1));
                    } else {
                        if (8 * nt < jrem) __stcs(orow + 8 * nt, x0);
                        if (8 * nt + 1 < jrem) __stcs(orow + 8 * nt + 1, x1);
                    }
                }
            }
        }
    }
}

template <int KS> int launch_v2h_mma(int flags, const V2HParams &p, dim3 grid, cudaStream_t stream);
template <int KS> int launch_v2h_dense_mma(const DenseParams &p, dim3 grid, cudaStream_t stream);

}  // namespace crbm
