// Training contractions of the cRBM on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory).
//
// Reference semantics (secomo/convRBM.py): A1/A2 :238-257 (both strands), A3 :259-267, A8-A10 :327-371, A11 :440-451 —
// the same quantities as v2h_mma.cuh (which stays the path for motif_length > 32 and unusual pooling factors).
//
// Tile orientation: MOTIF = TMEM LANE. One CTA owns a chunk of 64 motifs; MMA row r = 64 s + kl is (strand s, motif kl),
// so one instruction produces both strands. A unit of work is 64 consecutive hidden positions (two 32-position blocks).
//   GEMM 1   D1[r, pos]  = Wx[r, c] . X[pos, c]^T        c = 4m + a, X = one-hot of the M bases under position pos
//            A = Wx: W (forward rows) and W[k, 3-a, M-1-m] (reverse-complement rows), split into two fp16 terms hi + lo
//            (22 significand bits), pre-scaled by -log2(e), resident in shared memory, K-major SWIZZLE_128B;
//            B = X: built in shared memory from the packed 2-bit words (never read from HBM), K-major SWIZZLE_128B.
//   epilogue thread = one (strand, motif) row, registers = 32 consecutive positions (tcgen05.ld 32x32b): sigmoid or the
//            pooled softmax (pool groups are runs of registers: no shuffles), Q = p (1 - sum_group p), sum_j p in a
//            register, Bernoulli / multinomial sampling straight into one 32-bit word of the persistent chain (bit = position);
//            p (and Q) are scaled by 2^14, split into fp16 hi + lo (exact to 2^-22 relative above 1.5e-5, 2e-12 absolute
//            below) and written back to shared memory as the K-major A operand of GEMM 2.
//   GEMM 2   D2[r, c]   += P[r, pos] . X[pos, c]         the statistics sum_j p[k,j] [base(j+m) = a] (A8), accumulated in
//            TMEM for the whole persistent CTA; B = the SAME one-hot tile, read MN-major (contraction index = position);
//            a second accumulator takes Q (A11). Written once per CTA: partial[c][r].
// The strand fold (convRBM.py:367) is a row/column permutation of D2 applied by the reduction: VH[k,c] = D2[kl,c] + D2[64+kl,4M-1-c].
//
// Warp roles (640 threads, one CTA per SM, persistent over units blockIdx.x, blockIdx.x + gridDim.x, ...):
//   warps 0-1   build the one-hot tiles (ring of up to 8), one lane of warp 2 (elect.sync) issues the MMAs with descriptors advanced by
//               their low words, warp 3 owns the TMEM allocation (four producer warps with table-built rows, as in scan_rt.cuh, were
//               measured: no gain, the kernel is not bound by its producers),
//   warps 4-19  epilogue: two groups of 8 warps work on alternate units; inside a group the two warpgroups take the two
//               32-position halves. D1 has 4 buffers of 64 columns, the P/Q operands 2 sets (one per group).
#pragma once
#include "v2h.cuh"
#include "tc_common.cuh"

namespace crbm {

constexpr int TT_UNIT = 64;                   // hidden positions per unit
constexpr int TT_KC = 64;                     // motifs per CTA
constexpr int TT_ROWS = 128;                  // MMA rows: 2 strands x 64 motifs
constexpr int TT_MAX_XSTAGES = 8;
constexpr int TT_THREADS = 640;
constexpr int TT_D2P_COL = 0, TT_D2Q_COL = 128, TT_D1_COL = 256;      // TMEM columns
constexpr int TT_XATOM = TT_UNIT * 128;       // bytes of one 64-column swizzle atom of a one-hot tile (64 rows)
constexpr int TT_WATOM = TT_ROWS * 128;       // same for the 128-row operands
constexpr int TT_QHI = 0, TT_QLO = 8192, TT_PHI = 16384, TT_PLO = 32768, TT_PSET = 49152;
constexpr float TT_PSCALE = 16384.f;          // 2^14: moves probabilities into the fp16 normal range

struct TrainTcParams {
    const uint32_t *packed;
    int row_words, n_rows, L, Lh;
    const int *batch_index;        // device minibatch cursor or null
    const float *W, *b;
    int K, M, ds;
    int nblk32;                    // 32-position blocks per sequence: ceil(Lh / VP), VP = (32 / P) * P
    int ksteps, n_atoms, Kdim;     // Kdim = 4 M padded to 16; ksteps = Kdim / 16; n_atoms = ceil(Kdim / 64)
    int n_units;                   // ceil(n_rows * nblk32 / 2)
    float *partial;                // [gridDim.y][gridDim.x][part_floats]
    int part_floats;
    uint32_t *bits;                // linear chain planes [2][n_rows][hw][Kpad] (F_SAMPLE)
    int hw, Kpad;
    int xstages, npsets;           // one-hot tiles in the ring (<= 8); P/Q operand sets (2, or 1 when shared memory is short)
    const float *u_fwd, *u_rc;     // injected uniforms [(n*K+k)*(Lh/P) + group] or null
    uint64_t seed;
    uint32_t ctr_hi, iter;
    const unsigned long long *step_dev;
    int pdl_late;
};

__host__ __device__ inline int tt_part_floats(int Kdim) { return Kdim * 192 + 128 + 64; }
__host__ __device__ inline size_t tt_smem_bytes(int n_atoms, int xstages, int npsets) {
    return 1024 + 2 * (size_t)n_atoms * TT_WATOM + (size_t)xstages * n_atoms * TT_XATOM + (size_t)npsets * TT_PSET + 512 + 512;
}

// one-hot row of a 64-row tile (same chunk layout as scan_tc.cu: chunk ch = bases 2ch, 2ch+1 as eight fp16)
__device__ __forceinline__ void tt_build_row(unsigned char *X, int r, int nch, const uint32_t (&w)[4], int sh, int mlim) {
    uint32_t v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
#pragma unroll
    for (int ch = 0; ch < 16; ++ch) {
        if (ch < nch) {
            const uint32_t bits = v[ch >> 3] >> (4 * (ch & 7));
            const int b0 = bits & 3, b1 = (bits >> 2) & 3, m0 = 2 * ch;
            const uint32_t h0 = (m0 < mlim) ? (0x3C00u << (16 * (b0 & 1))) : 0u;
            const uint32_t h1 = (m0 + 1 < mlim) ? (0x3C00u << (16 * (b1 & 1))) : 0u;
            uint4 q;
            q.x = (b0 & 2) ? 0u : h0; q.y = (b0 & 2) ? h0 : 0u;
            q.z = (b1 & 2) ? 0u : h1; q.w = (b1 & 2) ? h1 : 0u;
            *reinterpret_cast<uint4 *>(X + (ch >> 3) * TT_XATOM + sw128_off(r, ch & 7)) = q;
        }
    }
}

constexpr int TT_LUT_BYTES = 21 * 16;         // one-hot patterns of two bases (16 entries), of one base (4), zero (1)
// threads 0-20 of the CTA write the pattern table: fp16 1.0 in the half selected by the low bit of the base, in the word selected by
// its high bit
__device__ __forceinline__ void tt_lut_init(unsigned char *lut, int tid) {
    if (tid < 21) {
        const int b0 = tid & 3, b1 = (tid >> 2) & 3;
        const uint32_t h0 = tid < 20 ? (0x3C00u << (16 * (b0 & 1))) : 0u, h1 = tid < 16 ? (0x3C00u << (16 * (b1 & 1))) : 0u;
        uint4 q;
        q.x = (b0 & 2) ? 0u : h0; q.y = (b0 & 2) ? h0 : 0u;
        q.z = (b1 & 2) ? 0u : h1; q.w = (b1 & 2) ? h1 : 0u;
        *reinterpret_cast<uint4 *>(lut + tid * 16) = q;
    }
}
// One-hot row of a 64-row tile from a table of the 16-byte patterns of two bases (same chunk layout as tt_build_row: chunk ch = bases
// 2 ch, 2 ch + 1 as eight fp16): a shift, a mask, a 16-byte table load and a 16-byte store per chunk instead of ~14 ALU instructions.
// lut: entries 0-15 pattern of (b0 | b1 << 2), 16-19 pattern of b0 alone (last chunk of an odd motif length), 20 zero.
// BATCH: one 64-column atom (eight chunks) at a time, the eight table loads issued together, then the eight stores; chunks beyond nch
// inside an atom that exists are stored as zeros. Written as one load and one store per chunk (!BATCH) the compiler reuses a single
// register quad and the row becomes a chain of 2 x nch shared-memory round trips (1600-2000 cycles per tile, measured): that is the
// better schedule for the kernels bound by their epilogue (fused reductions with the mean, pooled scans: the bursts of the batched
// builder cost them 5-7 %), while the kernels bound by the tile supply gain from the batch (un-pooled dense scan: +5 % of the HBM peak).
template <bool BATCH>
__device__ __forceinline__ void tt_build_row_lut(unsigned char *X, const unsigned char *lut, int r, int nch, int M, const uint32_t (&w)[4], int sh, bool valid) {
    uint32_t v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
    const int nfull = M >> 1;
    unsigned char *row = X + (r >> 3) * 1024 + (r & 7) * 128;
    if (BATCH) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (8 * half < nch) {
                uint4 q[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int ch = 8 * half + c;
                    const uint32_t bits = (v[ch >> 3] >> (4 * (ch & 7))) & 15u;
                    uint32_t e = ch < nfull ? bits : (ch == nfull && (M & 1)) ? 16u + (bits & 3u) : 20u;     // (uniform in ch: M is a kernel constant)
                    e = valid ? e : 20u;
                    q[c] = *reinterpret_cast<const uint4 *>(lut + e * 16);
                }
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4 *>(row + half * TT_XATOM + (((c ^ r) & 7) << 4)) = q[c];
            }
        }
    } else {
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
            if (ch < nch) {
                const uint32_t bits = (v[ch >> 3] >> (4 * (ch & 7))) & 15u;
                uint32_t e = ch < nfull ? bits : (ch == nfull && (M & 1)) ? 16u + (bits & 3u) : 20u;
                e = valid ? e : 20u;
                const uint4 q = *reinterpret_cast<const uint4 *>(lut + e * 16);
                *reinterpret_cast<uint4 *>(row + (ch >> 3) * TT_XATOM + (((ch ^ r) & 7) << 4)) = q;
            }
        }
    }
}

__device__ __forceinline__ void tmem_ld_fence16(uint32_t (&v)[16]) {
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                      "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}

static __device__ __noinline__ uint4 tt_philox(uint4 ctr, uint2 key) { return philox4x32(ctr, key); }

// eight values (already scaled by 2^14) -> fp16 hi chunk (the top 11 significand bits, by truncation: exact in fp16) and fp16 lo
// chunk (the exact remainder, rounded to 11 bits): 3 instructions per value
__device__ __forceinline__ uint32_t tt_pack_half2(float a, float b) {
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ void tt_split8(const float *v, uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
        h[i] = tt_pack_half2(ah, bh);
        l[i] = tt_pack_half2(a - ah, b - bh);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// wait with back-off: a spinning warp competes for issue slots with the warps that do the work
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
        if (ok) break;
        __nanosleep(40);
    }
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return ok != 0;
}

template <int FLAGS, int P>
__global__ void __launch_bounds__(TT_THREADS, 1) v2h_train_tc_kernel(const TrainTcParams p) {
    constexpr bool STAT_P = (FLAGS & F_STAT_P) != 0, STAT_Q = (FLAGS & F_STAT_Q) != 0, SAMPLE = (FLAGS & F_SAMPLE) != 0;
    constexpr bool STAT = STAT_P || STAT_Q;
    constexpr int VP = (32 / P) * P, NG = 32 / P;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const size_t w_bytes = (size_t)p.n_atoms * TT_WATOM, x_bytes = (size_t)p.n_atoms * TT_XATOM;
    unsigned char *Whi = smem, *Wlo = Whi + w_bytes;
    unsigned char *X = Wlo + w_bytes;
    unsigned char *PQ = X + (size_t)p.xstages * x_bytes;
    float *bias_s = reinterpret_cast<float *>(PQ + (STAT ? (size_t)p.npsets * TT_PSET : 0));
    uint64_t *x_full = reinterpret_cast<uint64_t *>(bias_s + TT_ROWS);
    uint64_t *x_free = x_full + TT_MAX_XSTAGES;
    uint64_t *d1_full = x_free + TT_MAX_XSTAGES;
    uint64_t *d1_free = d1_full + 4;
    uint64_t *p_full = d1_free + 4;
    uint64_t *p_free = p_full + 2;
    uint64_t *d2_done = p_free + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d2_done + 1);
    const int XS = p.xstages;
    // scale of the accumulators: P == 1 evaluates 1 / (1 + 2^x) with x = -log2(e) a; the pooled softmax needs 2^x with x = +log2(e) a
    constexpr float WSCALE = P == 1 ? -1.4426950408889634f : 1.4426950408889634f;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int chunk = blockIdx.y, G = gridDim.x;
    const int Kdim = p.Kdim;

    if (warp == 3) tmem_alloc(tmem_slot, 512u);
    if (tid == 64) {
        for (int i = 0; i < TT_MAX_XSTAGES; ++i) { mbar_init(x_full + i, 64); mbar_init(x_free + i, 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(d1_full + i, 1); mbar_init(d1_free + i, 8); }
        for (int i = 0; i < 2; ++i) { mbar_init(p_full + i, 8); mbar_init(p_free + i, 1); }
        mbar_init(d2_done, 1);
    }
    // ---- A of GEMM 1: both strands of the chunk's motifs, scaled, fp16 hi / lo, K-major SWIZZLE_128B; one 16-byte chunk
    // (motif positions 2 ch, 2 ch + 1, four bases each) per work item
    for (int idx = tid; idx < TT_ROWS * (Kdim >> 3); idx += TT_THREADS) {
        const int r = idx & (TT_ROWS - 1), ch = idx >> 7;
        const int s = r >> 6, k = chunk * TT_KC + (r & 63);
        float w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int m = 2 * ch + (e >> 2), a = e & 3;
            float v = 0.f;
            if (k < p.K && m < p.M && (s == 0 || p.ds))
                v = s == 0 ? __ldg(p.W + ((size_t)k * 4 + a) * p.M + m) : __ldg(p.W + ((size_t)k * 4 + (3 - a)) * p.M + (p.M - 1 - m));
            w[e] = v * WSCALE;
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
        const uint32_t off = (uint32_t)(ch >> 3) * TT_WATOM + sw128_off(r, ch & 7);
        *reinterpret_cast<uint4 *>(Whi + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
        *reinterpret_cast<uint4 *>(Wlo + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    }
    for (int r = tid; r < TT_ROWS; r += TT_THREADS) {
        const int k = chunk * TT_KC + (r & 63);
        bias_s[r] = k < p.K ? WSCALE * __ldg(p.b + k) : 0.f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_units = ((int)blockIdx.x < p.n_units) ? (p.n_units - 1 - (int)blockIdx.x) / G + 1 : 0;
    const uint32_t *packed = p.packed + (p.batch_index ? (size_t)(*p.batch_index) * p.n_rows * p.row_words : 0);
    // programmatic dependent launch: see v2h_mma.cuh (W, b and the cursors were complete before this grid could start)
    if (!p.pdl_late) pdl_wait();
    pdl_launch_dependents();

    if (warp < 2) {
        // ================================================= producers: one-hot tiles ========================================
        const int nch = 2 * p.ksteps;
        auto load_row = [&](int t, uint32_t (&w)[4], int &sh, int &mlim) {
            const int ug = (int)blockIdx.x + t * G;
            const int b32 = 2 * ug + (tid >> 5), i = tid & 31;
            const int n = b32 / p.nblk32, wb = b32 - n * p.nblk32;
            const int j = wb * VP + i;
            mlim = 0; sh = 0;
            w[0] = w[1] = w[2] = w[3] = 0u;
            if (n < p.n_rows && i < VP && j < p.Lh) {
                const int bit = 2 * (j + 16 * FRONT_PAD_WORDS);
                const uint32_t *src = packed + (size_t)n * p.row_words + (bit >> 5);
                sh = bit & 31;
#pragma unroll
                for (int e = 0; e < 4; ++e) w[e] = __ldg(src + e);
                mlim = p.M;
            }
        };
        uint32_t cw[4], nw[4];
        int csh = 0, cml = 0, nsh = 0, nml = 0;
        if (my_units > 0) load_row(0, cw, csh, cml);
        for (int t = 0; t < my_units; ++t) {
            const int stage = t % XS, use = t / XS;
            if (t + 1 < my_units) load_row(t + 1, nw, nsh, nml);
            if (use > 0) mbar_wait(x_free + stage, (uint32_t)((use - 1) & 1));
            tt_build_row(X + (size_t)stage * x_bytes, tid, nch, cw, csh, cml);
            fence_async_smem();
            mbar_arrive(x_full + stage);
#pragma unroll
            for (int e = 0; e < 4; ++e) cw[e] = nw[e];
            csh = nsh; cml = nml;
        }
    } else if (warp == 2) {
        // ================================================= MMA issue (one thread) ==========================================
        if (elect_one()) {
            const uint32_t whi_lo = umma_desc_lo(smem_u32(Whi)), wlo_lo = umma_desc_lo(smem_u32(Wlo)), xa = smem_u32(X), pq = smem_u32(PQ);
            const uint32_t idesc1 = (1u << 4) | ((uint32_t)(TT_UNIT >> 3) << 17) | ((uint32_t)(TT_ROWS >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (1u << 16) | ((uint32_t)(Kdim >> 3) << 17) | ((uint32_t)(TT_ROWS >> 4) << 24);
            // Polling issue loop: GEMM 2 of the oldest finished unit first (it frees a one-hot stage and an operand set), else the
            // next GEMM 1 whose tile is built and whose accumulator buffer has been drained. Nothing blocks behind a wait.
            const int npm = p.npsets - 1;                         // 0 or 1
            int g1 = 0, g2 = 0;
            const int g2_total = STAT ? my_units : 0;
            while (g1 < my_units || g2 < g2_total) {
                bool progressed = false;
                if (STAT && g2 < g1) {
                    const int set_i = g2 & npm, use = p.npsets == 2 ? (g2 >> 1) : g2;
                    if (mbar_test(p_full + set_i, (uint32_t)(use & 1))) {
                        const int t = g2, stage = t % XS;
                        tc_fence_after();
                        const uint32_t set_lo = umma_desc_lo(pq + set_i * TT_PSET);
                        // B = the one-hot tile read MN-major: 16 positions = two 8-row groups (1024 bytes apart), c atoms TT_XATOM apart
                        const uint32_t xmn_lo = umma_desc_lo(xa + (uint32_t)(stage * x_bytes), TT_XATOM);
#pragma unroll
                        for (int k2 = 0; k2 < TT_UNIT / 16; ++k2) {
                            const uint32_t bd = xmn_lo + (uint32_t)(k2 * 2048 >> 4), oa = (uint32_t)(k2 * 32 >> 4);
                            const uint32_t acc = (t > 0 || k2 > 0) ? 1u : 0u;
                            umma_f16_lo(tmem_base + TT_D2P_COL, set_lo + (TT_PHI >> 4) + oa, bd, idesc2, acc);
                            umma_f16_lo(tmem_base + TT_D2P_COL, set_lo + (TT_PLO >> 4) + oa, bd, idesc2, 1u);
                            if (STAT_Q) {
                                // rows 64-127 of the Q operands alias the buffers that follow them (finite values; those D2 lanes are unused)
                                umma_f16_lo(tmem_base + TT_D2Q_COL, set_lo + (TT_QHI >> 4) + oa, bd, idesc2, acc);
                                umma_f16_lo(tmem_base + TT_D2Q_COL, set_lo + (TT_QLO >> 4) + oa, bd, idesc2, 1u);
                            }
                        }
                        umma_commit(x_free + stage);
                        // "the set may be overwritten": with two sets the waiter is the same group's next unit; with one shared set it is
                        // the OTHER group's next unit, so each group has its own barrier and sees one phase at a time (a parity wait
                        // cannot tell phases two apart: scratch/tt_sim.py)
                        umma_commit(p_free + (p.npsets == 2 ? set_i : ((t + 1) & 1)));
                        ++g2;
                        progressed = true;
                    }
                }
                if (g1 < my_units) {
                    const int t = g1, stage = t % XS, buf = t & 3;
                    if (mbar_test(x_full + stage, (uint32_t)((t / XS) & 1)) &&
                        (t < 4 || mbar_test(d1_free + buf, (uint32_t)(((t >> 2) - 1) & 1)))) {
                        tc_fence_after();
                        const uint32_t d1 = tmem_base + TT_D1_COL + 64 * buf;
                        const uint32_t xs = xa + (uint32_t)(stage * x_bytes);
                        const uint32_t x_lo = umma_desc_lo(xs);
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            if (s < p.ksteps) {
                                const uint32_t oa = (uint32_t)(((s >> 2) * TT_WATOM + (s & 3) * 32) >> 4), ox = (uint32_t)(((s >> 2) * TT_XATOM + (s & 3) * 32) >> 4);
                                umma_f16_lo(d1, whi_lo + oa, x_lo + ox, idesc1, s > 0 ? 1u : 0u);
                                umma_f16_lo(d1, wlo_lo + oa, x_lo + ox, idesc1, 1u);
                            }
                        }
                        umma_commit(d1_full + buf);
                        if (!STAT) umma_commit(x_free + stage);
                        ++g1;
                        progressed = true;
                    }
                }
                if (!progressed) __nanosleep(20);
            }
            umma_commit(d2_done);
        }
    } else if (warp >= 4) {
        // ================================================= epilogue ========================================================
        // All probabilities are carried SCALED by 2^14 (pv = 2^14 p, qv = 2^14 Q): the scale is folded into the reciprocal, the
        // fp16 operands of GEMM 2 need it, and powers of two commute with every comparison and sum below.
        const int ew = warp - 4, q = warp & 3, wg = ew >> 2, pr = wg >> 1, h = wg & 1;
        const int r = 32 * q + lane, s = q >> 1, kl = r & 63, k = chunk * TT_KC + kl;
        const bool row_ok = k < p.K && (s == 0 || p.ds);
        const float bias = bias_s[r];
        const int groups_per_seq = p.Lh / P;
        const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        const uint32_t ctr_hi = p.ctr_hi + (p.step_dev ? (uint32_t)(*p.step_dev) : 0u);
        const uint32_t tag = (s == 0 ? RNG_HIDDEN : RNG_HIDDEN_RC) | (p.iter << 8);
        const float *inj = s == 0 ? p.u_fwd : p.u_rc;
        const int npm = p.npsets - 1;
        constexpr float INV_S = 1.0f / TT_PSCALE;
        float sum_p = 0.f, sum_q = 0.f;
        for (int t = pr; t < my_units; t += 2) {
            const int ug = (int)blockIdx.x + t * G;
            const int b32 = 2 * ug + h;
            const int n = b32 / p.nblk32, wb = b32 - n * p.nblk32;
            int nv = n < p.n_rows ? min(VP, p.Lh - wb * VP) : 0;
            nv = nv < 0 ? 0 : nv;
            const int buf = t & 3;
            mbar_wait_backoff(d1_full + buf, (uint32_t)((t >> 2) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + TT_D1_COL + 64 * buf + 32 * h;
            uint32_t va[16], vb[16];
            tmem_ld16(taddr, va);
            tmem_ld16(taddr + 16, vb);
            tmem_ld_wait();
            tmem_ld_fence16(va);
            tmem_ld_fence16(vb);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d1_free + buf);          // the accumulator buffer is in registers now
            float pv[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) { pv[i] = __uint_as_float(va[i]) + bias; pv[16 + i] = __uint_as_float(vb[i]) + bias; }
            float qv[STAT_Q ? 32 : 1];
            if (P == 1) {
                // 2^14 p = 2^14 / (1 + 2^x), x = -log2(e) (activity + b). Two outputs share one reciprocal (MUFU is a busy pipe):
                // a_i = 2^-14 (1 + e_i) (one FFMA), r = 1 / (a_0 a_1), 2^14 p_0 = r a_1, 2^14 p_1 = r a_0.
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float e0, e1, rr;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(pv[i], 40.f)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(pv[i + 1], 40.f)));
                    const float a0 = fmaf(e0, INV_S, INV_S), a1 = fmaf(e1, INV_S, INV_S);
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr) : "f"(a0 * a1));
                    pv[i] = rr * a1;
                    pv[i + 1] = rr * a0;
                }
                if (nv != 32) {                               // warp-uniform: the last block of a sequence, the tail of the input
#pragma unroll
                    for (int i = 0; i < 32; ++i) pv[i] = i < nv ? pv[i] : 0.f;
                }
                if constexpr (STAT_Q) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) qv[i] = fmaf(-(pv[i] * INV_S), pv[i], pv[i]);      // 2^14 p (1 - p)
                }
            } else {
                // p = exp(a) / (P + sum_group exp(a)) (convRBM.py:252), a clamped at 40; groups are P consecutive registers
                constexpr float fpool = (float)P;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(pv[i], 57.7f)));
                    pv[i] = i < VP ? e : 0.f;
                }
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float tot = 0.f;
#pragma unroll
                    for (int i = 0; i < P; ++i) tot += pv[g * P + i];
                    float inv;                                   // 2^14 / (P + tot)
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaf(tot, INV_S, fpool * INV_S)));
                    const float gq = inv * (fpool * INV_S);      // P / (P + tot) = 1 - sum_group p
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        pv[g * P + i] *= inv;
                        if constexpr (STAT_Q) qv[g * P + i] = pv[g * P + i] * gq;
                    }
                }
                if constexpr (STAT_Q) {
#pragma unroll
                    for (int i = VP; i < 32; ++i) qv[i] = 0.f;
                }
                if (nv != VP) {                               // warp-uniform; nv is a multiple of P: whole groups drop out
#pragma unroll
                    for (int i = 0; i < VP; ++i) {
                        pv[i] = i < nv ? pv[i] : 0.f;
                        if constexpr (STAT_Q) qv[i] = i < nv ? qv[i] : 0.f;
                    }
                }
            }
            // ---- hidden sums (A9), per unit first so that the running sums grow by a few dozen terms at a time
            {
                float us0 = 0.f, us1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 2) { us0 += pv[i]; us1 += pv[i + 1]; }
                sum_p += us0 + us1;
                if constexpr (STAT_Q) {
                    float uq0 = 0.f, uq1 = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) { uq0 += qv[i]; uq1 += qv[i + 1]; }
                    sum_q += uq0 + uq1;
                }
            }
            // ---- sampling (A3): the 32 positions of this (strand, motif, block) are one word of the chain planes
            if (SAMPLE) {
                uint32_t word = 0u;
                const uint32_t rowid = (uint32_t)(n * p.K + (row_ok ? k : 0));
                if (P == 1) {
                    if (inj) {
                        const float *up = inj + (size_t)rowid * p.Lh + wb * 32;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float u = (row_ok && i < nv) ? up[i] : 1.f;
                            word |= (pv[i] > u * TT_PSCALE ? 1u : 0u) << i;
                        }
                    } else {
                        // 16-bit uniforms, eight per Philox draw; exact Bernoulli(p) through a second draw on the rare ties
                        // (same scheme as v2h_mma.cuh); 65536 p = 4 pv
                        uint32_t ambw = 0u;
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const uint4 rnd = tt_philox(make_uint4(rowid, (uint32_t)(wb * 4 + c4), ctr_hi, tag), key);
                            const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                // floor(65536 p) against the 16 bits, both in the pattern of a float near 2^23 (floor_bits, common.cuh)
                                const uint32_t tb = floor_bits(pv[8 * c4 + i], 4.f);
                                const uint32_t ub = 0x4B000000u | ((rw[i >> 1] >> (16 * (i & 1))) & 0xffffu);
                                word |= (tb > ub ? 1u : 0u) << (8 * c4 + i);
                                ambw |= (tb == ub ? 1u : 0u) << (8 * c4 + i);
                            }
                        }
                        if (__any_sync(0xffffffffu, ambw != 0u)) {
#pragma unroll
                            for (int c4 = 0; c4 < 4; ++c4) {
                                const uint4 r2 = tt_philox(make_uint4(rowid, (uint32_t)(wb * 4 + c4), ctr_hi, tag + 16u), key);
                                const uint32_t r2w[4] = {r2.x, r2.y, r2.z, r2.w};
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float tt = pv[8 * c4 + i] * 4.f;
                                    const float fr = tt - (float)(int)tt;
                                    const bool hit2 = fr * 65536.f > (float)((r2w[i >> 1] >> (16 * (i & 1))) & 0xffffu);
                                    if ((ambw >> (8 * c4 + i)) & 1u) word |= (hit2 ? 1u : 0u) << (8 * c4 + i);
                                }
                            }
                        }
                    }
                } else {
                    // one uniform per pool group (convRBM.py:259-267): first unit whose running sum exceeds it, none if u >= sum
                    float ug4[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        float u;
                        if (inj) {
                            u = (row_ok && g * P < nv) ? inj[(size_t)rowid * groups_per_seq + wb * NG + g] : 1.f;
                        } else {
                            if ((g & 3) == 0) {
                                const uint4 rnd = tt_philox(make_uint4(rowid, (uint32_t)(wb * 4 + (g >> 2)), ctr_hi, tag), key);
                                ug4[0] = u01(rnd.x); ug4[1] = u01(rnd.y); ug4[2] = u01(rnd.z); ug4[3] = u01(rnd.w);
                            }
                            u = ug4[g & 3];
                        }
                        const float us = u * TT_PSCALE;
                        float run = 0.f;
                        bool done = false;
#pragma unroll
                        for (int i = 0; i < P; ++i) {
                            run += pv[g * P + i];
                            const bool hit = !done && run > us;
                            done = done || hit;
                            word |= (hit ? 1u : 0u) << (g * P + i);
                        }
                    }
                }
                if (n < p.n_rows) p.bits[(((size_t)s * p.n_rows + n) * p.hw + wb) * p.Kpad + chunk * TT_KC + kl] = row_ok ? word : 0u;
            }
            // ---- A operands of GEMM 2: 2^14 p (and 2^14 Q) as fp16 hi / lo, K-major SWIZZLE_128B rows of 64 positions
            if (STAT) {
                const int set_i = t & npm, use = p.npsets == 2 ? (t >> 1) : t;
                unsigned char *set = PQ + set_i * TT_PSET;
                // the GEMM 2 that read this set last has completed: unit t-2 (own set) or unit t-1 (shared set, own barrier)
                if (p.npsets == 2) { if (use >= 1) mbar_wait_backoff(p_free + set_i, (uint32_t)((use - 1) & 1)); }
                else if (t >= 1) mbar_wait_backoff(p_free + pr, (uint32_t)(((t - 1) >> 1) & 1));
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    uint4 hi, lo;
                    tt_split8(pv + 8 * c8, hi, lo);
                    const uint32_t off = sw128_off(r, 4 * h + c8);
                    *reinterpret_cast<uint4 *>(set + TT_PHI + off) = hi;
                    *reinterpret_cast<uint4 *>(set + TT_PLO + off) = lo;
                }
                if constexpr (STAT_Q) if (s == 0) {
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        uint4 hi, lo;
                        tt_split8(qv + 8 * c8, hi, lo);
                        const uint32_t off = sw128_off(kl, 4 * h + c8);
                        *reinterpret_cast<uint4 *>(set + TT_QHI + off) = hi;
                        *reinterpret_cast<uint4 *>(set + TT_QLO + off) = lo;
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full + set_i);
            }
        }
        // ---- statistics accumulators -> this CTA's partial: P[c][r], Q[c][kl], hidden sums per warpgroup ----------------------
        if (STAT) {
            float *part = p.partial + ((size_t)chunk * G + blockIdx.x) * p.part_floats;
            mbar_wait_backoff(d2_done, 0u);
            tc_fence_after();
            const float unscale = 1.0f / TT_PSCALE;
            const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16);
            for (int cc = wg; cc < (Kdim >> 3); cc += 4) {
                uint32_t v[8];
                tmem_ld8(trow + TT_D2P_COL + 8 * cc, v);
                tmem_ld_wait();
                tmem_ld_fence8(v);
#pragma unroll
                for (int j = 0; j < 8; ++j) part[(size_t)(8 * cc + j) * TT_ROWS + r] = my_units > 0 ? __uint_as_float(v[j]) * unscale : 0.f;
                if (STAT_Q && s == 0) {
                    tmem_ld8(trow + TT_D2Q_COL + 8 * cc, v);
                    tmem_ld_wait();
                    tmem_ld_fence8(v);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        part[(size_t)Kdim * TT_ROWS + (size_t)(8 * cc + j) * TT_KC + kl] = my_units > 0 ? __uint_as_float(v[j]) * unscale : 0.f;
                }
            }
            // hidden sums: the four warpgroups hold partial sums of the same rows; combined through shared memory (the operand sets
            // are free: every GEMM 2 has completed) in a fixed order
            float *comb = reinterpret_cast<float *>(PQ);
            comb[wg * TT_ROWS + r] = row_ok ? sum_p : 0.f;       // rows beyond K / of the unused strand hold values of no meaning
            if (s == 0) comb[4 * TT_ROWS + wg * TT_KC + kl] = row_ok ? sum_q : 0.f;
            asm volatile("bar.sync 2, 512;" ::: "memory");
            if (wg == 0) {
                float *sums = part + (size_t)Kdim * 192;
                sums[r] = ((comb[r] + comb[TT_ROWS + r]) + (comb[2 * TT_ROWS + r] + comb[3 * TT_ROWS + r])) * unscale;
                if (s == 0) {
                    const float *cq = comb + 4 * TT_ROWS;
                    sums[TT_ROWS + kl] = STAT_Q ? ((cq[kl] + cq[TT_KC + kl]) + (cq[2 * TT_KC + kl] + cq[3 * TT_KC + kl])) * unscale : 0.f;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc(tmem_base, 512u);
    if (p.pdl_late) pdl_wait();
}

template <int P> int launch_v2h_train_tc(int flags, const TrainTcParams &p, dim3 grid, cudaStream_t stream);

}  // namespace crbm
