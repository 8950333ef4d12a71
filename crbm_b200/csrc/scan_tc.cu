// Motif scan (getHitProbs / motifHitProbs, reference secomo/convRBM.py:238-257,507-514) on the 5th-generation tensor
// cores: tcgen05.mma with the accumulator in tensor memory.
//
//   x[n,k,j] = b_k + sum_{m,a} W[k,a,m] [base(n,j+m) == a]            (A1, gather form of the valid cross-correlation)
//
// is the GEMM  D[pos, k] = X[pos, c] * Wt[k, c]^T,  c = 4m + a,  X = one-hot of the M bases under hidden position pos.
//   * A (M = 128 hidden positions x Kdim = 4M padded to 16): the one-hot rows are exact in fp16 and are never read from
//     HBM: the CTA builds them in shared memory from the packed 2-bit sequences (two 64-bit shifts per 16-byte chunk),
//     directly in the K-major SWIZZLE_128B layout the MMA descriptor expects.
//   * B (N = up to 256 motifs x Kdim): W is split once per CTA into two fp16 terms, W = hi + lo (11 + 11 significand
//     bits, |W - hi - lo| <= 2^-22 |W| or 3e-8 absolute: fp32 accuracy with two tensor-core passes instead of three
//     bf16 ones). The reverse-complement strand is just another B (W[k,3-a,M-1-m], convRBM.py:241 filter_flip=True), the
//     sum of both strands (getHitProbs with doublestranded=False, :511-514) the sum of the two. For probabilities B and
//     the bias are pre-scaled by -log2(e), so the epilogue is ex2(acc + bias) -> rcp(1 + e).
//   * D (128 lanes x N fp32 columns) lives in TMEM, double buffered: the MMAs of tile i+1 run while all eight warps drain
//     tile i (tcgen05.ld 32x32b: a thread holds one position and 16 consecutive motifs, so for every motif a warp writes
//     32 consecutive floats of one output row = one 128-byte store).
// Tiles are 128 consecutive hidden positions of the flattened (sequence, position) index, so every MMA row is used and
// the work per tile is uniform; a thread finds its (n, j) with one division per tile.
// One CTA per SM slot (two per SM when N <= 128), persistent over tiles, one __syncthreads per tile.
// Bound: HBM write (4*K bytes per hidden position); tensor pipe ~20-30 % busy, MUFU (two per output) ~70 %.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "v2h.cuh"
#include "tc_common.cuh"

namespace crbm {

struct TcParams {
    DenseParams d;
    int NT;            // motifs per CTA (multiple of 16, <= 256)
    int ksteps;        // Kdim / 16
    int n_atoms;       // ceil(Kdim / 64): 128-byte swizzle atoms along K
    int tmem_stride;   // columns between accumulator buffers (power of two >= max(NT, 128))
    int nbuf;          // accumulator buffers in the 512 TMEM columns: 4 (NT <= 128) or 2
    int a_stages;      // one-hot A tiles in the shared-memory ring (2 or 3)
    int store_len;     // positions written per row: Lh, or the padded row (aligned mode: the pad floats are written too, so
                       // that no 32-byte sector is left partially written -> no read-modify-write in DRAM)
    int n_tiles;       // ceil(n_rows * LhV / 128)
    float scale;       // -log2(e) for probabilities, 1 for activities
    int LhV;           // virtual hidden length the tiles are cut from: Lh, or Lh rounded up to 32 when the output rows are
                       // sector aligned (then every warp of a tile sits inside one sequence at a multiple of 32 positions)
    long long row_stride;   // floats between consecutive (n,k) rows of the output (Lh for the reference's contiguous layout)
    int P;             // pooling factor of the probabilities (1: sigmoid)
    int VP;            // real positions per warp of a tile: 32, or (32 / P) * P so that pool groups never straddle warps. Virtual
                       // position v of a sequence <-> real position (v / 32) * VP + v % 32, valid iff v % 32 < VP
};

constexpr int TC_TILE = 128;
constexpr int A_ATOM_BYTES = TC_TILE * 128;
constexpr int TC_PRODUCER_WARPS = 4;                      // build the one-hot A tiles, one thread issues the MMAs
constexpr int TC_EPILOGUE_WARPS = 16;                     // four warpgroups draining accumulator buffers
constexpr int TC_THREADS = 32 * (TC_PRODUCER_WARPS + TC_EPILOGUE_WARPS);
constexpr int TC_TMEM_COLS = 512;                         // the CTA owns the SM: all of tensor memory
constexpr int TC_MAX_A_STAGES = 3;

__host__ __device__ inline size_t tc_smem_bytes(int NT, int n_atoms, int a_stages) {
    return 1024 /*alignment slack*/ + (size_t)a_stages * n_atoms * A_ATOM_BYTES + 2 * (size_t)n_atoms * NT * 128 + (size_t)NT * 4 + 256;
}

// (sequence, hidden position) of the first row of a tile; tiles advance by a fixed stride, so no 64-bit division is needed
struct TilePos {
    int n0, j0;
    __device__ __forceinline__ void init(int first_pos, int Lh) { n0 = first_pos / Lh; j0 = first_pos - n0 * Lh; }
    __device__ __forceinline__ void advance(int positions, int Lh) {
        const int jj = j0 + positions;
        const int dn = jj / Lh;
        n0 += dn;
        j0 = jj - dn * Lh;
    }
    __device__ __forceinline__ void row(int r, int Lh, int &n, int &j) const {
        const int jj = j0 + r, dn = jj / Lh;
        n = n0 + dn;
        j = jj - dn * Lh;
    }
    // same with the virtual -> real position map of pooled tiles (TcParams::VP); ok = the row is a real position of its warp
    __device__ __forceinline__ void row(int r, int Lh, int VP, int &n, int &j, bool &ok) const {
        row(r, Lh, n, j);
        ok = (j & 31) < VP;
        j = (j >> 5) * VP + (j & 31);
    }
};

// the packed words under hidden position row `r` of a tile: three 16-base windows starting at the position's first base
struct RowBases {
    uint32_t w[4];
    int sh, mlim;
    __device__ __forceinline__ void load(const TcParams &p, const TilePos &tp, int r) {
        const DenseParams &d = p.d;
        int n, j;
        bool ok;
        tp.row(r, p.LhV, p.VP, n, j, ok);
        mlim = 0; sh = 0;
        w[0] = w[1] = w[2] = w[3] = 0u;
        if (ok && n < d.n_rows && j < d.Lh) {
            const int bit = 2 * (j + 16 * FRONT_PAD_WORDS);
            const uint32_t *src = d.packed + (size_t)n * d.row_words + (bit >> 5);
            sh = bit & 31;
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = __ldg(src + i);
            mlim = d.M;
        }
    }
};

// One-hot row `r` of an A tile: 16-byte chunk ch holds the eight fp16 [base(j+2ch) == 0..3], [base(j+2ch+1) == 0..3].
__device__ __forceinline__ void tc_build_row(const TcParams &p, unsigned char *A, int r, const RowBases &rb) {
    uint32_t v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = __funnelshift_r(rb.w[i], rb.w[i + 1], rb.sh);     // bases j+16i .. j+16i+15
    const int nch = 2 * p.ksteps;
#pragma unroll
    for (int ch = 0; ch < 18; ++ch) {
        if (ch < nch) {
            const uint32_t bits = v[ch >> 3] >> (4 * (ch & 7));
            const int b0 = bits & 3, b1 = (bits >> 2) & 3, m0 = 2 * ch;
            const uint32_t h0 = (m0 < rb.mlim) ? (0x3C00u << (16 * (b0 & 1))) : 0u;   // fp16 1.0 in the half selected by the base
            const uint32_t h1 = (m0 + 1 < rb.mlim) ? (0x3C00u << (16 * (b1 & 1))) : 0u;
            uint4 q;
            q.x = (b0 & 2) ? 0u : h0; q.y = (b0 & 2) ? h0 : 0u;
            q.z = (b1 & 2) ? 0u : h1; q.w = (b1 & 2) ? h1 : 0u;
            *reinterpret_cast<uint4 *>(A + (ch >> 3) * A_ATOM_BYTES + sw128_off(r, ch & 7)) = q;
        }
    }
}

__device__ __forceinline__ void tc_issue_mma(const TcParams &p, uint32_t a_addr, uint32_t bhi_addr, uint32_t blo_addr,
                                             uint32_t tmem_d, uint32_t idesc) {
    const uint32_t b_atom = ((uint32_t)p.NT * 128u) >> 4;
    const uint32_t a_lo = umma_desc_lo(a_addr), bhi_lo = umma_desc_lo(bhi_addr), blo_lo = umma_desc_lo(blo_addr);
#pragma unroll
    for (int s = 0; s < 9; ++s) {
        if (s < p.ksteps) {
            const uint32_t atom = s >> 2, koff = ((s & 3) * 32) >> 4;     // 16 fp16 = 32 bytes per k-step inside the 128-byte row
            const uint32_t ad = a_lo + atom * (A_ATOM_BYTES >> 4) + koff;
            umma_f16_lo(tmem_d, ad, bhi_lo + atom * b_atom + koff, idesc, s > 0 ? 1u : 0u);
            umma_f16_lo(tmem_d, ad, blo_lo + atom * b_atom + koff, idesc, 1u);
        }
    }
}
__device__ __forceinline__ void producer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_PRODUCER_WARPS) : "memory"); }

// Warp-specialised, persistent, one CTA per SM:
//   warps 0-3   producers: thread r builds row r of the next one-hot A tile (ring of a_stages tiles; the packed words of the
//               tile after that are already in flight), then thread 0 issues the tile's MMAs into accumulator buffer i % nbuf
//               and commits to full[buf] and a_free[stage]
//   warps 4-19  four epilogue warpgroups: group g drains buffer g % nbuf (when nbuf < 4 the groups sharing a buffer take
//               alternate 8-motif chunks): wait full[buf] -> tcgen05.ld -> bias, sigmoid -> 128-byte row stores -> arrive empty[buf]
// Sum of v over the pool group of P consecutive lanes that contains this lane (groups start at lanes that are multiples of P;
// gs = this lane's group start). Window sums of length 1, 2, 4, ... by doubling, P composed from its binary digits, then the
// value of the group's first lane is broadcast: popcount(P) + floor(log2 P) shuffles instead of P.
__device__ __forceinline__ float pool_group_sum(float v, int P, int gs) {
    float w[5];
    w[0] = v;
#pragma unroll
    for (int b = 1; b < 5; ++b) {
        w[b] = w[b - 1];
        if ((1 << b) <= P) w[b] = w[b - 1] + __shfl_down_sync(0xffffffffu, w[b - 1], 1 << (b - 1));     // warp-uniform branch
    }
    float acc = 0.f;
    int off = 0;
#pragma unroll
    for (int b = 4; b >= 0; --b) {
        if (P & (1 << b)) {
            acc += off ? __shfl_down_sync(0xffffffffu, w[b], off) : w[b];
            off += 1 << b;
        }
    }
    return __shfl_sync(0xffffffffu, acc, gs);
}

template <bool PROB, bool POOL>
__global__ void __launch_bounds__(TC_THREADS, 1) v2h_dense_tc_kernel(const TcParams p) {
    const DenseParams &d = p.d;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const size_t a_tile_bytes = (size_t)p.n_atoms * A_ATOM_BYTES;
    unsigned char *A = smem;
    unsigned char *Bhi = A + (size_t)p.a_stages * a_tile_bytes;
    unsigned char *Blo = Bhi + (size_t)p.n_atoms * p.NT * 128;
    float *bias_s = reinterpret_cast<float *>(Blo + (size_t)p.n_atoms * p.NT * 128);
    uint64_t *full = reinterpret_cast<uint64_t *>(bias_s + p.NT);       // NT multiple of 16 -> 8-byte aligned
    uint64_t *empty = full + 4;
    uint64_t *a_free = empty + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_free + TC_MAX_A_STAGES);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);         // warp-uniform for the compiler: uniform registers / branches
    const int kbase = blockIdx.y * p.NT;
    const int Kdim = 16 * p.ksteps;
    const int nbuf = p.nbuf, groups_per_buf = 4 / nbuf;

    if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)TC_TMEM_COLS);
    if (tid == 32) {
        for (int i = 0; i < 4; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 4 * groups_per_buf); }
        for (int i = 0; i < TC_MAX_A_STAGES; ++i) mbar_init(a_free + i, 1);
    }

    // ---- B: W (strand-selected, scaled) -> fp16 hi / lo, K-major SWIZZLE_128B ------------------------------------------------
    for (int idx = tid; idx < p.NT * Kdim; idx += TC_THREADS) {
        const int nrow = idx / Kdim, c = idx - nrow * Kdim;
        const int k = kbase + nrow, m = c >> 2, a = c & 3;
        float w = 0.f;
        if (k < d.K && m < d.M) {
            if (d.strand_mode != CRBM_STRAND_RC) w += __ldg(d.W + ((size_t)k * 4 + a) * d.M + m);
            if (d.strand_mode != CRBM_STRAND_FWD) w += __ldg(d.W + ((size_t)k * 4 + (3 - a)) * d.M + (d.M - 1 - m));
        }
        w *= p.scale;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const uint32_t off = (uint32_t)(c >> 6) * (uint32_t)p.NT * 128u + sw128_off(nrow, (c >> 3) & 7) + (c & 7) * 2;
        *reinterpret_cast<__half *>(Bhi + off) = hi;
        *reinterpret_cast<__half *>(Blo + off) = lo;
    }
    for (int i = tid; i < p.NT; i += TC_THREADS) {
        const int k = kbase + i;
        const float bs = d.strand_mode == CRBM_STRAND_SUM ? 2.f : 1.f;       // both activities carry the bias (convRBM.py:513)
        bias_s[i] = k < d.K ? bs * p.scale * __ldg(d.b + k) : 0.f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int stride_pos = TC_TILE * (int)gridDim.x;                  // positions between consecutive tiles of this CTA
    // number of tiles this CTA processes: tiles blockIdx.x, blockIdx.x + gridDim.x, ...
    const int my_tiles = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp < TC_PRODUCER_WARPS) {
        // ================================================= producers ===================================================
        const uint32_t a_addr = smem_u32(A), bhi_addr = smem_u32(Bhi), blo_addr = smem_u32(Blo);
        // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B fp16, both K-major, N = NT, M = 128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);
        TilePos tp;
        tp.init((int)blockIdx.x * TC_TILE, p.LhV);
        RowBases cur, nxt;
        if (my_tiles > 0) cur.load(p, tp, tid);
        for (int i = 0; i < my_tiles; ++i) {
            const int stage = i % p.a_stages, use = i / p.a_stages;
            tp.advance(stride_pos, p.LhV);
            if (i + 1 < my_tiles) nxt.load(p, tp, tid);                              // in flight while this tile is built
            if (use > 0) mbar_wait(a_free + stage, (uint32_t)((use - 1) & 1));       // the MMAs that read this stage are done
            tc_build_row(p, A + (size_t)stage * a_tile_bytes, tid, cur);
            fence_async_smem();
            producer_bar_sync();
            if (warp == 0 && elect_one()) {
                const int buf = i % nbuf, u = i / nbuf;
                if (u > 0) mbar_wait(empty + buf, (uint32_t)((u - 1) & 1));          // the epilogue has drained this accumulator
                tc_fence_after();
                tc_issue_mma(p, a_addr + (uint32_t)(stage * a_tile_bytes), bhi_addr, blo_addr,
                             tmem_base + (uint32_t)(buf * p.tmem_stride), idesc);
                umma_commit(full + buf);
                umma_commit(a_free + stage);
            }
            cur = nxt;
        }
    } else {
        // ================================================= epilogue ====================================================
        const int ew = warp - TC_PRODUCER_WARPS;           // 0..15
        const int q = warp & 3;                            // TMEM lane quarter this warp may read: lanes 32q .. 32q+31
        const int grp = ew >> 2;                           // warpgroup 0..3
        const int buf = grp % nbuf, sub = grp / nbuf;      // accumulator buffer and chunk phase of this group
        const size_t stride8 = 8 * (size_t)p.row_stride;
        const int kvalid = min(p.NT, d.K - kbase);
        const int nch8 = (kvalid + 7) >> 3;
        const uint32_t trow = tmem_base + (uint32_t)(buf * p.tmem_stride) + ((uint32_t)(32 * q) << 16);
        TilePos tp;
        tp.init(((int)blockIdx.x + buf * (int)gridDim.x) * TC_TILE, p.LhV);
        for (int i = buf, u = 0; i < my_tiles; i += nbuf, ++u) {
            // output address = warp-uniform 64-bit row base (sequence tp.n0, motif kbase + 8 ch) + per-thread 32-bit byte offset
            // (its sequence relative to tp.n0, its position, the motif inside the chunk): one STG per value, no 64-bit math
            int n, j;
            bool ok;
            tp.row(32 * q + lane, p.LhV, p.VP, n, j, ok);
            const bool valid = ok && n < d.n_rows && j < p.store_len;
            const uint32_t rs4 = (uint32_t)p.row_stride * 4u;
            const uint32_t off0 = (uint32_t)(n - tp.n0) * (uint32_t)d.K * rs4 + (uint32_t)j * 4u;
            uint32_t offc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) offc[c] = off0 + (uint32_t)c * rs4;
            char *ub_tile = reinterpret_cast<char *>(d.out) + ((size_t)tp.n0 * d.K + kbase) * (size_t)p.row_stride * 4;
            tp.advance(stride_pos * nbuf, p.LhV);
            mbar_wait(full + buf, (uint32_t)(u & 1));
            tc_fence_after();
            auto finish = [&](uint32_t (&v)[8], int ch) {
                tmem_ld_fence8(v);
                const float4 b0 = reinterpret_cast<const float4 *>(bias_s)[2 * ch], b1 = reinterpret_cast<const float4 *>(bias_s)[2 * ch + 1];
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float x[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) x[c] = __uint_as_float(v[c]) + bb[c];
                if (PROB && POOL) {
                    // p = exp(a) / (P + sum_group exp(a)) (convRBM.py:252) with x = -log2(e) a; a group is P consecutive lanes of
                    // this warp. a is clamped at 40 so that the paired reciprocal below stays finite ((16 e^40)^2 < FLT_MAX).
                    const float fpool = (float)p.P;
                    const int gs = lane - lane % p.P;
                    float e[8], t[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[c]) : "f"(fminf(-x[c], 57.7f)));
#pragma unroll
                    for (int c = 0; c < 8; ++c) t[c] = fpool + pool_group_sum(e[c], p.P, gs);
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        float r;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t[c] * t[c + 1]));
                        x[c] = e[c] * (r * t[c + 1]);
                        x[c + 1] = e[c + 1] * (r * t[c]);
                    }
                } else if (PROB) {
                    // p = 1 / (1 + 2^x) with x = -log2(e) * activity. The MUFU pipe is the busiest unit of this kernel
                    // (two ops per output), so two outputs share one reciprocal: r = 1 / ((1+e0)(1+e1)), p0 = r (1+e1),
                    // p1 = r (1+e0). x is clamped to 60 so that the product stays finite (p < 1e-18 is unaffected at 1e-5).
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        float e0, e1, r;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(x[c], 60.f)));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(x[c + 1], 60.f)));
                        const float a0 = 1.0f + e0, a1 = 1.0f + e1;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a0 * a1));
                        x[c] = r * a1;
                        x[c + 1] = r * a0;
                    }
                }
                char *ub = ub_tile + (size_t)ch * stride8 * 4;
                const int krem = kvalid - ch * 8;                        // valid motifs in this chunk (warp-uniform)
                if (krem >= 8) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (valid) __stcs(reinterpret_cast<float *>(ub + offc[c]), x[c]);
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (valid && c < krem) __stcs(reinterpret_cast<float *>(ub + offc[c]), x[c]);
                }
            };
            // the TMEM load of the next chunk is in flight while the current one is finished (tcgen05.wait::ld waits for every
            // outstanding load, so the pipeline is one load deep; 16-column loads were measured: 67.7 % instead of 68.6 % of HBM)
            uint32_t va[8], vb[8];
            int ch = sub;
            if (ch < nch8) tmem_ld8(trow + ch * 8, va);
            while (ch < nch8) {
                tmem_ld_wait();
                if (ch + groups_per_buf < nch8) tmem_ld8(trow + (ch + groups_per_buf) * 8, vb);
                finish(va, ch);
                ch += groups_per_buf;
                if (ch >= nch8) break;
                tmem_ld_wait();
                if (ch + groups_per_buf < nch8) tmem_ld8(trow + (ch + groups_per_buf) * 8, va);
                finish(vb, ch);
                ch += groups_per_buf;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)TC_TMEM_COLS);
}

static int pow2_at_least(int x, int lo) { int v = lo; while (v < x) v <<= 1; return v; }

int launch_v2h_dense_tc(const DenseParams &d, long long row_stride, cudaStream_t stream) {
    TcParams p{};
    p.d = d;
    const int Kdim = 16 * ceil_div(4 * d.M, 16);
    p.ksteps = Kdim / 16;
    p.n_atoms = ceil_div(Kdim, 64);
    const int Kpad = 16 * ceil_div(d.K, 16);
    // motifs per CTA: as many as fit in shared memory (B resident for the whole kernel) and in one MMA (N <= 256)
    const size_t smem_cap = 220 * 1024;
    int n_mtiles = ceil_div(Kpad, 256);
    int NT = 16 * ceil_div(ceil_div(Kpad, n_mtiles), 16);
    while (tc_smem_bytes(NT, p.n_atoms, 2) > smem_cap && NT > 16) {
        ++n_mtiles;
        NT = 16 * ceil_div(ceil_div(Kpad, n_mtiles), 16);
    }
    p.NT = NT;
    n_mtiles = ceil_div(d.K, NT);
    p.a_stages = tc_smem_bytes(NT, p.n_atoms, TC_MAX_A_STAGES) <= smem_cap ? TC_MAX_A_STAGES : 2;
    p.tmem_stride = pow2_at_least(NT, 128);
    p.nbuf = TC_TMEM_COLS / p.tmem_stride;                           // 4 accumulator buffers (N <= 128) or 2
    // sector-aligned output rows (32-byte multiples): cut the tiles from rows padded to 32 positions, so that every warp
    // store is 128 contiguous bytes starting on a sector boundary (scratch/store_align.cu: 5.9 TB/s vs 3.5 TB/s otherwise)
    p.row_stride = row_stride;
    const bool aligned = (row_stride % 8 == 0) && (((uintptr_t)d.out & 31) == 0);
    const bool prob = d.out_kind == CRBM_OUT_PROBABILITY;
    const bool pool = prob && d.pooling > 1;
    p.P = pool ? d.pooling : 1;
    p.VP = pool ? (32 / p.P) * p.P : 32;
    // pooled: every warp must hold whole pool groups of one sequence, so the tiles are always cut from the virtual rows
    p.LhV = (aligned || pool) ? 32 * ceil_div(d.Lh, p.VP) : d.Lh;
    const int covered = (p.LhV / 32) * p.VP;                       // real positions the virtual row covers (>= Lh)
    p.store_len = aligned ? (int)(row_stride < covered ? row_stride : covered) : d.Lh;
    const long long total = (long long)d.n_rows * p.LhV;            // (virtual) hidden positions
    CRBM_CHECK_ARG((long long)d.K * row_stride * (TC_TILE / p.LhV + 3) < (1ll << 29), "v2h_dense_tc: K * row stride too large for 32-bit tile offsets");
    CRBM_CHECK_ARG((total + TC_TILE - 1) / TC_TILE < (1ll << 24), "v2h_dense_tc: more than 2^31 hidden positions in one call");
    p.n_tiles = (int)((total + TC_TILE - 1) / TC_TILE);
    p.scale = d.out_kind == CRBM_OUT_PROBABILITY ? -1.4426950408889634f : 1.0f;
    // the CTA allocates all 512 TMEM columns: ask for more than half of the shared memory so that two CTAs never share an SM
    const size_t smem = tc_smem_bytes(NT, p.n_atoms, p.a_stages) > 120 * 1024 ? tc_smem_bytes(NT, p.n_atoms, p.a_stages) : 120 * 1024;
    const int variant = pool ? 2 : prob ? 1 : 0;
    void (*kernel)(const TcParams) = variant == 2 ? v2h_dense_tc_kernel<true, true> : variant == 1 ? v2h_dense_tc_kernel<true, false>
                                                                                                  : v2h_dense_tc_kernel<false, false>;
    static size_t configured[3] = {0, 0, 0};
    if (smem > configured[variant]) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[variant] = smem;
    }
    int gx = num_sms() / n_mtiles;
    if (gx < 1) gx = 1;
    if (gx > p.n_tiles) gx = p.n_tiles;
    kernel<<<dim3(gx, n_mtiles), TC_THREADS, smem, stream>>>(p);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

}  // namespace crbm
