// Dense motif scan (getHitProbs / motifHitProbs, reference secomo/convRBM.py:238-257,507-514) with MOTIF = TMEM LANE and rows of
// the output staged in shared memory and written by the TMA engine.
//
// The position-major kernel of scan_tc.cu gives a thread one position and eight motifs: a pool group is P lanes of a warp (shuffles
// for every value) and a warp's store is 4 * (32 / P) * P bytes of one output row (120 bytes for P = 5: partial 32-byte sectors,
// measured 19.6 % of the HBM peak). Here the machinery of scan_rt.cuh is used: MMA row r = motif 128 chunk + r (strand-selected
// W as fp16 hi + lo, A operand), columns = hidden positions (one-hot tile built from the packed words, B operand), a thread of the
// epilogue owns one motif and 32 consecutive positions: the pooled softmax p = e / (P + sum_group e) (:252) is a run of registers.
// Two store paths, by the block width VP = (32 / P) * P:
//   VP = 32 (P = 1, 2, 4): a block is a whole 128-byte line of an output row and the epilogue warps are independent of each other:
//     the CTA's (sequence, block) pairs are dealt out two per unit without regard to sequence boundaries, a warp stages its 32 motifs x
//     32 positions in its own 4 KB of shared memory (SWIZZLE_128B: conflict-free 16-byte stores) and ONE lane stores the tile with
//     cp.async.bulk.tensor.3d (tensor map over (position, motif, sequence): rows beyond K and positions beyond the row are clipped by
//     the TMA engine). The four warps of an SM sub-partition drift apart, so the exponentials of one overlap the stores of another.
//   VP = 30 (P = 5, 10): 120 bytes are no legal TMA box. A CTA follows ONE sequence at a time: a step is four consecutive blocks (two
//     units of 64 columns, one per epilogue group) whose 4 VP probabilities per motif are staged as one run [motif][4 VP] by the four
//     warps of a lane quarter (named barriers of 128 threads) and stored with one cp.async.bulk of 480 bytes per motif row.
// Whole 32-byte sectors are written when the output rows are padded to multiples of 8 floats, which crbm_b200/convRBM.py does.
// No per-value store instruction, no shuffle. Tile supply: scan_rt.cuh (four producer warps, rows from the pattern table).
// Bound: HBM write, 4 K bytes per hidden position.
#pragma once
#include <cuda.h>
#include "train_tc.cuh"

namespace crbm {

constexpr int RD_D1_BUFS = 8;
constexpr int RD_SROW = 132;                  // floats per staged row: 4 VP (<= 128) + up to 7 pad floats; 132 = 4 mod 32: conflict-free 16-byte stores
constexpr int RD_XSTAGES = 4;
constexpr int RD_THREADS = 672;               // warps 0-3 build the one-hot tiles (two pairs on alternate units), warp 4 owns TMEM and issues
                                              // the MMAs, warps 5-20 are the epilogue (warp & 3 = TMEM lane quarter)

struct RowDenseParams {
    const uint32_t *packed;
    int row_words, n_rows, L, Lh;
    const float *W, *b;
    int K, M;
    int strand_mode, out_kind;
    int nblk32, nsteps, ksteps, n_atoms, Kdim, xstages;
    float *out;
    long long row_stride;          // floats between consecutive (n, k) rows (multiple of 4)
    int store_len;                 // floats written per row: min(row_stride, Lh rounded up to 8)
    int debug;                     // CRBM_RD_DEBUG bits (timing experiments only): 1 = no stores, 2 = no exponentials
    int tensor_store;              // VP = 32: per-warp tiles stored through the tensor map (else one bulk copy per motif row and step)
};

__host__ __device__ inline size_t rd_smem_bytes(int n_atoms, int xstages) {
    return 1024 + 2 * (size_t)n_atoms * TT_WATOM + (size_t)xstages * n_atoms * TT_XATOM + (size_t)TT_ROWS * RD_SROW * 4 + 512 + 512 + 512;
}

__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tensor_store_3d(const CUtensorMap *map, const void *src_smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src_smem)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void quarter_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

template <int P>
__global__ void __launch_bounds__(RD_THREADS, 1) v2h_rowdense_tc_kernel(const RowDenseParams p, const __grid_constant__ CUtensorMap tmap) {
    constexpr int VP = (32 / P) * P, NG = 32 / P;
    const bool tensor = VP == 32 && p.tensor_store;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const size_t w_bytes = (size_t)p.n_atoms * TT_WATOM, x_bytes = (size_t)p.n_atoms * TT_XATOM;
    unsigned char *Whi = smem, *Wlo = Whi + w_bytes;
    unsigned char *X = Wlo + w_bytes;
    float *stage_s = reinterpret_cast<float *>(X + (size_t)p.xstages * x_bytes);
    float *bias_s = stage_s + TT_ROWS * RD_SROW;
    uint64_t *x_full = reinterpret_cast<uint64_t *>(bias_s + TT_ROWS);
    uint64_t *x_free = x_full + TT_MAX_XSTAGES;
    uint64_t *d1_full = x_free + TT_MAX_XSTAGES;
    uint64_t *d1_free = d1_full + RD_D1_BUFS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d1_free + RD_D1_BUFS);
    unsigned char *lut_s = reinterpret_cast<unsigned char *>(tmem_slot) + 16 + ((16 - ((smem_u32(tmem_slot) + 16) & 15)) & 15);
    const int XS = p.xstages;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int chunk = blockIdx.y, G = gridDim.x;
    const int Kdim = p.Kdim;
    const bool prob = p.out_kind == CRBM_OUT_PROBABILITY;
    // accumulator scale: sigmoid as 1 / (1 + 2^x) needs x = -log2(e) a; the pooled softmax needs 2^x = e^a
    const float wscale = !prob ? 1.0f : P == 1 ? -1.4426950408889634f : 1.4426950408889634f;

    if (warp == 4) tmem_alloc(tmem_slot, 512u);
    if (tid == 64) {
        for (int i = 0; i < TT_MAX_XSTAGES; ++i) { mbar_init(x_full + i, 64); mbar_init(x_free + i, 1); }
        for (int i = 0; i < RD_D1_BUFS; ++i) { mbar_init(d1_full + i, 1); mbar_init(d1_free + i, 8); }
    }
    tt_lut_init(lut_s, tid);
    // ---- A operand: the chunk's rows of W (strand-selected), scaled, fp16 hi / lo, K-major SWIZZLE_128B ---------------------------
    for (int idx = tid; idx < TT_ROWS * (Kdim >> 3); idx += RD_THREADS) {
        const int r = idx & (TT_ROWS - 1), ch = idx >> 7;
        const int k = chunk * TT_ROWS + r;
        const bool want_f = p.strand_mode != CRBM_STRAND_RC, want_r = p.strand_mode != CRBM_STRAND_FWD;
        float w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int m = 2 * ch + (e >> 2), a = e & 3;
            float v = 0.f;
            if (k < p.K && m < p.M) {
                if (want_f) v += __ldg(p.W + ((size_t)k * 4 + a) * p.M + m);
                if (want_r) v += __ldg(p.W + ((size_t)k * 4 + (3 - a)) * p.M + (p.M - 1 - m));
            }
            w[e] = v * wscale;
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
    for (int r = tid; r < TT_ROWS; r += RD_THREADS) {
        const int k = chunk * TT_ROWS + r;
        const float bs = p.strand_mode == CRBM_STRAND_SUM ? 2.f : 1.f;      // both activities carry the bias (:513)
        bias_s[r] = k < p.K ? bs * wscale * __ldg(p.b + k) : 0.f;
    }
    for (int i = tid; i < TT_ROWS * RD_SROW; i += RD_THREADS) stage_s[i] = 0.f;     // the pad columns behind the last block of a run stay zero
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // work of this CTA: sequences x, x + G, ...; a sequence is nsteps steps of two units (one per epilogue group)
    const int my_seqs = ((int)blockIdx.x < p.n_rows) ? (p.n_rows - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int my_blocks = my_seqs * p.nblk32;                                       // tensor-store mode: two blocks per unit
    const int my_units = tensor ? (my_blocks + 1) >> 1 : 2 * my_seqs * p.nsteps;

    if (warp < 4) {
        // ================================================= producers: one-hot tiles ========================================
        // Two pairs of warps on alternate units (pair = warp >> 1): the proxy fence behind a tile waits for the thread's outstanding
        // global loads (the packed words of its next tile), so one pair delivers a tile per load latency; two pairs hide it.
        const int nch = 2 * p.ksteps, pair = warp >> 1, ptid = tid & 63;
        // unit t, tile row rr -> (sequence, position): bulk mode: group t & 1 of step t >> 1, block 4 step + 2 (t & 1) + (rr >> 5) of the
        // step's sequence; tensor mode: block g = 2 t + (rr >> 5) of the CTA's list, sequence g / nblk
        auto load_row = [&](int t, uint32_t (&w)[4], int &sh, bool &valid) {
            const int i = ptid & 31;
            int n, j;
            if (tensor) {
                const int g = 2 * t + (ptid >> 5), si = g / p.nblk32;
                n = g < my_blocks ? (int)blockIdx.x + si * G : p.n_rows;
                j = (g - si * p.nblk32) * 32 + i;
            } else {
                const int step = t >> 1, si = step / p.nsteps, st = step - si * p.nsteps;
                n = (int)blockIdx.x + si * G;
                j = (4 * st + 2 * (t & 1) + (ptid >> 5)) * VP + i;
            }
            valid = i < VP && j < p.Lh && n < p.n_rows;
            sh = 0;
            w[0] = w[1] = w[2] = w[3] = 0u;
            if (valid) {
                const int bit = 2 * (j + 16 * FRONT_PAD_WORDS);
                const uint32_t *src = p.packed + (size_t)n * p.row_words + (bit >> 5);
                sh = bit & 31;
#pragma unroll
                for (int e = 0; e < 4; ++e) w[e] = __ldg(src + e);
            }
        };
        uint32_t cw[4], nw[4];
        int csh = 0, nsh = 0;
        bool cv = false, nv = false;
        if (pair < my_units) load_row(pair, cw, csh, cv);
        for (int t = pair; t < my_units; t += 2) {
            const int stage = t % XS, use = t / XS;
            if (t + 2 < my_units) load_row(t + 2, nw, nsh, nv);
            if (use > 0) mbar_wait(x_free + stage, (uint32_t)((use - 1) & 1));
            tt_build_row_lut<VP == 32>(X + (size_t)stage * x_bytes, lut_s, ptid, nch, p.M, cw, csh, cv);
            fence_async_smem();
            mbar_arrive(x_full + stage);
#pragma unroll
            for (int e = 0; e < 4; ++e) cw[e] = nw[e];
            csh = nsh; cv = nv;
        }
    } else if (warp == 4) {
        // ================================================= MMA issue (one thread) ==========================================
        if (elect_one()) {
            const uint32_t whi_lo = umma_desc_lo(smem_u32(Whi)), wlo_lo = umma_desc_lo(smem_u32(Wlo)), xa = smem_u32(X);
            const uint32_t idesc1 = (1u << 4) | ((uint32_t)(TT_UNIT >> 3) << 17) | ((uint32_t)(TT_ROWS >> 4) << 24);
            for (int t = 0; t < my_units; ++t) {
                const int stage = t % XS, buf = t & (RD_D1_BUFS - 1);
                mbar_wait(x_full + stage, (uint32_t)((t / XS) & 1));
                if (t >= RD_D1_BUFS) mbar_wait(d1_free + buf, (uint32_t)(((t >> 3) - 1) & 1));
                tc_fence_after();
                const uint32_t d1 = tmem_base + 64 * buf;
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
                umma_commit(x_free + stage);
            }
        }
    } else {
        // ================================================= epilogue ========================================================
        // (any four consecutive warps cover the four TMEM lane quarters: warp & 3)
        const int ew = warp - 5, q = warp & 3, wg = ew >> 2, pr = wg >> 1, h = wg & 1;
        const int r = 32 * q + lane;
        const float bias = bias_s[r];
        float *my_stage = stage_s + (size_t)r * RD_SROW + (2 * pr + h) * VP;
        int qi = 0, st = 0;
        for (int t = pr; t < my_units; t += 2) {
            const int n = (int)blockIdx.x + qi * G;
            const int buf = t & (RD_D1_BUFS - 1);
            mbar_wait_backoff(d1_full + buf, (uint32_t)((t >> 3) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + 64 * buf + 32 * h;
            uint32_t va[16], vb[16];
            tmem_ld16(taddr, va);
            tmem_ld16(taddr + 16, vb);
            tmem_ld_wait();
            tmem_ld_fence16(va);
            tmem_ld_fence16(vb);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d1_free + buf);          // eight warps per buffer: the four lane quarters of both halves
            float pv[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) { pv[i] = __uint_as_float(va[i]) + bias; pv[16 + i] = __uint_as_float(vb[i]) + bias; }
            if (p.debug & 2) {
            } else if (prob && P == 1) {
                // p = 1 / (1 + 2^x), x = -log2(e) a; two values share one reciprocal (x clamped at 60: the product stays finite)
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float e0, e1, rr;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(pv[i], 60.f)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(pv[i + 1], 60.f)));
                    const float a0 = 1.0f + e0, a1 = 1.0f + e1;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr) : "f"(a0 * a1));
                    pv[i] = rr * a1;
                    pv[i + 1] = rr * a0;
                }
            } else if (prob) {
                // pooled probabilities p = e^a / (P + sum_group e^a) (convRBM.py:252), evaluated as 2^(x - m) / (P 2^-m + sum 2^(x - m))
                // with m = max(0, max_group x): finite and accurate for any activity (the reference overflows beyond a = 88)
                constexpr float fpool = (float)P;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float m = 0.f;
#pragma unroll
                    for (int i = 0; i < P; ++i) m = fmaxf(m, pv[g * P + i]);
                    float tot, e0;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-m));
                    tot = fpool * e0;
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        float e;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(pv[g * P + i] - m));
                        pv[g * P + i] = e;
                        tot += e;
                    }
                    float inv;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(tot));
#pragma unroll
                    for (int i = 0; i < P; ++i) pv[g * P + i] *= inv;
                }
            }
            if (tensor) {
                // this warp's own tile [32 motifs][32 positions], rows of 128 bytes, 16-byte chunks XOR-swizzled with the row (SWIZZLE_128B)
                const int g = 2 * t + h, si = g / p.nblk32, blk = g - si * p.nblk32;
                unsigned char *tile = reinterpret_cast<unsigned char *>(stage_s) + (size_t)ew * 4096;
                if (lane == 0) bulk_wait_read();                // the previous store of this warp has read the tile
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4 *>(tile + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_float4(pv[4 * c], pv[4 * c + 1], pv[4 * c + 2], pv[4 * c + 3]);
                fence_async_smem();
                __syncwarp();
                if (lane == 0 && g < my_blocks && !(p.debug & 1)) {
                    tensor_store_3d(&tmap, tile, blk * 32, chunk * TT_ROWS + 32 * q, (int)blockIdx.x + si * G);
                    bulk_commit();
                }
                continue;
            }
            // the staging rows of this lane quarter are free again when the bulk copies of the previous step have read them; only the
            // four warps of a lane quarter (one per block of the step) synchronise, so the quarters drift apart and the exponentials
            // of one overlap the shared-memory stores and copy issue of another
            if (lane < 8) bulk_wait_read();
            quarter_bar(4 + q);
            if (VP == 32) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4 *>(my_stage + i) = make_float4(pv[i], pv[i + 1], pv[i + 2], pv[i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < VP; i += 2) *reinterpret_cast<float2 *>(my_stage + i) = make_float2(pv[i], pv[i + 1]);
            }
            fence_async_smem();
            quarter_bar(8 + q);
            if (lane < 8) {
                // warp wg of the quarter stores rows 32 q + 8 wg + lane: one bulk copy per motif row and step
                const int rr = 32 * q + 8 * wg + lane, kk = chunk * TT_ROWS + rr;
                if (kk < p.K && !(p.debug & 1)) {
                    const int start = 4 * VP * st;
                    const int run = st == p.nsteps - 1 ? p.store_len - start : 4 * VP;
                    bulk_store(p.out + ((size_t)n * p.K + kk) * (size_t)p.row_stride + start, stage_s + (size_t)rr * RD_SROW, (uint32_t)run * 4u);
                }
                bulk_commit();
            }
            if (++st == p.nsteps) { st = 0; ++qi; }
        }
        if (lane < 8) bulk_wait_all();                         // (tensor-store mode: lane 0 holds the groups)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512u);
}

template <int P> int launch_v2h_rowdense_tc(const RowDenseParams &p, const CUtensorMap &tmap, dim3 grid, cudaStream_t stream);

}  // namespace crbm
