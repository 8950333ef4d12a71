// Scan with per-sequence reductions, and free energy / evaluation, on the 5th-generation tensor cores with MOTIF = TMEM LANE.
//
//   RT_REDUCE  (reference secomo/utils.py:155,245-247 max over positions of motifHitProbs, :306 mean):
//              per sequence and motif the maximum hit probability (or activity), the position of the first maximum and the mean
//              over positions. MMA row r = motif 128 chunk + r of ONE strand selection (forward / reverse complement / sum of
//              both: getHitProbs' two branches, convRBM.py:507-514).
//   RT_FE      (convRBM.py:657-697 free energy, :466-472 evaluateData): per sequence and motif
//              sum_groups log(1 + sum_group exp(x)) over both strands, plus the number of units a forward-strand hidden sample
//              switches on. MMA row r = (strand r / 64, motif 64 chunk + r % 64), as in train_tc.cuh.
//
// Same machinery as train_tc.cuh (one-hot tiles built in shared memory from packed 2-bit words, W as fp16 hi/lo A operand,
// D1[row, position] in TMEM, a thread owns one row and 32 consecutive positions), without the second GEMM: all of tensor memory is
// eight accumulator buffers. The per-sequence reductions stay in registers because a warpgroup follows ONE sequence: a CTA works on
// four sequences at a time (two per epilogue group; rows 0-31 / 32-63 of every one-hot tile belong to the two sequences of the group),
// walking their 32-position blocks in lock step, so thread (row, stream) sees all blocks of its sequence in order and writes the
// result once. Nothing but (N,K) results reaches HBM: the kernel is bound by the epilogue's instruction issue, not by memory.
#pragma once
#include "train_tc.cuh"

namespace crbm {

enum : int { RT_REDUCE = 0, RT_FE = 1, RT_REDUCE_NOMEAN = 2 };      // NOMEAN: maximum and position only (out_mean == null)
constexpr int RT_D1_BUFS = 8;
constexpr int RT_THREADS = 672;               // warps 0-3 build the one-hot tiles (two pairs on alternate units), warp 4 owns TMEM and issues
                                              // the MMAs, warps 5-20 are the epilogue (warp & 3 = TMEM lane quarter)

struct RowsTcParams {
    const uint32_t *packed;
    int row_words, n_rows, L, Lh;
    const float *W, *b;
    int K, M;
    int strand_mode, out_kind;     // RT_REDUCE
    int ds;                        // RT_FE
    int nblk32, ksteps, n_atoms, Kdim, xstages;
    int n_quads;                   // ceil(n_rows / 4)
    float *out_max, *out_mean;     // RT_REDUCE: (N,K) each, may be null
    int *out_argmax;
    float *fe_motif, *fe_seq, *count;   // RT_FE: accumulated (zeroed by the caller); each may be null
    const float *uniforms;         // RT_FE count: injected uniforms [(n*K+k)*(Lh/P) + group] or null
    uint64_t seed;
    uint32_t ctr_hi;
    long long *dbg;                // CRBM_RT_DEBUG=1: CTA (0,0) records clock64() per role, unit and stage: [role][RT_DBG_UNITS][4]
};
constexpr int RT_DBG_UNITS = 64, RT_DBG_ROLES = 5;
#ifdef CRBM_RT_TIMELINE          // (the marks cost the reduction kernels ~5 %: compiled in only with -DCRBM_RT_TIMELINE)
#define RT_MARK(role, t, i) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (t) < RT_DBG_UNITS) p.dbg[((role) * RT_DBG_UNITS + (t)) * 4 + (i)] = clock64(); } while (0)
#else
#define RT_MARK(role, t, i) do { } while (0)
#endif

__host__ __device__ inline size_t rt_smem_bytes(int n_atoms, int xstages) {
    return 1024 + 2 * (size_t)n_atoms * TT_WATOM + (size_t)xstages * n_atoms * TT_XATOM + 512 + 512 + 512;
}

// ln(1 + t) for t = 2^y >= 0 given y (so that huge arguments need no overflow handling): series below 1/16, lg2 above, y ln 2 beyond 2^24
__device__ __forceinline__ float rt_softplus(float t, float y) {
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + t));
    const float mid = l * 0.6931471805599453f;
    const float small = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 0.2f, -0.25f), 0.3333333333f), -0.5f), 1.0f);
    return y > 24.f ? y * 0.6931471805599453f : (t < 0.0625f ? small : mid);
}

template <int MODE_T, int P>
__global__ void __launch_bounds__(RT_THREADS, 1) v2h_rows_tc_kernel(const RowsTcParams p) {
    constexpr int MODE = MODE_T == RT_FE ? RT_FE : RT_REDUCE;
    constexpr bool want_mean = MODE_T != RT_REDUCE_NOMEAN;
    constexpr int VP = (32 / P) * P, NG = 32 / P;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const size_t w_bytes = (size_t)p.n_atoms * TT_WATOM, x_bytes = (size_t)p.n_atoms * TT_XATOM;
    unsigned char *Whi = smem, *Wlo = Whi + w_bytes;
    unsigned char *X = Wlo + w_bytes;
    float *bias_s = reinterpret_cast<float *>(X + (size_t)p.xstages * x_bytes);
    uint64_t *x_full = reinterpret_cast<uint64_t *>(bias_s + TT_ROWS);
    uint64_t *x_free = x_full + TT_MAX_XSTAGES;
    uint64_t *d1_full = x_free + TT_MAX_XSTAGES;
    uint64_t *d1_free = d1_full + RT_D1_BUFS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d1_free + RT_D1_BUFS);
    unsigned char *lut_s = reinterpret_cast<unsigned char *>(tmem_slot) + 16 + ((16 - ((smem_u32(tmem_slot) + 16) & 15)) & 15);
    const int XS = p.xstages;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int chunk = blockIdx.y, G = gridDim.x;
    const int Kdim = p.Kdim;
    const bool prob = MODE == RT_FE || p.out_kind == CRBM_OUT_PROBABILITY;
    // accumulator scale: sigmoid as 1 / (1 + 2^x) needs x = -log2(e) a; the pooled softmax and the free energy need 2^x = e^a
    const float wscale = !prob ? 1.0f : (MODE == RT_REDUCE && P == 1) ? -1.4426950408889634f : 1.4426950408889634f;

    if (warp == 4) tmem_alloc(tmem_slot, 512u);
    tt_lut_init(lut_s, tid);
    if (tid == 64) {
        for (int i = 0; i < TT_MAX_XSTAGES; ++i) { mbar_init(x_full + i, 64); mbar_init(x_free + i, 1); }
        for (int i = 0; i < RT_D1_BUFS; ++i) { mbar_init(d1_full + i, 1); mbar_init(d1_free + i, 8); }
    }
    // ---- A operand: the chunk's rows of W (strand-selected), scaled, fp16 hi / lo, K-major SWIZZLE_128B ---------------------------
    for (int idx = tid; idx < TT_ROWS * (Kdim >> 3); idx += RT_THREADS) {
        const int r = idx & (TT_ROWS - 1), ch = idx >> 7;
        int k, want_f, want_r;
        if (MODE == RT_REDUCE) {
            k = chunk * TT_ROWS + r;
            want_f = p.strand_mode != CRBM_STRAND_RC; want_r = p.strand_mode != CRBM_STRAND_FWD;
        } else {
            const int s = r >> 6;
            k = chunk * TT_KC + (r & 63);
            want_f = s == 0; want_r = s == 1 && p.ds;
        }
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
    for (int r = tid; r < TT_ROWS; r += RT_THREADS) {
        const int k = MODE == RT_REDUCE ? chunk * TT_ROWS + r : chunk * TT_KC + (r & 63);
        const float bs = (MODE == RT_REDUCE && p.strand_mode == CRBM_STRAND_SUM) ? 2.f : 1.f;      // both activities carry the bias (:513)
        bias_s[r] = k < p.K ? bs * wscale * __ldg(p.b + k) : 0.f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // work of this CTA: sequence quads x, x + G, ...; every quad is nblk32 steps of two units (one per epilogue group)
    const int my_quads = ((int)blockIdx.x < p.n_quads) ? (p.n_quads - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int my_units = 2 * my_quads * p.nblk32;

    if (warp < 4) {
        // ================================================= producers: one-hot tiles ========================================
        // Two pairs of warps on alternate units (pair = warp >> 1): the proxy fence behind a tile waits for the thread's outstanding
        // global loads (the packed words of its next tile), so one pair delivers a tile per load latency; two pairs hide it. Rows come
        // from the pattern table (tt_build_row_lut).
        const int nch = 2 * p.ksteps, pair = warp >> 1, ptid = tid & 63;
        // unit t: group t & 1, step t >> 1 = (quad iteration, block); tile row rr: stream (group, rr >> 5), position rr & 31 of the block
        auto load_row = [&](int t, uint32_t (&w)[4], int &sh, bool &valid) {
            const int step = t >> 1, qi = step / p.nblk32, wb = step - qi * p.nblk32;
            const int n = 4 * ((int)blockIdx.x + qi * G) + 2 * (t & 1) + (ptid >> 5), i = ptid & 31;
            const int j = wb * VP + i;
            valid = n < p.n_rows && i < VP && j < p.Lh;
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
            if (ptid == 0) RT_MARK(pair, t, 0);
            if (use > 0) mbar_wait(x_free + stage, (uint32_t)((use - 1) & 1));
            if (ptid == 0) RT_MARK(pair, t, 1);
            tt_build_row_lut<false>(X + (size_t)stage * x_bytes, lut_s, ptid, nch, p.M, cw, csh, cv);
            if (ptid == 0) RT_MARK(pair, t, 2);
            fence_async_smem();
            mbar_arrive(x_full + stage);
            if (ptid == 0) RT_MARK(pair, t, 3);
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
                const int stage = t % XS, buf = t & (RT_D1_BUFS - 1);
                RT_MARK(2, t, 0);
                mbar_wait(x_full + stage, (uint32_t)((t / XS) & 1));
                RT_MARK(2, t, 1);
                if (t >= RT_D1_BUFS) mbar_wait(d1_free + buf, (uint32_t)(((t >> 3) - 1) & 1));
                RT_MARK(2, t, 2);
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
                RT_MARK(2, t, 3);
            }
        }
    } else {
        // ================================================= epilogue ========================================================
        // (any four consecutive warps cover the four TMEM lane quarters: warp & 3)
        const int ew = warp - 5, q = warp & 3, wg = ew >> 2, pr = wg >> 1, h = wg & 1;
        const int r = 32 * q + lane;
        const int s = MODE == RT_FE ? (q >> 1) : 0;
        const int k = MODE == RT_REDUCE ? chunk * TT_ROWS + r : chunk * TT_KC + (r & 63);
        const bool row_ok = k < p.K && (MODE == RT_REDUCE || s == 0 || p.ds);
        const float bias = bias_s[r];
        const int groups_per_seq = p.Lh / P;
        const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        const float inv_lh = 1.0f / (float)p.Lh;
        // running reductions of the current sequence of this stream
        float best = -INFINITY, sum = 0.f, cnt = 0.f;
        int bpos = 0;
        int qi = 0, wb = 0;
        for (int t = pr; t < my_units; t += 2) {
            const int n = 4 * ((int)blockIdx.x + qi * G) + 2 * pr + h;
            int nv = n < p.n_rows ? min(VP, p.Lh - wb * VP) : 0;
            nv = nv < 0 ? 0 : nv;
            const int buf = t & (RT_D1_BUFS - 1);
            if (q == 0 && h == 0 && lane == 0) RT_MARK(3 + pr, t, 0);
            mbar_wait_backoff(d1_full + buf, (uint32_t)((t >> 3) & 1));
            if (q == 0 && h == 0 && lane == 0) RT_MARK(3 + pr, t, 1);
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
            if (q == 0 && h == 0 && lane == 0) RT_MARK(3 + pr, t, 2);
            float pv[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) { pv[i] = __uint_as_float(va[i]) + bias; pv[16 + i] = __uint_as_float(vb[i]) + bias; }
            if (wb == 0) { best = -INFINITY; sum = 0.f; cnt = 0.f; bpos = 0; }
            if (MODE == RT_REDUCE) {
                if (!prob || P == 1) {
                    // key = the activity (up to the positive factor log2 e): its maximum is the maximum of the sigmoid too
                    // (for probabilities the accumulator is -log2(e) a: the minimum; for activities the maximum)
                    const bool fullblk = nv == 32;                 // warp-uniform; the masked path runs once per sequence
                    float bm;
                    if (prob) {
                        bm = INFINITY;
                        if (fullblk) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) bm = fminf(bm, pv[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) bm = fminf(bm, i < nv ? pv[i] : INFINITY);
                        }
                    } else {
                        bm = -INFINITY;
                        if (fullblk) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) bm = fmaxf(bm, pv[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) bm = fmaxf(bm, i < nv ? pv[i] : -INFINITY);
                        }
                    }
                    const float key = prob ? -bm : bm;
                    if (key > best) {                           // strict: an earlier block keeps a tie (numpy.argmax)
                        int at = 0;
#pragma unroll
                        for (int i = 31; i >= 0; --i) at = (i < nv && pv[i] == bm) ? i : at;
                        best = key;
                        bpos = wb * VP + at;
                    }
                    // the probabilities themselves are only needed for the mean: a caller that asks for the maximum and its position
                    // alone (utils.py:155,245-247) gets them from the activities, one sigmoid per sequence and motif at the end
                    if (prob && want_mean) {
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
                    }
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    if (!want_mean) {
                    } else if (fullblk) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) { s0 += pv[i]; s1 += pv[i + 1]; s2 += pv[i + 2]; s3 += pv[i + 3]; }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) { s0 += i < nv ? pv[i] : 0.f; s1 += i + 1 < nv ? pv[i + 1] : 0.f; }
                    }
                    sum += (s0 + s1) + (s2 + s3);
                } else {
                    // pooled probabilities: p = e / (P + sum_group e) (convRBM.py:252); the maximum is taken over p itself
                    constexpr float fpool = (float)P;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float e;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(pv[i], 57.7f)));
                        pv[i] = i < VP ? e : 0.f;
                    }
                    if (nv != VP) {                             // warp-uniform; whole groups drop out (nv is a multiple of P)
#pragma unroll
                        for (int i = 0; i < VP; ++i) pv[i] = i < nv ? pv[i] : 0.f;
                    }
                    float bm = -INFINITY, s0 = 0.f;
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        float tot = 0.f;
#pragma unroll
                        for (int i = 0; i < P; ++i) tot += pv[g * P + i];
                        float inv;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fpool + tot));
#pragma unroll
                        for (int i = 0; i < P; ++i) {
                            pv[g * P + i] *= inv;
                            s0 += pv[g * P + i];
                            bm = fmaxf(bm, pv[g * P + i]);          // (masked positions hold 0 <= every probability)
                        }
                    }
                    sum += s0;
                    if (bm > best) {
                        int at = 0;
#pragma unroll
                        for (int i = VP - 1; i >= 0; --i) at = (i < nv && pv[i] == bm) ? i : at;
                        best = bm;
                        bpos = wb * VP + at;
                    }
                }
            } else {
                // free energy: sum over pool groups of log(1 + sum_group exp(a)); count: one uniform per group, first unit whose
                // running probability e / (P + sum e) exceeds it (forward strand only, convRBM.py:469-472)
                constexpr float fpool = (float)P;
                float fe = 0.f;
                const bool counting = p.count != nullptr && s == 0;
                const bool fullblk = nv == VP;
                const uint32_t rowid = (uint32_t)(n * p.K + (row_ok ? k : 0));
                float ug4[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float e[P], tot = 0.f, ymax = -INFINITY;
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const float y = fminf(pv[g * P + i], 100.f);
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[i]) : "f"(y));
                        tot += e[i];
                        ymax = fmaxf(ymax, y);
                    }
                    // (for P > 1 the exponent of the group sum is bounded by ymax + log2 P: the y * ln 2 branch uses lg2 of the sum)
                    float lt;
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lt) : "f"(tot));
                    const float sp = rt_softplus(tot, P == 1 ? ymax : lt);
                    if (fullblk || g * P < nv) fe += sp;
                    if (counting && P == 1 && !p.uniforms) {
                        // one Bernoulli(e / (1 + e)) per position from 16-bit uniforms, eight per Philox draw; a tie of the 16 bits
                        // (2^-16 per position) is resolved towards "off": the count is a monitoring statistic (NumH, convRBM.py:469)
                        if ((g & 7) == 0) {
                            const uint4 rnd = tt_philox(make_uint4(rowid, (uint32_t)(wb * 8 + (g >> 3)), p.ctr_hi, RNG_EVAL), key);
                            ug4[0] = __uint_as_float(rnd.x); ug4[1] = __uint_as_float(rnd.y); ug4[2] = __uint_as_float(rnd.z); ug4[3] = __uint_as_float(rnd.w);
                        }
                        const uint32_t u16 = (__float_as_uint(ug4[(g & 7) >> 1]) >> (16 * (g & 1))) & 0xffffu;
                        float inv;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fpool + tot));
                        const bool hit = (int)(e[0] * inv * 65536.f) > (int)u16 && (fullblk || g < nv);
                        cnt += hit ? 1.f : 0.f;
                    } else if (counting) {
                        float u;
                        if (p.uniforms) u = (row_ok && g * P < nv) ? p.uniforms[(size_t)rowid * groups_per_seq + wb * NG + g] : 1.f;
                        else {
                            if ((g & 3) == 0) {
                                const uint4 rnd = tt_philox(make_uint4(rowid, (uint32_t)(wb * 8 + (g >> 2)), p.ctr_hi, RNG_EVAL), key);
                                ug4[0] = u01(rnd.x); ug4[1] = u01(rnd.y); ug4[2] = u01(rnd.z); ug4[3] = u01(rnd.w);
                            }
                            u = ug4[g & 3];
                        }
                        float inv;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fpool + tot));
                        float run = 0.f;
                        bool done = false;
#pragma unroll
                        for (int i = 0; i < P; ++i) {
                            run += e[i] * inv;
                            const bool hit = !done && run > u && (fullblk || g * P + i < nv);
                            done = done || hit;
                            cnt += hit ? 1.f : 0.f;
                        }
                    }
                }
                sum += fe;
            }
            // ---- end of the sequence: write / accumulate its results ------------------------------------------------------------
            if (wb == p.nblk32 - 1 && n < p.n_rows) {
                if (MODE == RT_REDUCE) {
                    if (row_ok) {
                        const size_t o = (size_t)n * p.K + k;
                        float mx = best;
                        if (prob && P == 1) {                   // best = -x~ of the maximum: its sigmoid
                            float e0;
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(-best, 60.f)));
                            mx = __fdividef(1.0f, 1.0f + e0);
                        }
                        if (p.out_max) p.out_max[o] = mx;
                        if (p.out_argmax) p.out_argmax[o] = bpos;
                        if (p.out_mean) p.out_mean[o] = sum * inv_lh;
                    }
                } else {
                    const float v = row_ok ? sum : 0.f;
                    if (p.fe_motif && row_ok) atomicAdd(p.fe_motif + (size_t)n * p.K + k, v);     // two addends (the strands): order-free
                    if (p.fe_seq) {
                        const float tot = warp_sum(v);
                        if (lane == 0 && tot != 0.f) atomicAdd(p.fe_seq + n, tot);
                    }
                    if (p.count && s == 0) {
                        const float tot = warp_sum(row_ok ? cnt : 0.f);
                        if (lane == 0 && tot != 0.f) atomicAdd(p.count + n, tot);
                    }
                }
            }
            if (q == 0 && h == 0 && lane == 0) RT_MARK(3 + pr, t, 3);
            if (++wb == p.nblk32) { wb = 0; ++qi; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512u);
}

template <int P> int launch_v2h_rows_tc(int mode, const RowsTcParams &p, dim3 grid, cudaStream_t stream);

}  // namespace crbm
