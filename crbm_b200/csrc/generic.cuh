// Shape-generic kernels: any motif_length, any pooling factor.
//
// The reference has no shape limits (secomo/convRBM.py:68-108 only asks for positive sizes); the tuned kernels of this library
// cover motif_length <= 32 (tcgen05: 4 M <= 128 contraction columns) / <= 33 (mma.sync, CUDA cores) and pooling <= 16. Everything
// beyond runs here: plain CUDA-core kernels with the motif's weights in shared memory and runtime loops over the motif and the
// pool group — the same quantities (A1-A3 convRBM.py:238-267, A7 :507-514, A8-A11 :327-371,440-451, A13 :657-697), one CTA per
// motif, no register-resident statistics. The statistics are taken in a second pass from the probabilities of the first (a warp per
// (motif, base, offset) sums its entries in a fixed order: deterministic, no floating-point atomics), so the path is a few times
// slower than the tuned ones per unit of work; it exists so that no model the reference accepts is refused.
#pragma once
#include "v2h.cuh"

namespace crbm {

enum : int { GEN_DENSE = 0, GEN_TRAIN = 1, GEN_REDUCE = 2, GEN_FE = 3 };

struct GenParams {
    const uint32_t *packed;
    int row_words, n_rows, L, Lh, row0;        // rows row0 .. row0 + n_rows - 1 of the (cursor-selected) minibatch
    const int *batch_index;
    int batch_rows;                            // rows per minibatch (cursor stride)
    const float *W, *b;
    int K, M, P, ds, mode;
    // GEN_DENSE / GEN_REDUCE
    int strand_mode, out_kind;
    float *out;
    long long out_row_stride;
    float *out_max, *out_mean;
    int *out_argmax;
    // GEN_TRAIN: probabilities of the chunk, [n_rows][K][Lh]
    float *pf, *pr, *q;
    int sample;
    uint32_t *bits;                            // fragment-ordered planes [2][chain rows][ktiles][FANTASY_WORDS], zeroed by the caller
    int chain_rows, ktiles;
    // GEN_FE
    float *fe_motif, *fe_seq, *count;
    // randomness: one uniform per (row, motif, pool group), as in v2h.cuh
    const float *u_fwd, *u_rc;
    uint64_t seed;
    uint32_t ctr_hi, iter;
    const unsigned long long *step_dev;
};

__device__ __forceinline__ float gen_uniform(const float *inj, uint32_t row, int g, int gps, uint2 key, uint32_t hi, uint32_t tag) {
    if (inj) return inj[(size_t)row * gps + g];
    const uint4 r = philox4x32(make_uint4(row, (uint32_t)(g >> 2), hi, tag), key);
    const int c = g & 3;
    return u01(c == 0 ? r.x : c == 1 ? r.y : c == 2 ? r.z : r.w);
}

// activities of hidden position j on both strands; Wf[a * M + m] = W[k, a, m], Wr[a * M + m] = W[k, 3-a, M-1-m]
__device__ __forceinline__ void gen_activity(const uint32_t *row, int j, int M, const float *Wf, const float *Wr, float bias, bool want_r,
                                             float &xf, float &xr) {
    float f = bias, r = bias;
    for (int m = 0; m < M; ++m) {
        const int a = base_at(row, j + m);
        f += Wf[a * M + m];
        if (want_r) r += Wr[a * M + m];
    }
    xf = f;
    xr = r;
}

__global__ void __launch_bounds__(256) gen_v2h_kernel(const GenParams p) {
    extern __shared__ float gen_smem[];
    const int k = blockIdx.x, M = p.M, P = p.P;
    float *Wf = gen_smem, *Wr = gen_smem + 4 * M;
    for (int e = threadIdx.x; e < 4 * M; e += blockDim.x) {
        const int a = e / M, m = e - a * M;
        Wf[e] = __ldg(p.W + ((size_t)k * 4 + a) * M + m);
        Wr[e] = __ldg(p.W + ((size_t)k * 4 + (3 - a)) * M + (M - 1 - m));
    }
    __syncthreads();
    const float bias = __ldg(p.b + k);
    const uint32_t *packed = p.packed + (p.batch_index ? (size_t)(*p.batch_index) * p.batch_rows * p.row_words : 0) + (size_t)p.row0 * p.row_words;
    const int gps = p.Lh / P;                                  // pool groups per sequence
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const uint32_t hi = p.ctr_hi + (p.step_dev ? (uint32_t)(*p.step_dev) : 0u);
    const int tid = blockIdx.y * blockDim.x + threadIdx.x, nthreads = gridDim.y * blockDim.x;
    const float fpool = (float)P;

    if (p.mode == GEN_DENSE || p.mode == GEN_TRAIN) {
        const bool prob = p.mode == GEN_TRAIN || p.out_kind == CRBM_OUT_PROBABILITY;
        const int Pe = prob ? P : 1, groups = p.Lh / Pe;
        const bool dense = p.mode == GEN_DENSE;
        const bool want_f = !dense || p.strand_mode != CRBM_STRAND_RC, want_r = dense ? p.strand_mode != CRBM_STRAND_FWD : (p.ds != 0);
        const bool summed = dense && p.strand_mode == CRBM_STRAND_SUM;
        for (long long item = tid; item < (long long)p.n_rows * groups; item += nthreads) {
            const int n = (int)(item / groups), g = (int)(item - (long long)n * groups), j0 = g * Pe;
            const uint32_t *row = packed + (size_t)n * p.row_words;
            if (!prob) {
                float xf, xr;
                gen_activity(row, j0, M, Wf, Wr, bias, want_r, xf, xr);
                p.out[((size_t)(p.row0 + n) * p.K + k) * p.out_row_stride + j0] = summed ? xf + xr : want_f ? xf : xr;
                continue;
            }
            // pass 1: the group's normaliser (maximum subtracted: finite whatever the weights); pass 2 recomputes the activities
            float mxf = 0.f, mxr = 0.f;
            for (int i = 0; i < Pe; ++i) {
                float xf, xr;
                gen_activity(row, j0 + i, M, Wf, Wr, bias, want_r, xf, xr);
                if (summed) xf += xr;
                mxf = fmaxf(mxf, xf);
                mxr = fmaxf(mxr, xr);
            }
            float sf = 0.f, sr = 0.f;
            for (int i = 0; i < Pe; ++i) {
                float xf, xr;
                gen_activity(row, j0 + i, M, Wf, Wr, bias, want_r, xf, xr);
                if (summed) xf += xr;
                sf += __expf(xf - mxf);
                sr += __expf(xr - mxr);
            }
            const float invf = 1.f / (fpool * __expf(-mxf) + sf), invr = 1.f / (fpool * __expf(-mxr) + sr);
            const float qf = fpool * __expf(-mxf) * invf;                  // 1 - sum_group p (forward strand)
            const uint32_t rowid = (uint32_t)((p.row0 + n) * p.K + k);
            float uf = 2.f, ur = 2.f, runf = 0.f, runr = 0.f;
            if (p.mode == GEN_TRAIN && p.sample) {
                uf = gen_uniform(p.u_fwd, rowid, g, gps, key, hi, RNG_HIDDEN | (p.iter << 8));
                if (want_r) ur = gen_uniform(p.u_rc, rowid, g, gps, key, hi, RNG_HIDDEN_RC | (p.iter << 8));
            }
            bool donef = false, doner = false;
            for (int i = 0; i < Pe; ++i) {
                float xf, xr;
                gen_activity(row, j0 + i, M, Wf, Wr, bias, want_r, xf, xr);
                if (summed) xf += xr;
                const float vf = __expf(xf - mxf) * invf, vr = __expf(xr - mxr) * invr;
                const int j = j0 + i;
                if (dense) {
                    p.out[((size_t)(p.row0 + n) * p.K + k) * p.out_row_stride + j] = (want_f || summed) ? vf : vr;
                    continue;
                }
                const size_t o = ((size_t)n * p.K + k) * p.Lh + j;
                p.pf[o] = vf;
                if (p.pr) p.pr[o] = vr;
                if (p.q) p.q[o] = vf * qf;
                if (p.sample) {
                    // first unit of the group whose running sum exceeds the group's uniform (convRBM.py:259-267)
                    runf += vf;
                    runr += vr;
                    const int kt = k >> 5, kl = k & 31;
                    if (!donef && runf > uf) {
                        donef = true;
                        atomicOr(p.bits + (((size_t)0 * p.chain_rows + p.row0 + n) * p.ktiles + kt) * FANTASY_WORDS + hbit_word(j, kl), 1u << hbit_bit(j, kl));
                    }
                    if (want_r && !doner && runr > ur) {
                        doner = true;
                        atomicOr(p.bits + (((size_t)1 * p.chain_rows + p.row0 + n) * p.ktiles + kt) * FANTASY_WORDS + hbit_word(j, kl), 1u << hbit_bit(j, kl));
                    }
                }
            }
        }
        return;
    }
    // GEN_REDUCE / GEN_FE: one thread per sequence walks its pool groups in order
    const bool reduce = p.mode == GEN_REDUCE;
    const bool prob = !reduce || p.out_kind == CRBM_OUT_PROBABILITY;
    const int Pe = prob ? P : 1, groups = p.Lh / Pe;
    const bool want_r = reduce ? p.strand_mode != CRBM_STRAND_FWD : (p.ds != 0);
    const bool want_f = !reduce || p.strand_mode != CRBM_STRAND_RC;
    const bool summed = reduce && p.strand_mode == CRBM_STRAND_SUM;
    for (int n = tid; n < p.n_rows; n += nthreads) {
        const uint32_t *row = packed + (size_t)n * p.row_words;
        float best = -3.4e38f, total = 0.f, fe = 0.f, fe_c = 0.f, cnt = 0.f;      // fe_c: Kahan compensation (hundreds of O(1) terms per row)
        int best_j = 0;
        for (int g = 0; g < groups; ++g) {
            const int j0 = g * Pe;
            float mxf = 0.f, mxr = 0.f;
            if (prob) {
                for (int i = 0; i < Pe; ++i) {
                    float xf, xr;
                    gen_activity(row, j0 + i, M, Wf, Wr, bias, want_r, xf, xr);
                    if (summed) xf += xr;
                    mxf = fmaxf(mxf, xf);
                    mxr = fmaxf(mxr, xr);
                }
            }
            float sf = 0.f, sr = 0.f;
            if (prob) {
                for (int i = 0; i < Pe; ++i) {
                    float xf, xr;
                    gen_activity(row, j0 + i, M, Wf, Wr, bias, want_r, xf, xr);
                    if (summed) xf += xr;
                    sf += expf(xf - mxf);
                    sr += expf(xr - mxr);
                }
            }
            if (!reduce) {
                // free energy (convRBM.py:657-697): log(1 + sum_group exp(x)) per strand
                float term = mxf + logf(expf(-mxf) + sf);
                if (want_r) term += mxr + logf(expf(-mxr) + sr);
                const float y = term - fe_c, t2 = fe + y;
                fe_c = (t2 - fe) - y;
                fe = t2;
                if (p.count) {
                    // number of forward-strand units a hidden sample switches on (evaluateData, :466-472): P(group on) = sum_group p
                    const uint32_t rowid = (uint32_t)((p.row0 + n) * p.K + k);
                    const float u = gen_uniform(p.u_fwd, rowid, g, gps, key, hi, RNG_EVAL | (p.iter << 8));
                    if (sf / (fpool * __expf(-mxf) + sf) > u) cnt += 1.f;
                }
                continue;
            }
            const float invf = prob ? 1.f / (fpool * __expf(-mxf) + sf) : 1.f, invr = prob ? 1.f / (fpool * __expf(-mxr) + sr) : 1.f;
            for (int i = 0; i < Pe; ++i) {
                float xf, xr;
                gen_activity(row, j0 + i, M, Wf, Wr, bias, want_r, xf, xr);
                if (summed) xf += xr;
                float v = (want_f || summed) ? xf : xr;
                if (prob) v = (want_f || summed) ? __expf(xf - mxf) * invf : __expf(xr - mxr) * invr;
                total += v;
                if (v > best) { best = v; best_j = j0 + i; }
            }
        }
        const size_t o = (size_t)(p.row0 + n) * p.K + k;
        if (reduce) {
            if (p.out_max) p.out_max[o] = best;
            if (p.out_argmax) p.out_argmax[o] = best_j;
            if (p.out_mean) p.out_mean[o] = total / (float)p.Lh;
        } else {
            if (p.fe_motif) p.fe_motif[o] += fe;
            if (p.fe_seq) atomicAdd(p.fe_seq + p.row0 + n, fe);
            if (p.count && cnt != 0.f) atomicAdd(p.count + p.row0 + n, cnt);
        }
    }
}

// ---- statistics from the probabilities of a chunk ------------------------------------------------------------------------
// warp (k, a, m): vh_p[k,a,m] (+)= sum_{n,j} pf[n,k,j] [base(n, j+m) = a] + sum pr[n,k,j] [base(n, j+M-1-m) = 3-a]   (A8 + A10)
//                 vh_q[k,a,m] (+)= sum_{n,j} q[n,k,j] [base(n, j+m) = a]                                                (A11)
// warps K*4*M .. K*4*M + K - 1: the hidden sums of motif k (A9). Lanes stride over (n, j); partial sums are combined by shuffles in a
// fixed order and written (first chunk) or added (later chunks) by lane 0: deterministic.
struct GenStatsParams {
    const uint32_t *packed;
    int row_words, n_rows, Lh, row0;
    const int *batch_index;
    int batch_rows;
    int K, M, accumulate;
    const float *pf, *pr, *q;
    float *vh_p, *vh_q;            // K*4*M each; vh_q may be null
    float *h_sum, *h_fwd, *h_q;    // K each; h_fwd / h_q may be null
};
__global__ void __launch_bounds__(256) gen_stats_kernel(const GenStatsParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nW = p.K * 4 * p.M;
    if (warp >= nW + p.K) return;
    const uint32_t *packed = p.packed + (p.batch_index ? (size_t)(*p.batch_index) * p.batch_rows * p.row_words : 0) + (size_t)p.row0 * p.row_words;
    const int total = p.n_rows * p.Lh;
    if (warp < nW) {
        const int k = warp / (4 * p.M), c = warp - k * 4 * p.M, a = c / p.M, m = c - a * p.M;
        float sp = 0.f, sq = 0.f;
        for (int e = lane; e < total; e += 32) {
            const int n = e / p.Lh, j = e - n * p.Lh;
            const uint32_t *row = packed + (size_t)n * p.row_words;
            const size_t o = ((size_t)n * p.K + k) * p.Lh + j;
            if (base_at(row, j + m) == a) {
                sp += p.pf[o];
                if (p.q) sq += p.q[o];
            }
            if (p.pr && base_at(row, j + p.M - 1 - m) == 3 - a) sp += p.pr[o];
        }
        sp = warp_sum(sp);
        sq = warp_sum(sq);
        if (lane == 0) {
            p.vh_p[warp] = (p.accumulate ? p.vh_p[warp] : 0.f) + sp;
            if (p.vh_q) p.vh_q[warp] = (p.accumulate ? p.vh_q[warp] : 0.f) + sq;
        }
    } else {
        const int k = warp - nW;
        float sf = 0.f, sr = 0.f, sq = 0.f;
        for (int e = lane; e < total; e += 32) {
            const int n = e / p.Lh, j = e - n * p.Lh;
            const size_t o = ((size_t)n * p.K + k) * p.Lh + j;
            sf += p.pf[o];
            if (p.pr) sr += p.pr[o];
            if (p.q) sq += p.q[o];
        }
        sf = warp_sum(sf);
        sr = warp_sum(sr);
        sq = warp_sum(sq);
        if (lane == 0) {
            p.h_sum[k] = (p.accumulate ? p.h_sum[k] : 0.f) + (sf + sr);
            if (p.h_fwd) p.h_fwd[k] = (p.accumulate ? p.h_fwd[k] : 0.f) + sf;
            if (p.h_q) p.h_q[k] = (p.accumulate ? p.h_q[k] : 0.f) + sq;
        }
    }
}

}  // namespace crbm
