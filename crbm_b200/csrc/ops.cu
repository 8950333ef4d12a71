// Sequence ingest (A0) and the dense "eager builder" ops (A2, A3, A5, A6) of the reference, as plain
// coalesced elementwise / gather kernels. None of these is on the training hot path; they exist so that the
// reference's builder-level API (`_bottomUpProbability`, `_bottomUpSample`, `_topDownActivity`, ...) is available
// eagerly on the GPU and so that tests can drive every stage separately.
#include "common.cuh"

namespace crbm {

// ---- A0: pack ------------------------------------------------------------------------------------------------
// One thread per output word (16 bases). one-hot layout (N,1,4,L).
__global__ void pack_onehot_kernel(const float *__restrict__ oh, int N, int L, uint32_t *__restrict__ packed,
                                   int row_words, int *__restrict__ bad) {
    const int wpr = row_words;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)N * wpr) return;
    const int n = (int)(gid / wpr), w = (int)(gid - (long long)n * wpr);
    uint32_t word = 0u;
    int nbad = 0;
    const int i0 = (w - FRONT_PAD_WORDS) * 16;
    if (w >= FRONT_PAD_WORDS && i0 < L) {
        const float *base = oh + (size_t)n * 4 * L;
        for (int q = 0; q < 16; ++q) {
            const int i = i0 + q;
            if (i >= L) break;
            const float a = base[i], c = base[L + i], g = base[2 * L + i], t = base[3 * L + i];
            const int ones = (a == 1.f) + (c == 1.f) + (g == 1.f) + (t == 1.f);
            const int zeros = (a == 0.f) + (c == 0.f) + (g == 0.f) + (t == 0.f);
            if (ones != 1 || zeros != 3) ++nbad;
            const uint32_t code = (c == 1.f) ? 1u : (g == 1.f) ? 2u : (t == 1.f) ? 3u : 0u;
            word |= code << (2 * q);
        }
    }
    packed[gid] = word;
    if (nbad) atomicAdd(bad, nbad);
}

__global__ void pack_codes_kernel(const uint8_t *__restrict__ codes, int N, int L, uint32_t *__restrict__ packed,
                                  int row_words, int *__restrict__ bad) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)N * row_words) return;
    const int n = (int)(gid / row_words), w = (int)(gid - (long long)n * row_words);
    uint32_t word = 0u;
    int nbad = 0;
    const int i0 = (w - FRONT_PAD_WORDS) * 16;
    if (w >= FRONT_PAD_WORDS && i0 < L) {
        const uint8_t *src = codes + (size_t)n * L;
        for (int q = 0; q < 16; ++q) {
            const int i = i0 + q;
            if (i >= L) break;
            const uint32_t c = src[i];
            if (c > 3u) ++nbad;
            word |= (c & 3u) << (2 * q);
        }
    }
    packed[gid] = word;
    if (nbad) atomicAdd(bad, nbad);
}

__global__ void unpack_onehot_kernel(const uint32_t *__restrict__ packed, int row_words, int N, int L,
                                     float *__restrict__ oh) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)N * 4 * L) return;
    const int i = (int)(gid % L);
    const int a = (int)((gid / L) % 4);
    const int n = (int)(gid / (4LL * L));
    oh[gid] = (base_at(packed + (size_t)n * row_words, i) == a) ? 1.f : 0.f;
}

__global__ void random_bases_kernel(uint32_t *__restrict__ packed, int row_words, int N, int L, uint64_t seed) {
    // one thread per 4 words (Philox yields 4 x 32 random bits = 64 bases)
    const int quads = row_words / 4;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)N * quads) return;
    const int n = (int)(gid / quads), q = (int)(gid - (long long)n * quads);
    const uint4 r = philox4x32(make_uint4((uint32_t)n, (uint32_t)q, 0u, RNG_BASES),
                               make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int word = 4 * q + e;
        const int i0 = (word - FRONT_PAD_WORDS) * 16;
        uint32_t v = w[e];
        if (word < FRONT_PAD_WORDS || i0 >= L) v = 0u;
        else if (i0 + 16 > L) v &= (1u << (2 * (L - i0))) - 1u;
        packed[(size_t)n * row_words + word] = v;
    }
}

// ---- A2 on a dense activity tensor ----------------------------------------------------------------------------
__global__ void pool_probability_kernel(const float *__restrict__ act, long long ngroups, int pooling,
                                        float *__restrict__ prob) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const float *x = act + g * pooling;
    float mx = 0.f;
    for (int i = 0; i < pooling; ++i) mx = fmaxf(mx, x[i]);
    float tot = 0.f;
    for (int i = 0; i < pooling; ++i) tot += __expf(x[i] - mx);
    const float inv = 1.f / ((float)pooling * __expf(-mx) + tot);
    for (int i = 0; i < pooling; ++i) prob[g * pooling + i] = __expf(x[i] - mx) * inv;
}

// ---- A3 on a dense probability tensor -------------------------------------------------------------------------
__global__ void sample_hidden_kernel(const float *__restrict__ prob, long long ngroups, int pooling,
                                     const float *__restrict__ uniforms, uint64_t seed, uint64_t offset,
                                     float *__restrict__ sample) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    float u;
    if (uniforms) u = uniforms[g];
    else {
        const uint64_t c = (uint64_t)g + offset;
        u = u01(philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, RNG_API),
                           make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))).x);
    }
    float cum = 0.f;
    bool done = false;
    for (int i = 0; i < pooling; ++i) {
        cum += prob[g * pooling + i];
        const bool s = !done && cum > u;
        done = done || s;
        sample[g * pooling + i] = s ? 1.f : 0.f;
    }
}

// ---- A5 dense: one thread per visible position, real-valued hidden input -------------------------------------------
__global__ void h2v_dense_kernel(const float *__restrict__ h, const float *__restrict__ hp, int N, int K, int Lh,
                                 const float *__restrict__ W, const float *__restrict__ c, int M,
                                 float *__restrict__ act) {
    const int Lv = Lh + M - 1;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)N * Lv) return;
    const int n = (int)(gid / Lv), i = (int)(gid - (long long)n * Lv);
    float v0 = c[0], v1 = c[1], v2 = c[2], v3 = c[3];
    const int mlo = max(0, i - (Lh - 1)), mhi = min(M - 1, i);
    for (int k = 0; k < K; ++k) {
        const float *hk = h + ((size_t)n * K + k) * Lh;
        const float *hpk = hp ? hp + ((size_t)n * K + k) * Lh : nullptr;
        const float *w = W + (size_t)k * 4 * M;
        for (int m = mlo; m <= mhi; ++m) {
            const float hv = hk[i - m];
            if (hv != 0.f) {
                v0 += hv * w[m]; v1 += hv * w[M + m]; v2 += hv * w[2 * M + m]; v3 += hv * w[3 * M + m];
            }
            if (hpk) {
                const float hv2 = hpk[i - m];
                if (hv2 != 0.f) {   // W[k, 3-a, M-1-m]
                    const int mm = M - 1 - m;
                    v0 += hv2 * w[3 * M + mm]; v1 += hv2 * w[2 * M + mm]; v2 += hv2 * w[M + mm]; v3 += hv2 * w[mm];
                }
            }
        }
    }
    float *o = act + (size_t)n * 4 * Lv + i;
    o[0] = v0; o[Lv] = v1; o[2 * Lv] = v2; o[3 * Lv] = v3;
}

// ---- A6 dense ---------------------------------------------------------------------------------------------------
__global__ void visible_softmax_sample_kernel(const float *__restrict__ act, int N, int Lv, float *__restrict__ prob,
                                              const float *__restrict__ uniforms, uint64_t seed, uint64_t offset,
                                              float *__restrict__ sample) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)N * Lv) return;
    const int n = (int)(gid / Lv), i = (int)(gid - (long long)n * Lv);
    const float *a = act + (size_t)n * 4 * Lv + i;
    const float v0 = a[0], v1 = a[Lv], v2 = a[2 * Lv], v3 = a[3 * Lv];
    const float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
    const float e0 = __expf(v0 - mx), e1 = __expf(v1 - mx), e2 = __expf(v2 - mx), e3 = __expf(v3 - mx);
    const float inv = 1.f / (e0 + e1 + e2 + e3);
    const float p0 = e0 * inv, p1 = e1 * inv, p2 = e2 * inv, p3 = e3 * inv;
    if (prob) {
        float *o = prob + (size_t)n * 4 * Lv + i;
        o[0] = p0; o[Lv] = p1; o[2 * Lv] = p2; o[3 * Lv] = p3;
    }
    if (sample) {
        float u;
        if (uniforms) u = uniforms[gid];
        else {
            const uint64_t cc = (uint64_t)gid + offset;
            u = u01(philox4x32(make_uint4((uint32_t)cc, (uint32_t)(cc >> 32), 1u, RNG_API),
                               make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))).x);
        }
        // first base whose running sum exceeds u; rounding can leave the sum a hair below u -> last base
        const int pick = (p0 > u) ? 0 : (p0 + p1 > u) ? 1 : (p0 + p1 + p2 > u) ? 2 : 3;
        float *o = sample + (size_t)n * 4 * Lv + i;
        o[0] = pick == 0; o[Lv] = pick == 1; o[2 * Lv] = pick == 2; o[3 * Lv] = pick == 3;
    }
}

// ---- fantasy particles: bit planes [n][kt][FANTASY_WORDS] (layout: common.cuh) <-> dense (B,K,1,200) -----------------
__global__ void bits_to_dense_kernel(const uint32_t *__restrict__ bits, int Bn, int K, int kw, float *__restrict__ dense) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)Bn * K * FANTASY_LEN) return;
    const int j = (int)(gid % FANTASY_LEN);
    const int k = (int)((gid / FANTASY_LEN) % K);
    const int n = (int)(gid / ((long long)FANTASY_LEN * K));
    const uint32_t w = bits[((size_t)n * kw + (k >> 5)) * FANTASY_WORDS + hbit_word(j, k & 31)];
    dense[gid] = (float)((w >> hbit_bit(j, k & 31)) & 1u);
}

__global__ void dense_to_bits_kernel(const float *__restrict__ dense, int Bn, int K, int kw, uint32_t *__restrict__ bits) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)Bn * kw * FANTASY_WORDS) return;
    const int wi = (int)(gid % FANTASY_WORDS);
    const int t = (int)((gid / FANTASY_WORDS) % kw);
    const int n = (int)(gid / ((long long)FANTASY_WORDS * kw));
    const int jb = wi >> 4, f = wi & 15, mt = f >> 3, h = (f >> 2) & 1, nt = (f >> 1) & 1, e = f & 1;
    uint32_t w = 0u;
    for (int b = 0; b < 32; ++b) {
        const int g = b >> 2, q = b & 3;
        const int k = t * 32 + 16 * mt + 8 * h + g, j = 16 * jb + 8 * nt + 2 * q + e;
        if (k < K && j < FANTASY_LEN && dense[((size_t)n * K + k) * FANTASY_LEN + j] != 0.f) w |= 1u << b;
    }
    bits[gid] = w;
}

// ---- sum over sequences of a dense (n,K,1,Lh) chunk (rows `row_stride` floats apart), accumulated into acc[K][Lh] ------------
// One thread per (motif, position): consecutive threads read consecutive positions of one row (coalesced), a thread walks the
// sequences of the chunk in order, so the sum is deterministic. Four independent partial sums hide the load latency.
__global__ void position_sum_kernel(const float *__restrict__ dense, int n, int K, int Lh, long long row_stride, float *__restrict__ acc) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)K * Lh) return;
    const int k = (int)(gid / Lh), j = (int)(gid - (long long)k * Lh);
    const float *src = dense + (size_t)k * row_stride + j;
    const size_t seq_stride = (size_t)K * row_stride;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
        a0 += src[(size_t)i * seq_stride];
        a1 += src[(size_t)(i + 1) * seq_stride];
        a2 += src[(size_t)(i + 2) * seq_stride];
        a3 += src[(size_t)(i + 3) * seq_stride];
    }
    for (; i < n; ++i) a0 += src[(size_t)i * seq_stride];
    acc[gid] += (a0 + a1) + (a2 + a3);
}

static inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace crbm

using namespace crbm;

extern "C" {

static int pack_onehot_impl(const float *onehot, int N, int L, uint32_t *packed, int row_words, int *bad_count, void *stream, bool reset) {
    CRBM_CHECK_ARG(onehot && packed && bad_count && N > 0 && L > 0, "crbm_pack_onehot: bad arguments");
    CRBM_CHECK_ARG(row_words >= packed_row_words(L), "crbm_pack_onehot: row_words %d < %d", row_words, packed_row_words(L));
    cudaStream_t s = (cudaStream_t)stream;
    if (reset) CRBM_CHECK_CUDA(cudaMemsetAsync(bad_count, 0, sizeof(int), s));
    pack_onehot_kernel<<<blocks_for((long long)N * row_words, 256), 256, 0, s>>>(onehot, N, L, packed, row_words, bad_count);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_pack_onehot(const float *onehot, int N, int L, uint32_t *packed, int row_words, int *bad_count, void *stream) {
    return pack_onehot_impl(onehot, N, L, packed, row_words, bad_count, stream, true);
}
int crbm_pack_onehot_accumulate(const float *onehot, int N, int L, uint32_t *packed, int row_words, int *bad_count, void *stream) {
    return pack_onehot_impl(onehot, N, L, packed, row_words, bad_count, stream, false);
}

int crbm_pack_codes(const uint8_t *codes, int N, int L, uint32_t *packed, int row_words, int *bad_count, void *stream) {
    CRBM_CHECK_ARG(codes && packed && bad_count && N > 0 && L > 0, "crbm_pack_codes: bad arguments");
    CRBM_CHECK_ARG(row_words >= packed_row_words(L), "crbm_pack_codes: row_words %d < %d", row_words, packed_row_words(L));
    cudaStream_t s = (cudaStream_t)stream;
    CRBM_CHECK_CUDA(cudaMemsetAsync(bad_count, 0, sizeof(int), s));
    pack_codes_kernel<<<blocks_for((long long)N * row_words, 256), 256, 0, s>>>(codes, N, L, packed, row_words, bad_count);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_unpack_onehot(const uint32_t *packed, int row_words, int N, int L, float *onehot, void *stream) {
    CRBM_CHECK_ARG(onehot && packed && N > 0 && L > 0, "crbm_unpack_onehot: bad arguments");
    unpack_onehot_kernel<<<blocks_for((long long)N * 4 * L, 256), 256, 0, (cudaStream_t)stream>>>(packed, row_words, N, L, onehot);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_random_bases(uint32_t *packed, int row_words, int N, int L, uint64_t seed, void *stream) {
    CRBM_CHECK_ARG(packed && N > 0 && L > 0, "crbm_random_bases: bad arguments");
    CRBM_CHECK_ARG(row_words >= packed_row_words(L) && row_words % 4 == 0, "crbm_random_bases: bad row_words %d", row_words);
    random_bases_kernel<<<blocks_for((long long)N * (row_words / 4), 256), 256, 0, (cudaStream_t)stream>>>(packed, row_words, N, L, seed);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_position_sum(const float *dense, int n, int K, int Lh, long long row_stride, float *acc, void *stream) {
    CRBM_CHECK_ARG(dense && acc && n > 0 && K > 0 && Lh > 0 && row_stride >= Lh, "crbm_position_sum: bad arguments");
    position_sum_kernel<<<blocks_for((long long)K * Lh, 128), 128, 0, (cudaStream_t)stream>>>(dense, n, K, Lh, row_stride, acc);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_pool_probability(const float *act, int N, int K, int Lh, int pooling, float *prob, void *stream) {
    CRBM_CHECK_ARG(act && prob && N > 0 && K > 0 && Lh > 0 && pooling > 0, "crbm_pool_probability: bad arguments");
    CRBM_CHECK_ARG(Lh % pooling == 0, "crbm_pool_probability: hidden length %d not divisible by pooling %d", Lh, pooling);
    const long long ng = (long long)N * K * (Lh / pooling);
    pool_probability_kernel<<<blocks_for(ng, 256), 256, 0, (cudaStream_t)stream>>>(act, ng, pooling, prob);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_sample_hidden(const float *prob, int N, int K, int Lh, int pooling, const float *uniforms,
                       uint64_t seed, uint64_t offset, float *sample, void *stream) {
    CRBM_CHECK_ARG(prob && sample && N > 0 && K > 0 && Lh > 0 && pooling > 0, "crbm_sample_hidden: bad arguments");
    CRBM_CHECK_ARG(Lh % pooling == 0, "crbm_sample_hidden: hidden length %d not divisible by pooling %d", Lh, pooling);
    const long long ng = (long long)N * K * (Lh / pooling);
    sample_hidden_kernel<<<blocks_for(ng, 256), 256, 0, (cudaStream_t)stream>>>(prob, ng, pooling, uniforms, seed, offset, sample);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_h2v_dense(const float *h, const float *hprime, int N, int K, int Lh, const float *W, const float *c, int M,
                   float *act, void *stream) {
    CRBM_CHECK_ARG(h && W && c && act && N > 0 && K > 0 && Lh > 0 && M > 0, "crbm_h2v_dense: bad arguments");
    const long long n = (long long)N * (Lh + M - 1);
    h2v_dense_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(h, hprime, N, K, Lh, W, c, M, act);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_visible_softmax_sample(const float *act, int N, int Lv, float *prob, const float *uniforms,
                                uint64_t seed, uint64_t offset, float *sample, void *stream) {
    CRBM_CHECK_ARG(act && N > 0 && Lv > 0 && (prob || sample), "crbm_visible_softmax_sample: bad arguments");
    visible_softmax_sample_kernel<<<blocks_for((long long)N * Lv, 256), 256, 0, (cudaStream_t)stream>>>(
        act, N, Lv, prob, uniforms, seed, offset, sample);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_hidden_bits_to_dense(const uint32_t *bits, int B, int K, float *dense, void *stream) {
    CRBM_CHECK_ARG(bits && dense && B > 0 && K > 0, "crbm_hidden_bits_to_dense: bad arguments");
    bits_to_dense_kernel<<<blocks_for((long long)B * K * FANTASY_LEN, 256), 256, 0, (cudaStream_t)stream>>>(
        bits, B, K, ceil_div(K, 32), dense);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

int crbm_hidden_dense_to_bits(const float *dense, int B, int K, uint32_t *bits, void *stream) {
    CRBM_CHECK_ARG(bits && dense && B > 0 && K > 0, "crbm_hidden_dense_to_bits: bad arguments");
    const int kw = ceil_div(K, 32);
    dense_to_bits_kernel<<<blocks_for((long long)B * kw * FANTASY_WORDS, 256), 256, 0, (cudaStream_t)stream>>>(
        dense, B, K, kw, bits);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

}  // extern "C"
