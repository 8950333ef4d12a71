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
#include "v2h.cuh"

namespace crbm {

struct TcParams {
    DenseParams d;
    int NT;            // motifs per CTA (multiple of 16, <= 256)
    int ksteps;        // Kdim / 16
    int n_atoms;       // ceil(Kdim / 64): 128-byte swizzle atoms along K
    int tmem_stride;   // columns between the two accumulator buffers (power of two >= NT)
    int tmem_cols;     // 2 * tmem_stride
    int n_tiles;       // ceil(total / 128)
    float scale;       // -log2(e) for probabilities, 1 for activities
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int COLS_UNUSED = 0>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate; issued by one thread for the CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all tcgen05.mma issued so far by this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 bytes, 8-row groups
// 1024 bytes apart (stride byte offset), 16-byte chunks XOR-swizzled with the row index inside a group.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                              // leading byte offset (ignored for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `r` inside one swizzle atom column (rows x 128 bytes)
__device__ __forceinline__ uint32_t sw128_off(int r, int chunk) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((chunk ^ r) & 7) << 4));
}

constexpr int TC_THREADS = 256;
constexpr int TC_TILE = 128;
constexpr int A_ATOM_BYTES = TC_TILE * 128;

__host__ __device__ inline size_t tc_smem_bytes(int NT, int n_atoms) {
    return 1024 /*alignment slack*/ + (size_t)n_atoms * A_ATOM_BYTES + 2 * (size_t)n_atoms * NT * 128 + (size_t)NT * 4 + 64;
}

// (sequence, hidden position) of the first row of a tile; tiles advance by a fixed stride, so no 64-bit division is needed
struct TilePos {
    int n0, j0;
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
};

// One-hot A tile of the 128 hidden positions starting at `tp`: thread t builds the even (t < 128) or odd chunks of row t & 127.
__device__ __forceinline__ void tc_build_a(const TcParams &p, unsigned char *A, const TilePos &tp) {
    const DenseParams &d = p.d;
    const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
    uint32_t v[3] = {0u, 0u, 0u};
    int mlim = 0;
    int n, j;
    tp.row(r, d.Lh, n, j);
    if (n < d.n_rows) {
        const uint32_t *row = d.packed + (size_t)n * d.row_words;
        const int bit = 2 * (j + 16 * FRONT_PAD_WORDS);
        const uint32_t *w = row + (bit >> 5);
        const int sh = bit & 31;
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
        v[0] = __funnelshift_r(w0, w1, sh);               // bases j .. j+15
        v[1] = __funnelshift_r(w1, w2, sh);               // bases j+16 .. j+31
        v[2] = __funnelshift_r(w2, w3, sh);               // bases j+32 .. j+47
        mlim = d.M;
    }
    const int nch = 2 * p.ksteps;                         // 16-byte chunks per row: chunk ch holds m = 2ch, 2ch+1
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const int ch = 2 * i + half;
        if (ch < nch) {
            // bases 2ch, 2ch+1 sit in window (2ch >> 4) == (i >> 2) at bit 2 * (2ch & 15)
            const uint32_t bits = (v[i >> 2] >> (2 * ((4 * i) & 15))) >> (4 * half);
            const int b0 = bits & 3, b1 = (bits >> 2) & 3, m0 = 2 * ch;
            const uint32_t h0 = (m0 < mlim) ? (0x3C00u << (16 * (b0 & 1))) : 0u;      // fp16 1.0 in the half selected by the base
            const uint32_t h1 = (m0 + 1 < mlim) ? (0x3C00u << (16 * (b1 & 1))) : 0u;
            uint4 q;
            q.x = (b0 & 2) ? 0u : h0; q.y = (b0 & 2) ? h0 : 0u;
            q.z = (b1 & 2) ? 0u : h1; q.w = (b1 & 2) ? h1 : 0u;
            *reinterpret_cast<uint4 *>(A + (ch >> 3) * A_ATOM_BYTES + sw128_off(r, ch & 7)) = q;
        }
    }
}

__device__ __forceinline__ void tc_issue_mma(const TcParams &p, uint32_t a_addr, uint32_t bhi_addr, uint32_t blo_addr,
                                             uint32_t tmem_d, uint32_t idesc, uint64_t *bar) {
    const uint32_t b_atom = (uint32_t)p.NT * 128u;
    for (int s = 0; s < p.ksteps; ++s) {
        const uint32_t atom = s >> 2, koff = (s & 3) * 32;            // 16 fp16 = 32 bytes per k-step inside the 128-byte row
        const uint64_t ad = umma_desc_sw128(a_addr + atom * A_ATOM_BYTES + koff);
        umma_f16(tmem_d, ad, umma_desc_sw128(bhi_addr + atom * b_atom + koff), idesc, s > 0 ? 1u : 0u);
        umma_f16(tmem_d, ad, umma_desc_sw128(blo_addr + atom * b_atom + koff), idesc, 1u);
    }
    umma_commit(bar);
}

__global__ void __launch_bounds__(TC_THREADS, 2) v2h_dense_tc_kernel(const TcParams p) {
    const DenseParams &d = p.d;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char *A = smem;
    unsigned char *Bhi = A + (size_t)p.n_atoms * A_ATOM_BYTES;
    unsigned char *Blo = Bhi + (size_t)p.n_atoms * p.NT * 128;
    float *bias_s = reinterpret_cast<float *>(Blo + (size_t)p.n_atoms * p.NT * 128);
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + p.NT);       // NT multiple of 16 -> 8-byte aligned
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kbase = blockIdx.y * p.NT;
    const int Kdim = 16 * p.ksteps;

    if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    if (tid == 32) { mbar_init(bars, 1); mbar_init(bars + 1, 1); }

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

    int tile = blockIdx.x;
    const int stride_pos = TC_TILE * (int)gridDim.x;
    TilePos cur;                                                     // first row of `tile`
    cur.n0 = (tile * TC_TILE) / d.Lh;
    cur.j0 = tile * TC_TILE - cur.n0 * d.Lh;
    if (tile < p.n_tiles) tc_build_a(p, A, cur);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_addr = smem_u32(A), bhi_addr = smem_u32(Bhi), blo_addr = smem_u32(Blo);
    // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B fp16, both K-major, N = NT, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);
    if (tid == 0 && tile < p.n_tiles) tc_issue_mma(p, a_addr, bhi_addr, blo_addr, tmem_base, idesc, bars);

    const bool prob = d.out_kind == CRBM_OUT_PROBABILITY;
    const int q = warp & 3, chalf = warp >> 2;
    const int nchunks = p.NT >> 4;
    for (int it = 0; tile < p.n_tiles; ++it, tile += (int)gridDim.x) {
        const int buf = it & 1;
        mbar_wait(bars + buf, (uint32_t)((it >> 1) & 1));           // MMAs of this tile are done: A is free, D[buf] is full
        tc_fence_after();
        const int next = tile + (int)gridDim.x;
        TilePos nxt = cur;
        nxt.advance(stride_pos, d.Lh);
        if (next < p.n_tiles) {
            tc_build_a(p, A, nxt);
            fence_async_smem();
        }
        tc_fence_before();
        __syncthreads();                                             // also: everyone has drained D[buf ^ 1] (previous tile)
        if (tid == 0 && next < p.n_tiles) {
            tc_fence_after();
            tc_issue_mma(p, a_addr, bhi_addr, blo_addr, tmem_base + (uint32_t)((buf ^ 1) * p.tmem_stride), idesc, bars + (buf ^ 1));
        }
        // ---- epilogue of `tile` from D[buf]: this thread owns position row 32q + lane ------------------------------------------
        int n, j;
        cur.row(32 * q + lane, d.Lh, n, j);
        const bool valid = n < d.n_rows;
        float *orow = d.out + ((size_t)n * d.K + kbase) * d.Lh + j;
        const uint32_t trow = tmem_base + (uint32_t)(buf * p.tmem_stride) + ((uint32_t)(32 * q) << 16);
        for (int ch = chalf; ch < nchunks; ch += 2) {
            uint32_t v[16];
            tmem_ld16(trow + ch * 16, v);
            tmem_ld_wait();
            const int krem = d.K - kbase - ch * 16;                  // valid motifs in this chunk (warp-uniform)
            const float4 *bs4 = reinterpret_cast<const float4 *>(bias_s + ch * 16);
            float *o = orow + (size_t)(ch * 16) * d.Lh;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 b4 = bs4[c4];
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * c4 + e;
                    float x = __uint_as_float(v[c]) + bb[e];
                    if (prob) {
                        float t, rr;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr) : "f"(1.0f + t));
                        x = rr;
                    }
                    if (valid && c < krem) __stcs(o + (size_t)c * d.Lh, x);
                }
            }
        }
        cur = nxt;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

static int pow2_at_least(int x, int lo) { int v = lo; while (v < x) v <<= 1; return v; }

int launch_v2h_dense_tc(const DenseParams &d, cudaStream_t stream) {
    TcParams p{};
    p.d = d;
    const int Kdim = 16 * ceil_div(4 * d.M, 16);
    p.ksteps = Kdim / 16;
    p.n_atoms = ceil_div(Kdim, 64);
    const int Kpad = 16 * ceil_div(d.K, 16);
    // motifs per CTA: as many as fit in shared memory (B resident for the whole kernel) and in one MMA (N <= 256)
    int n_mtiles = ceil_div(Kpad, 256);
    int NT = 16 * ceil_div(ceil_div(Kpad, n_mtiles), 16);
    while (tc_smem_bytes(NT, p.n_atoms) > 200 * 1024 && NT > 16) {
        ++n_mtiles;
        NT = 16 * ceil_div(ceil_div(Kpad, n_mtiles), 16);
    }
    p.NT = NT;
    n_mtiles = ceil_div(d.K, NT);
    p.tmem_stride = pow2_at_least(NT, 32);
    p.tmem_cols = 2 * p.tmem_stride;
    const long long total = (long long)d.n_rows * d.Lh;             // hidden positions
    CRBM_CHECK_ARG((total + TC_TILE - 1) / TC_TILE < (1ll << 30), "v2h_dense_tc: more than 2^37 hidden positions in one call");
    p.n_tiles = (int)((total + TC_TILE - 1) / TC_TILE);
    p.scale = d.out_kind == CRBM_OUT_PROBABILITY ? -1.4426950408889634f : 1.0f;
    const size_t smem = tc_smem_bytes(NT, p.n_atoms);
    static size_t configured = 0;
    if (smem > configured) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(v2h_dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > 512 / p.tmem_cols) per_sm = 512 / p.tmem_cols;
    if (per_sm > 2) per_sm = 2;
    if (per_sm < 1) per_sm = 1;
    int gx = (per_sm * num_sms()) / n_mtiles;
    if (gx < 1) gx = 1;
    if (gx > p.n_tiles) gx = p.n_tiles;
    v2h_dense_tc_kernel<<<dim3(gx, n_mtiles), TC_THREADS, smem, stream>>>(p);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

}  // namespace crbm
