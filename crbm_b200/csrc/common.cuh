// Shared host/device helpers for the crbm_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/crbm_b200.h"

namespace crbm {

constexpr int FRONT_PAD_WORDS = 1;      // word 0 of every packed row is zero padding (16 bases)
constexpr int BACK_PAD_WORDS  = 7;      // kernels read up to 6 words starting at the word before a block's first position
constexpr int FANTASY_LEN     = CRBM_FANTASY_LEN;
constexpr int BLOCK_POS       = 16;     // hidden positions a warp processes per work item
constexpr int KTILE           = 32;     // motifs per k-tile: one lane per motif

// ---- host-side error plumbing ------------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
#define CRBM_CHECK_ARG(cond, ...) do { if (!(cond)) { crbm::set_error(__VA_ARGS__); return CRBM_ERR_ARG; } } while (0)
#define CRBM_CHECK_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { \
    crbm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    return CRBM_ERR_CUDA; } } while (0)
#define CRBM_CHECK_LAUNCH() do { crbm::count_launch(); CRBM_CHECK_CUDA(cudaPeekAtLastError()); } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// The kernels of a training step form a chain (h2v -> v2h -> reduce -> update -> next h2v ...). Launched with the
// programmatic-stream-serialization attribute, a kernel may start as soon as every CTA of its predecessor has executed
// griddepcontrol.launch_dependents; it runs its prologue (staging W, building fragments) and blocks in griddepcontrol.wait
// until the predecessor has completed and flushed. Stream capture records this as a programmatic graph edge.
// CRBM_PDL=0 turns the attribute off (the device-side instructions are then no-ops).
bool pdl_enabled();
template <typename P>
inline cudaError_t launch_chain(void (*kernel)(P), dim3 grid, dim3 block, size_t smem, cudaStream_t s, const P &params, bool pdl = true) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, params);
}
inline int packed_row_words(int L) { return 4 * ceil_div(FRONT_PAD_WORDS + ceil_div(L, 16) + BACK_PAD_WORDS, 4); }
int num_sms();

// ---- hidden bit planes (fantasy particles) --------------------------------------------------------------------------
// uint32 H[strand][n][kt][jb][16], jb = j/16. The 16 words of a (k-tile, position block) are the 16 warp ballots of the
// MMA accumulator fragments: word ((mt*2+h)*2+nt)*2+e, bit 4g+q  <->  motif 32kt+16mt+8h+g, position 16jb+8nt+2q+e.
constexpr int FANTASY_BLOCKS = (FANTASY_LEN + 15) / 16;
constexpr int FANTASY_WORDS = FANTASY_BLOCKS * 16;          // words per (strand, particle, k-tile)
__host__ __device__ inline int hbit_word(int j, int kl) {   // kl = motif index inside its k-tile (0..31)
    const int jb = j >> 4, nt = (j >> 3) & 1, e = j & 1, mt = kl >> 4, h = (kl >> 3) & 1;
    return jb * 16 + ((mt * 2 + h) * 2 + nt) * 2 + e;
}
__host__ __device__ inline int hbit_bit(int j, int kl) { return 4 * (kl & 7) + ((j >> 1) & 3); }

// ---- Philox4x32-10 (counter-based; Salmon et al. 2011) ----------------------------------------------------
__host__ __device__ inline uint4 philox4x32(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)M0 * ctr.x, p1 = (uint64_t)M1 * ctr.z;
        uint4 n;
        n.x = (uint32_t)(p1 >> 32) ^ ctr.y ^ key.x;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ ctr.w ^ key.y;
        n.w = (uint32_t)p0;
        ctr = n;
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__host__ __device__ inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }   // [0,1)

// Random stream tags (ctr.w high bits) so different consumers never share counters.
enum : uint32_t { RNG_BASES = 1, RNG_HIDDEN = 2, RNG_HIDDEN_RC = 3, RNG_VISIBLE = 4, RNG_EVAL = 5, RNG_API = 6 };

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- packed 2-bit sequence access -----------------------------------------------------------------------------
// Base i (i >= -16) of a row sits at bit 2*(i+16).
__device__ __forceinline__ int base_at(const uint32_t *__restrict__ row, int i) {
    int bit = 2 * (i + 16 * FRONT_PAD_WORDS);
    return (row[bit >> 5] >> (bit & 31)) & 3;
}

// Load NW consecutive 32-bit windows of bases starting at base `start` (>= -16): a[q] holds bases
// start+16q .. start+16q+15. All lanes of a warp pass the same arguments (broadcast loads).
template <int NW>
__device__ __forceinline__ void load_windows(const uint32_t *__restrict__ row, int start, uint32_t (&a)[NW]) {
    int bit = 2 * (start + 16 * FRONT_PAD_WORDS);
    const uint32_t *p = row + (bit >> 5);
    int sh = bit & 31;
    uint32_t lo = __ldg(p);
#pragma unroll
    for (int q = 0; q < NW; ++q) {
        uint32_t hi = __ldg(p + q + 1);
        a[q] = __funnelshift_r(lo, hi, sh);
        lo = hi;
    }
}

// 2*NB bits starting at (static) base offset t of the window array.
template <int NB, int NW>
__device__ __forceinline__ uint32_t window_bits(const uint32_t (&a)[NW], int t) {
    const int w = t >> 4, s = 2 * (t & 15);
    uint32_t v = (s + 2 * NB <= 32) ? (a[w] >> s) : __funnelshift_r(a[w], a[(w + 1 < NW) ? w + 1 : w], s);
    return v & ((1u << (2 * NB)) - 1u);
}

// floor(t) of 0 <= t < 2^23 in the low 23 bits of a float's pattern: t + 2^23 rounded toward zero has an ulp of one. A Bernoulli
// decision "floor(65536 p) > 16-bit uniform" is then one fma and one integer compare against (0x4B000000 | uniform) — no float-to-int
// conversion (F2I runs on the XU pipe, beside the exponentials), and the same decision bit for bit.
__device__ __forceinline__ uint32_t floor_bits(float v, float scale) {
    float x;
    asm("fma.rz.f32 %0, %1, %2, 0f4B000000;" : "=f"(x) : "f"(v), "f"(scale));
    return __float_as_uint(x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- 1-D bulk copy global -> shared through the TMA engine (cp.async.bulk), completion on an mbarrier ------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif  // __CUDACC__

}  // namespace crbm
