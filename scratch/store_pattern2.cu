// Write-bandwidth microbenchmark for the scan output out[n][k][j] (K = 100 rows of LH floats per sequence): which store
// shape reaches the HBM write ceiling? No arithmetic. Tiles of 128 consecutive positions of one sequence x all K rows.
//   mode 0: ceiling — every warp writes 512 contiguous, 16-byte aligned bytes per instruction (float4), flat array
//   mode 1: scalar row pieces: warp writes 32 consecutive floats of one row per instruction (tcgen05 epilogue, round 1)
//   mode 2: scalar, 4 consecutive instructions cover one 128-position run of a row (better locality per warp)
//   mode 3: float4 body + scalar head/tail per 128-position run (needs a staging transpose in the real kernel)
//   mode 4: cp.async.bulk shared -> global, one 512-byte (aligned part of a) run per instruction, issued by lane 0
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int K = 100;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int n_seq, int LH) {
    __shared__ __align__(128) float stage[8][136];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (MODE == 0) {
        const size_t total4 = (size_t)n_seq * K * LH / 4;
        float4 *o = reinterpret_cast<float4 *>(out);
        for (size_t i = (size_t)(blockIdx.x * nw + warp) * 32 + lane; i < total4; i += (size_t)gridDim.x * nw * 32)
            __stcs(o + i, make_float4(1.f, 2.f, 3.f, 4.f));
        return;
    }
    for (int i = lane; i < 136; i += 32) stage[warp][i] = 1.f;
    __syncwarp();
    const int tiles_per_seq = (LH + 127) / 128;
    const long n_tiles = (long)n_seq * tiles_per_seq;
    // a CTA takes a tile; its 8 warps split the K rows
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_seq, j0 = (tile - (long)n * tiles_per_seq) * 128;
        const int len = min(128, LH - j0);
        for (int kk = warp; kk < K; kk += nw) {
            float *row = out + ((size_t)n * K + kk) * LH + j0;
            if (MODE == 1) {
                // 4 pieces of 32, but interleaved over rows like the TMEM epilogue: piece q of row kk
                for (int q = 0; q < 4; ++q)
                    if (32 * q + lane < len) __stcs(row + 32 * q + lane, 1.f);
            } else if (MODE == 2) {
                for (int q = 0; q < 4; ++q)
                    if (32 * q + lane < len) __stcs(row + 32 * q + lane, 1.f);
            } else if (MODE == 3) {
                const int a = (int)(((uintptr_t)row >> 2) & 3);        // misalignment in floats
                const int head = (4 - a) & 3;
                if (lane < head && lane < len) __stcs(row + lane, 1.f);
                const int body4 = (len - head) >> 2;
                float4 *b = reinterpret_cast<float4 *>(row + head);
                if (lane < body4) __stcs(b + lane, make_float4(1.f, 2.f, 3.f, 4.f));
                const int tail0 = head + 4 * body4;
                if (tail0 + lane < len) __stcs(row + tail0 + lane, 1.f);
            } else if (MODE == 4) {
                const int a = (int)(((uintptr_t)row >> 2) & 3);
                const int head = (4 - a) & 3;
                if (lane < head && lane < len) __stcs(row + lane, 1.f);
                const int body4 = (len - head) >> 2;
                if (lane == 0 && body4 > 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(row + head),
                                 "r"(smem_u32(&stage[warp][0])), "r"(body4 * 16) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                const int tail0 = head + 4 * body4;
                if (tail0 + lane < len) __stcs(row + tail0 + lane, 1.f);
            }
        }
        if (MODE == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    if (MODE == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE> void run(float *d, int n_seq, int LH, int ctas_per_sm, const char *name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t bytes = (size_t)n_seq * K * LH * 4;
    float ms = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<148 * ctas_per_sm, 256>>>(d, n_seq, LH);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("LH=%d ctas/SM=%d mode %d (%s): %.3f ms, %.0f GB/s  [%s]\n", LH, ctas_per_sm, MODE, name, ms, bytes / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int n_seq = 100000;
    float *d; cudaMalloc(&d, (size_t)n_seq * K * 192 * 4 + 1024);
    cudaMemset(d, 0, (size_t)n_seq * K * 192 * 4);
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float ms;
        for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); cudaMemsetAsync(d, 1, (size_t)n_seq * K * 186 * 4); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
        printf("cudaMemset: %.3f ms, %.0f GB/s\n", ms, (double)n_seq * K * 186 * 4 / ms / 1e6);
    }
    for (int LH : {186, 188, 192}) {
        for (int c : {2, 4}) {
            run<0>(d, n_seq, LH, c, "float4 flat ceiling");
            run<1>(d, n_seq, LH, c, "scalar row pieces");
            run<3>(d, n_seq, LH, c, "float4 body + scalar head/tail");
            run<4>(d, n_seq, LH, c, "bulk copy smem->global body");
        }
    }
    return 0;
}
