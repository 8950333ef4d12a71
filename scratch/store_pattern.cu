// Store-pattern microbenchmark for the scan output out[n][k][j] (rows of Lh=186 floats, K=100): which per-warp store
// shape sustains HBM write bandwidth? Each warp writes 16 rows x 32 positions per item, like v2h_dense_mma_kernel.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int K = 100, LH = 186;
template <int MODE>
__global__ void k(float *out, int n_seq) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int nblk2 = (LH + 31) / 32;
    const long n_items = (long)n_seq * nblk2;
    for (long item = blockIdx.x * nw + warp; item < n_items; item += (long)gridDim.x * nw) {
        const int n = item / nblk2, j0 = (item - (long)n * nblk2) * 32;
        for (int mt = 0; mt < 7; ++mt) {
            for (int h = 0; h < 2; ++h) {
                const int kk = mt * 16 + 8 * h + g;
                if (kk >= K) continue;
                float *row = out + ((size_t)n * K + kk) * LH + j0;
                if (MODE == 0) {            // current: per n-tile, quad writes 8 floats (float2 per lane)
                    for (int nt = 0; nt < 4; ++nt) {
                        const int j = 8 * nt + 2 * q;
                        if (j0 + j + 1 < LH) __stcs(reinterpret_cast<float2 *>(row + j), make_float2(1.f, 2.f));
                    }
                } else if (MODE == 1) {     // permuted columns: lane holds 8 consecutive positions -> 4 x float2 contiguous
                    for (int e = 0; e < 4; ++e) {
                        const int j = 8 * q + 2 * e;
                        if (j0 + j + 1 < LH) __stcs(reinterpret_cast<float2 *>(row + j), make_float2(1.f, 2.f));
                    }
                }
            }
            if (MODE == 2) {                // transposed through smem: one row of 32 positions per instruction
                for (int r = 0; r < 16; ++r) {
                    const int kk = mt * 16 + r;
                    if (kk >= K) continue;
                    float *row = out + ((size_t)n * K + kk) * LH + j0;
                    if (j0 + lane < LH) __stcs(row + lane, 1.f);
                }
            }
        }
    }
}
int main() {
    const int n_seq = 100000;
    const size_t bytes = (size_t)n_seq * K * LH * 4;
    float *d; cudaMalloc(&d, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 3, 256>>>(d, n_seq);
            if (mode == 1) k<1><<<148 * 3, 256>>>(d, n_seq);
            if (mode == 2) k<2><<<148 * 3, 256>>>(d, n_seq);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 2) printf("mode %d: %.3f ms, %.1f GB/s\n", mode, ms, bytes / ms / 1e6);
        }
    }
    return 0;
}
