// Flat contiguous array written in 128-byte warp pieces (32 lanes x 4 B) that start `off` bytes after a 128-byte boundary:
// which misalignment granularity (line / sector / sub-sector) costs write bandwidth on B200?
// Also: the same pieces written in a row order that delays the neighbouring piece (stride between consecutive pieces of a warp).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(float *out, size_t n_pieces, int off_floats, int vec) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float *base = out + off_floats;
    if (vec == 1) {
        for (size_t p = (size_t)blockIdx.x * nw + warp; p < n_pieces; p += (size_t)gridDim.x * nw)
            __stcs(base + p * 32 + lane, 1.f);
    } else {       // float4: 512-byte pieces (needs off % 16 bytes == 0)
        for (size_t p = (size_t)blockIdx.x * nw + warp; p < n_pieces / 4; p += (size_t)gridDim.x * nw)
            __stcs(reinterpret_cast<float4 *>(base + p * 128) + lane, make_float4(1.f, 2.f, 3.f, 4.f));
    }
}
int main() {
    const size_t bytes = (size_t)7 << 30;
    float *d; cudaMalloc(&d, bytes + 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t n_pieces = bytes / 128;
    for (int vec : {1, 4})
        for (int off : {0, 64, 32, 16, 8, 4}) {
            if (vec == 4 && off % 16) continue;
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                k<<<148 * 4, 256>>>(d, n_pieces, off / 4, vec);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            printf("vec=%d offset %3d B: %.3f ms, %.0f GB/s [%s]\n", vec, off, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
