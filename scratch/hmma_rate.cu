// Microbenchmark: legacy mma.sync.m16n8k16 bf16 throughput on sm_100a (per SM), to size the fused cRBM kernels.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__global__ void k(float *out, int iters, int nacc_unused) {
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 0x3f803f80, b1 = 0x3f800000;
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *d; cudaMalloc(&d, 148 * 1024 * 4);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000;
        k<<<148, warps * 32>>>(d, 100, 0);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<148, warps * 32>>>(d, iters, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mmas = 148.0 * warps * iters * 8;
        printf("warps/SM=%2d: %.3f ms, %.2f HMMA.16816/ns total, %.1f TFLOP/s, cycles per HMMA per SM @1.965GHz = %.2f\n", warps, ms,
               mmas / (ms * 1e6), mmas * 4096 * 2 / (ms * 1e-3) / 1e12, (ms * 1e-3 * 1.965e9) / (warps * iters * 8.0));
    }
    return 0;
}
