// Microbenchmark: throughput of the legacy warp-level int8 tensor-core instruction
// (mma.sync.m16n8k32.s32.u8.s8 -> SASS IMMA.16832.U8.S8) on sm_100a.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
__global__ void bench(const uint32_t *a, const uint32_t *b, int *d, long long *cycles) {
    uint32_t a0 = a[threadIdx.x & 31], a1 = a0 ^ 0x01020304u, a2 = a0 + 7, a3 = a0 * 3;
    uint32_t b0 = b[threadIdx.x & 31], b1 = b0 ^ 0x7f7f7f7fu;
    int c[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) c[j][k] = j + k;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) s += c[j][k];
    d[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
int main() {
    uint32_t *a, *b; int *d; long long *cyc;
    cudaMalloc(&a, 128); cudaMalloc(&b, 128); cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    cudaMemset(a, 1, 128); cudaMemset(b, 2, 128);
    for (int threads : {128, 256, 512}) {
        for (int r = 0; r < 2; r++) { bench<<<148, threads>>>(a, b, d, cyc); cudaDeviceSynchronize(); }
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; i++) avg += h[i];
        avg /= 148;
        double immas = 8.0 * ITERS * (threads / 32);
        printf("threads=%4d cycles=%9.0f  IMMA.16832 per clk per SM=%6.3f  MAC/clk/SM=%8.1f  clk per IMMA per SMSP=%6.2f\n", threads, avg,
               immas / avg, immas * 4096 / avg, avg / (8.0 * ITERS * (threads / 128.0)));
    }
    return 0;
}
