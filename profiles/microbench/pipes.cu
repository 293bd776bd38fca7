// Microbenchmark of the sm_100a issue/pipe rates the resample kernels depend on:
// FFMA vs FFMA2 (fma.rn.f32x2), I2F.U8 byte->float, PRMT+FADD magic conversion,
// and the cost of predicated-off FFMA2.  Prints ops per clock per SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2 &acc, float2 a, float w) {
    float2 b = make_float2(w, w);
    unsigned long long &A = reinterpret_cast<unsigned long long &>(acc);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(b)));
}

constexpr int ITERS = 4096;

template <int MODE>
__global__ void bench(float *out, const uint32_t *in, long long *cycles, uint32_t pmask) {
    float2 acc[16];
    float sc[32];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
#pragma unroll
    for (int i = 0; i < 32; i++) sc[i] = threadIdx.x * 0.001f + i;
    uint32_t word = in[threadIdx.x & 31];
    float w = 1.0001f + in[32] * 1e-9f;
    float2 v = make_float2(0.5f + in[33], 0.25f);
    float conv = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {  // 32 FFMA
#pragma unroll
            for (int i = 0; i < 32; i++) sc[i] = fmaf(sc[i], w, v.x);
        } else if (MODE == 1) {  // 16 FFMA2 (= 32 FMA)
#pragma unroll
            for (int i = 0; i < 16; i++) ffma2(acc[i], v, w);
        } else if (MODE == 2) {  // 16 I2F.U8
#pragma unroll
            for (int i = 0; i < 4; i++) {
                conv += (float)(word & 0xff); conv += (float)((word >> 8) & 0xff);
                conv += (float)((word >> 16) & 0xff); conv += (float)(word >> 24);
                word = word * 1664525u + 1013904223u;
            }
        } else if (MODE == 3) {  // 16 FFMA2 + 8 I2F.U8 (the V-stage mix per 8 bytes, 4 live slots)
            float f[8];
            f[0] = (float)(word & 0xff); f[1] = (float)((word >> 8) & 0xff); f[2] = (float)((word >> 16) & 0xff); f[3] = (float)(word >> 24);
            uint32_t w2 = word ^ 0x5a5a5a5a;
            f[4] = (float)(w2 & 0xff); f[5] = (float)((w2 >> 8) & 0xff); f[6] = (float)((w2 >> 16) & 0xff); f[7] = (float)(w2 >> 24);
#pragma unroll
            for (int i = 0; i < 16; i++) ffma2(acc[i], make_float2(f[(2 * i) & 7], f[(2 * i + 1) & 7]), w);
            word += 0x01010101u;
        } else if (MODE == 4) {  // 16 predicated FFMA2, predicate from pmask (runtime): off when pmask == 0
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (pmask & (1u << (i & 7))) ffma2(acc[i], v, w);
        } else if (MODE == 5) {  // PRMT magic + FADD: 8 conversions
            float f[8];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                f[b] = __uint_as_float(__byte_perm(word, 0x4b000000u, 0x7440 + b)) - 8388608.0f;
                f[4 + b] = __uint_as_float(__byte_perm(word ^ 0x3c3c3c3cu, 0x4b000000u, 0x7440 + b)) - 8388608.0f;
            }
#pragma unroll
            for (int b = 0; b < 8; b++) conv += f[b];
            word += 0x01010101u;
        } else if (MODE == 6) {  // 32 FFMA2 + 8 I2F.U8 (8 live slots x 8 bytes)
            float f[8];
            f[0] = (float)(word & 0xff); f[1] = (float)((word >> 8) & 0xff); f[2] = (float)((word >> 16) & 0xff); f[3] = (float)(word >> 24);
            uint32_t w2 = word ^ 0x5a5a5a5a;
            f[4] = (float)(w2 & 0xff); f[5] = (float)((w2 >> 8) & 0xff); f[6] = (float)((w2 >> 16) & 0xff); f[7] = (float)(w2 >> 24);
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) ffma2(acc[i], make_float2(f[(2 * i) & 7], f[(2 * i + 1) & 7]), w);
            word += 0x01010101u;
        }
    }
    long long t1 = clock64();
    float s = conv;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
#pragma unroll
    for (int i = 0; i < 32; i++) s += sc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, double ops_per_iter_per_thread, int threads, uint32_t pmask, float *out, uint32_t *in, long long *cyc) {
    bench<MODE><<<148, threads>>>(out, in, cyc, pmask);
    cudaDeviceSynchronize();
    bench<MODE><<<148, threads>>>(out, in, cyc, pmask);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += h[i];
    avg /= 148;
    printf("%-44s threads=%4d cycles=%9.0f  ops/clk/SM=%8.2f  clk/warp-iter/SMSP=%7.2f\n", name, threads, avg,
           ops_per_iter_per_thread * ITERS * threads / avg, avg / ITERS / (threads / 128.0));
}

int main() {
    float *out; uint32_t *in; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 64 * 4); cudaMalloc(&cyc, 148 * 8);
    cudaMemset(in, 0, 64 * 4);
    for (int threads : {128, 256, 512}) {
        run<0>("32 FFMA            (ops = FMA)", 32, threads, 0xff, out, in, cyc);
        run<1>("16 FFMA2           (ops = FMA)", 32, threads, 0xff, out, in, cyc);
        run<2>("16 I2F.U8          (ops = cvt)", 16, threads, 0xff, out, in, cyc);
        run<5>("8 PRMT+FADD magic  (ops = cvt)", 8, threads, 0xff, out, in, cyc);
        run<3>("16 FFMA2 + 8 I2F.U8 (ops = FMA)", 32, threads, 0xff, out, in, cyc);
        run<6>("32 FFMA2 + 8 I2F.U8 (ops = FMA)", 64, threads, 0xff, out, in, cyc);
        run<4>("16 FFMA2 all predicated ON  (ops = FMA)", 32, threads, 0xff, out, in, cyc);
        run<4>("16 FFMA2 half predicated OFF (ops = FMA issued)", 32, threads, 0x0f, out, in, cyc);
        run<4>("16 FFMA2 all predicated OFF (ops = FMA issued)", 32, threads, 0x00, out, in, cyc);
    }
    return 0;
}
