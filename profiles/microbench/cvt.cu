// Microbenchmark: ways to turn 8 packed bytes into 8 floats on sm_100a, alone and
// mixed with the 24 FFMA2 (6 live output rows x 8 bytes) of the vertical pass.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2 &acc, float2 a, float w) {
    float2 b = make_float2(w, w);
    unsigned long long &A = reinterpret_cast<unsigned long long &>(acc);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(b)));
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long &>(r)) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(b)));
    return r;
}

constexpr int ITERS = 4096;

// CVT: 0 = I2F.U8, 1 = PRMT extract + I2FP.F32.U32, 2 = PRMT magic + FADD, 3 = PRMT magic + FADD2, 4 = none (bytes reinterpreted)
template <int CVT>
__device__ __forceinline__ void conv8(uint32_t a, uint32_t b, float2 f[4]) {
    if (CVT == 0) {
        f[0] = make_float2((float)(a & 0xff), (float)((a >> 8) & 0xff));
        f[1] = make_float2((float)((a >> 16) & 0xff), (float)(a >> 24));
        f[2] = make_float2((float)(b & 0xff), (float)((b >> 8) & 0xff));
        f[3] = make_float2((float)((b >> 16) & 0xff), (float)(b >> 24));
    } else if (CVT == 1) {
        uint32_t x[8];
#pragma unroll
        for (int k = 0; k < 4; k++) { x[k] = __byte_perm(a, 0, 0x4440 + k); x[4 + k] = __byte_perm(b, 0, 0x4440 + k); }
        float y[8];
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(y[k]) : "r"(x[k]));
        f[0] = make_float2(y[0], y[1]); f[1] = make_float2(y[2], y[3]); f[2] = make_float2(y[4], y[5]); f[3] = make_float2(y[6], y[7]);
    } else if (CVT == 2) {
        float y[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            y[k] = __uint_as_float(__byte_perm(a, 0x4b000000u, 0x7440 + k)) - 8388608.0f;
            y[4 + k] = __uint_as_float(__byte_perm(b, 0x4b000000u, 0x7440 + k)) - 8388608.0f;
        }
        f[0] = make_float2(y[0], y[1]); f[1] = make_float2(y[2], y[3]); f[2] = make_float2(y[4], y[5]); f[3] = make_float2(y[6], y[7]);
    } else if (CVT == 3) {
        const float2 m = make_float2(-8388608.0f, -8388608.0f);
        float2 t[4];
        t[0] = make_float2(__uint_as_float(__byte_perm(a, 0x4b000000u, 0x7440)), __uint_as_float(__byte_perm(a, 0x4b000000u, 0x7441)));
        t[1] = make_float2(__uint_as_float(__byte_perm(a, 0x4b000000u, 0x7442)), __uint_as_float(__byte_perm(a, 0x4b000000u, 0x7443)));
        t[2] = make_float2(__uint_as_float(__byte_perm(b, 0x4b000000u, 0x7440)), __uint_as_float(__byte_perm(b, 0x4b000000u, 0x7441)));
        t[3] = make_float2(__uint_as_float(__byte_perm(b, 0x4b000000u, 0x7442)), __uint_as_float(__byte_perm(b, 0x4b000000u, 0x7443)));
#pragma unroll
        for (int k = 0; k < 4; k++) f[k] = fadd2(t[k], m);
    } else {
        f[0] = make_float2(__uint_as_float(a), __uint_as_float(b)); f[1] = f[0]; f[2] = f[0]; f[3] = f[0];
    }
}

template <int CVT, int NSLOT>
__global__ void bench(float *out, const uint32_t *in, long long *cycles) {
    float2 acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[j][e] = make_float2(threadIdx.x * 0.001f + j, e * 0.5f);
    uint32_t a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    float w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) w[j] = 0.01f * (j + 1) + in[40] * 1e-9f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        float2 f[4];
        conv8<CVT>(a, b, f);
#pragma unroll
        for (int j = 0; j < NSLOT; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) ffma2(acc[j][e], f[e], w[j]);
        if (NSLOT == 0) { acc[0][0].x += f[0].x + f[1].x + f[2].x + f[3].x; acc[0][1].x += f[0].y + f[1].y + f[2].y + f[3].y; }
        a += 0x01010101u; b += 0x02020202u;
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) s += acc[j][e].x + acc[j][e].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CVT, int NSLOT>
void run(const char *name, int threads, float *out, uint32_t *in, long long *cyc) {
    for (int r = 0; r < 2; r++) { bench<CVT, NSLOT><<<148, threads>>>(out, in, cyc); cudaDeviceSynchronize(); }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += h[i];
    avg /= 148;
    double fma = 8.0 * NSLOT;
    printf("%-34s slots=%d threads=%4d  clk per warp-row per SMSP=%7.2f  bytes/clk/SM=%7.2f  FMA/clk/SM=%7.2f (FMA-only bound %6.2f clk)\n", name, NSLOT, threads,
           avg / ITERS / (threads / 128.0), 8.0 * ITERS * threads / avg, fma * ITERS * threads / avg, fma * 32 / 32.0);
}

int main() {
    float *out; uint32_t *in; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 64 * 4); cudaMalloc(&cyc, 148 * 8);
    cudaMemset(in, 0, 64 * 4);
    for (int threads : {256, 512}) {
        run<0, 0>("I2F.U8 only", threads, out, in, cyc);
        run<1, 0>("PRMT + I2FP.U32 only", threads, out, in, cyc);
        run<2, 0>("PRMT magic + FADD only", threads, out, in, cyc);
        run<3, 0>("PRMT magic + FADD2 only", threads, out, in, cyc);
        run<4, 6>("no conversion", threads, out, in, cyc);
        run<0, 6>("I2F.U8", threads, out, in, cyc);
        run<1, 6>("PRMT + I2FP.U32", threads, out, in, cyc);
        run<2, 6>("PRMT magic + FADD", threads, out, in, cyc);
        run<3, 6>("PRMT magic + FADD2", threads, out, in, cyc);
        run<1, 7>("PRMT + I2FP.U32", threads, out, in, cyc);
        run<3, 7>("PRMT magic + FADD2", threads, out, in, cyc);
        run<1, 8>("PRMT + I2FP.U32", threads, out, in, cyc);
        run<3, 8>("PRMT magic + FADD2", threads, out, in, cyc);
    }
    return 0;
}
