// (variant: the issuing thread is chosen with elect.sync instead of tid == 0)
// Microbenchmark: issue rate of tcgen05.mma kind::i8 (cta_group::1, M = 128, K = 32 per
// instruction) on B200 for different N and for K-major vs MN-major A operands.  Data is
// garbage; only the time from first issue to the commit's mbarrier completion matters.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = (uint64_t)layout << 61;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
constexpr int NMMA = 256;
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

__global__ void __launch_bounds__(128) k(long long *out, uint32_t n, int a_mn_major, int a_swz, uint32_t m, uint32_t nacc) {
    extern __shared__ __align__(1024) uint8_t smem[];  // A: 32 KB, B: 96 KB
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32768; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((n >> 3) << 17) | ((m >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 32768;
        long long t0 = clock64();
        for (int i = 0; i < NMMA; i++) {
            const uint32_t ks = i & 7;
            uint64_t da, db;
            if (a_mn_major) da = a_swz ? make_desc(a0 + ks * 4096, 16, 1024, 2) : make_desc(a0 + ks * 512, 128, 4096, 0);
            else da = a_swz ? make_desc(a0 + ks * 32, 16, 1024, 2) : make_desc(a0 + ks * 256, 128, 1024, 0);  // K-major [128 rows][K]
            db = make_desc(b0 + ks * 256, 128, 2048, 0);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm), "l"(da),
                         "l"(db), "r"(idesc), "r"(1)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
    long long *d;
    cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 132096);
    for (uint32_t m : {128u})
            for (uint32_t n : {16u, 48u, 64u, 96u, 128u, 192u, 256u}) {
                const int mn = 1, swz = 1; const uint32_t nacc = 1;
                for (int r = 0; r < 2; r++) { k<<<148, 128, 132096>>>(d, n, mn, swz, m, nacc); cudaDeviceSynchronize(); }
                long long h[148];
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                double avg = 0;
                for (int i = 0; i < 148; i++) avg += h[i];
                avg /= 148;
                printf("elect.sync issue: i8 M=%3u N=%3u: %7.1f clk per MMA   [%s]\n", m, n, avg / NMMA, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
