// Where does tcgen05.mma cta_group::1 with M = 64 put the rows of D in TMEM?  (Same operands and descriptors as
// umma_i8.cu variant 3, M = 64: the 128 lanes are dumped and matched against the CPU result.)
// Bring-up test for tcgen05.mma kind::i8 on sm_100a with hand-built shared-memory
// descriptors (no swizzle): D[128 x N] (s32, TMEM) = A[128 x K] (u8, MN-major: the natural
// row-major image tile, K = image rows) * B[K x N] (s8, K-major).  Tries the candidate
// meanings of the descriptor's LBO/SBO fields and reports mismatches against the CPU.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 64, N = 48, K = 96, LANES = 128;  // K multiple of 32

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128) k(const uint8_t *gA /*[K][M]*/, const int8_t *gB /*[N][K]*/, int *gD /*[M][N]*/,
                                         int variant, uint32_t idesc) {
    __shared__ __align__(128) uint8_t sA[K * M];  // [k/8][m/16][k%8][16]
    __shared__ __align__(128) uint8_t sB[N * K];  // [n/8][k/16][n%8][16]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < K * (M / 16); i += 128) {  // one 16-byte segment each
        const int kk = i / (M / 16), j = i % (M / 16);
        const uint4 v = *reinterpret_cast<const uint4 *>(gA + kk * M + j * 16);
        *reinterpret_cast<uint4 *>(sA + ((kk / 8) * (M / 16) + j) * 128 + (kk % 8) * 16) = v;
    }
    for (int i = tid; i < N * (K / 16); i += 128) {
        const int n = i / (K / 16), kc = i % (K / 16);
        const uint4 v = *reinterpret_cast<const uint4 *>(gB + n * K + kc * 16);
        *reinterpret_cast<uint4 *>(sB + ((n / 8) * (K / 16) + kc) * 128 + (n % 8) * 16) = v;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        // strides between core matrices: A along MN = 128 B, along K = (M/16)*128; B along K = 128 B, along N = (K/16)*128
        const uint32_t a_mn = 128, a_k = (M / 16) * 128, b_k = 128, b_n = (K / 16) * 128;
        for (int ks = 0; ks < K / 32; ks++) {
            const uint32_t a_addr = smem_u32(sA) + ks * 4 * a_k;  // 32 k-rows = 4 groups of 8
            const uint32_t b_addr = smem_u32(sB) + ks * 2 * b_k;  // 32 k = 2 groups of 16
            uint64_t da, db;
            if (variant == 0) { da = make_desc(a_addr, a_mn, a_k); db = make_desc(b_addr, b_k, b_n); }
            else if (variant == 1) { da = make_desc(a_addr, a_k, a_mn); db = make_desc(b_addr, b_n, b_k); }
            else if (variant == 2) { da = make_desc(a_addr, a_mn, a_k); db = make_desc(b_addr, b_n, b_k); }
            else { da = make_desc(a_addr, a_k, a_mn); db = make_desc(b_addr, b_k, b_n); }
            const uint32_t acc = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reads TMEM lanes [32w, 32w+32): thread = row m, 16 columns at a time
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; j++) gD[tid * N + c0 + j] = (int)r[j];  // tid = TMEM lane
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64));
}

int main() {
    std::vector<uint8_t> hA(K * M);
    std::vector<int8_t> hB(N * K);
    srand(1);
    for (auto &v : hA) v = rand() & 255;
    for (auto &v : hB) v = (rand() % 255) - 127;
    std::vector<int> ref(M * N, 0);
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            int s = 0;
            for (int kk = 0; kk < K; kk++) s += (int)hA[kk * M + m] * (int)hB[n * K + kk];
            ref[m * N + n] = s;
        }
    uint8_t *dA; int8_t *dB; int *dD;
    cudaMalloc(&dA, hA.size()); cudaMalloc(&dB, hB.size()); cudaMalloc(&dD, LANES * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    cudaMemset(dD, 0x7f, LANES * N * 4);
    k<<<1, 128>>>(dA, dB, dD, 3, idesc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("M = 64, N = %d: %s\n", N, cudaGetErrorString(e));
    std::vector<int> out(LANES * N);
    cudaMemcpy(out.data(), dD, LANES * N * 4, cudaMemcpyDeviceToHost);
    for (int lane = 0; lane < LANES; lane++) {
        int row = -1, col_shift = 0;
        for (int m = 0; m < M && row < 0; m++)
            for (int sh = 0; sh <= 0; sh++) {
                bool all = true;
                for (int n = 0; n < N; n++) all = all && out[lane * N + n] == ref[m * N + n];
                if (all) { row = m; col_shift = sh; }
            }
        if (lane % 16 == 0) printf("\nlanes %3d..%3d hold rows:", lane, lane + 15);
        printf(" %3d", row);
    }
    printf("\n(-1 = the lane holds none of the rows in columns [0, N))\n");
    return 0;
}
