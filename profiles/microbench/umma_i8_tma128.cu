// (SW128 variant: 2-D tensor map {row bytes, rows}, box {128 B, K rows}, SWIZZLE_128B, so the TMA
// unit moves 128-byte lines instead of 16-byte pieces; the A descriptor uses layout type 2.)
// Bring-up test, TMA variant: the A tile is fetched by ONE cp.async.bulk.tensor.3d from an
// image-like global array ([rows][pitch] u8) through a tensor map with dims {16 B, rows,
// 16-byte segments} (strides {pitch, 16}), which lands it as [segment][row][16 B] -- a valid
// no-swizzle core-matrix layout (K-direction stride 128 B, M-direction stride rows*16 B).
// Rows past the image are zero-filled by the TMA unit.  B comes by one cp.async.bulk.
// Original header: tcgen05.mma kind::i8 on sm_100a with hand-built shared-memory
// descriptors (no swizzle): D[128 x N] (s32, TMEM) = A[128 x K] (u8, MN-major: the natural
// row-major image tile, K = image rows) * B[K x N] (s8, K-major).  Tries the candidate
// meanings of the descriptor's LBO/SBO fields and reports mismatches against the CPU.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

constexpr int M = 128, N = 96, K = 96;  // K multiple of 32
constexpr int PITCH = 5760, IMG_ROWS = 90, ROW0 = 3, COL0 = 208;  // tile = rows [3, 99) (last 9 past the image), bytes [208, 336)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 0) {
    uint64_t d = (uint64_t)layout << 61;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tmapA, const CUtensorMap *tmap_g, const int8_t *gB /*core layout*/, int *gD /*[M][N]*/,
                                         int variant, uint32_t idesc) {
    __shared__ __align__(1024) uint8_t sA[K * M];  // [k][128 B], 16-byte chunks XOR-swizzled by k % 8
    __shared__ __align__(128) uint8_t sB[N * K];  // [n/8][k/16][n%8][16]
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t full;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = K * M + N * K;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full)), "r"(bytes) : "memory");
        const CUtensorMap *tm = variant >= 4 ? tmap_g : &tmapA;  // variant >= 4: tensor map read from global memory
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sA)), "l"(tm), "r"(smem_u32(&full)), "r"(COL0), "r"(ROW0) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sB)), "l"(gB), "r"(uint32_t(N * K)), "r"(smem_u32(&full)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&full)), "r"(0) : "memory");
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
        const uint32_t a_mn = K * 16, a_k = 128, b_k = 128, b_n = (K / 16) * 128;
        for (int ks = 0; ks < K / 32; ks++) {
            const uint32_t a_addr = smem_u32(sA) + ks * 32 * 128;  // 32 k-rows of 128 B
            const uint32_t b_addr = smem_u32(sB) + ks * 2 * b_k;  // 32 k = 2 groups of 16
            uint64_t da, db;
            const int v = variant & 3;
            da = v == 0 ? make_desc(a_addr, 1024, 128, 2) : v == 1 ? make_desc(a_addr, 128, 1024, 2) : v == 2 ? make_desc(a_addr, 1024, 1024, 2) : make_desc(a_addr, 16, 1024, 2);
            db = make_desc(b_addr, b_k, b_n);
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
        for (int j = 0; j < 16; j++) gD[tid * N + c0 + j] = (int)r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64));
}

int main() {
    std::vector<uint8_t> img(size_t(IMG_ROWS) * PITCH);
    srand(1);
    for (auto &v : img) v = rand() & 255;
    std::vector<int8_t> hB(N * K), hBc(N * K);
    for (auto &v : hB) v = (rand() % 255) - 127;
    for (int n = 0; n < N; n++)
        for (int kk = 0; kk < K; kk++) hBc[((n / 8) * (K / 16) + kk / 16) * 128 + (n % 8) * 16 + kk % 16] = hB[n * K + kk];
    std::vector<int> ref(M * N, 0);
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            int s = 0;
            for (int kk = 0; kk < K; kk++) {
                const int row = ROW0 + kk;
                const int a = row < IMG_ROWS ? img[size_t(row) * PITCH + COL0 + m] : 0;  // TMA zero-fills rows past the image
                s += a * (int)hB[n * K + kk];
            }
            ref[m * N + n] = s;
        }
    uint8_t *dImg; int8_t *dB; int *dD; CUtensorMap *dMap;
    cudaMalloc(&dImg, img.size()); cudaMalloc(&dB, hBc.size()); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dMap, sizeof(CUtensorMap));
    cudaMemcpy(dImg, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hBc.data(), hBc.size(), cudaMemcpyHostToDevice);
    PFN_cuTensorMapEncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr);
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {PITCH, IMG_ROWS};
    const cuuint64_t gstr[1] = {PITCH};
    const cuuint32_t box[2] = {128, K};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dImg, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled -> %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    cudaMemcpy(dMap, &tmap, sizeof(tmap), cudaMemcpyHostToDevice);
    for (int variant : {0, 1, 2, 3, 4}) {
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        cudaMemset(dD, 0xff, M * N * 4);
        k<<<1, 128>>>(tmap, dMap, dB, dD, variant, idesc);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<int> out(M * N);
        cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < M * N; i++) bad += out[i] != ref[i];
        printf("variant %d, tensor map in %s: %s, mismatches %d / %d  (D[0][0]=%d ref %d, D[5][3]=%d ref %d)\n", variant, variant >= 4 ? "global memory" : "kernel param",
               cudaGetErrorString(e), bad, M * N, out[0], ref[0], out[5 * N + 3], ref[5 * N + 3]);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
