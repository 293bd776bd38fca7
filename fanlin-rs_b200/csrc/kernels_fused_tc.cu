// Fused separable Lanczos3 resample with the vertical pass on the sm_100a tensor
// cores (tcgen05.mma kind::i8, accumulators in TMEM).  See fused_tc.h.
//
// Per CTA (8 warps, 1 CTA / SM): one band of <= 192 output rows of one image, swept
// left to right in chunks of 128 source bytes per row.  Per chunk, per group of 32
// output rows:
//   * all threads cp.async the group's source rows (<= 256 x 128 B, straight from the
//     image, placed in the no-swizzle core-matrix layout) and its s8 weight-digit tile
//     into one of two shared-memory buffers -- the next group's copies are in flight
//     while this group is computed;
//   * one thread issues kg/32 tcgen05.mma (M = 128 bytes of the row, N = 96 = 3 digits
//     x 32 output rows, K = 32 source rows each) and commits to an mbarrier;
//   * all 8 warps read their quarter of TMEM (tcgen05.ld 32x32b), recombine the three
//     s32 digit sums into the f32 value of the crate's vertical pass and store it to the
//     tile tmp[element][row];
// then the horizontal stage runs on the CUDA cores exactly as in kernels_fused.cu (one
// thread per output row, scatter into <= 8 live output pixels, epilogue).
#include "fused_device.cuh"
#include "fused_tc.h"
#include "kernels.h"

namespace fanlin {

namespace {

constexpr int NT = 256;
constexpr int S = FUSED_SLOTS;
constexpr uint32_t TMEM_COLS = 256;  // two accumulator buffers of 96 columns at 0 and 128

// Shared-memory matrix descriptor, no swizzle.  Measured on B200
// (profiles/microbench/umma_i8.cu): LBO = byte stride between core matrices along K,
// SBO = along M/N, for both the MN-major A tile and the K-major B tile.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = uint64_t((saddr & 0x3FFFFu) >> 4);
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= uint64_t(1) << 46;  // descriptor version of sm_100
    return d;
}

// Instruction descriptor: D = S32 (2 @ bit 4), A = U8 (0 @ bit 7), B = S8 (1 @ bit 10),
// A is MN-major (bit 15), B is K-major, N >> 3 @ bit 17, M >> 4 @ bit 24.
constexpr uint32_t UMMA_IDESC = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | ((TC_N >> 3) << 17) | ((TC_M >> 4) << 24);

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

template <int C>
__global__ void __launch_bounds__(NT, 1) fused_resample_tc_kernel(const FusedTcItem *__restrict__ items,
                                                                  const uint8_t *__restrict__ tb,
                                                                  const float *__restrict__ tw,
                                                                  const uint32_t *__restrict__ tinfo) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ FusedTcItem it_s;
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        it_s = items[blockIdx.x];
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const FusedTcItem &it = it_s;
    const uint32_t tmem_base = tmem_base_s;

    fill_bars(it, warp, lane, NT / 32);

    // ---- shared-memory carve-up
    const uint32_t r_pad = it.r_pad, kg_max = it.kg_max, pitch = it.src_pitch;
    float *tmp = reinterpret_cast<float *>(smem);                      // [128][r_pad]
    uint8_t *sA = smem + size_t(TC_M) * r_pad * 4;                     // 2 x [kg_max rows][128 B], core-matrix layout
    uint8_t *sB = sA + 2 * size_t(kg_max) * TC_M;                      // 2 x [96][kg_max], core-matrix layout
    float *hw_s = reinterpret_cast<float *>(sB + 2 * size_t(TC_N) * kg_max);  // [chunk_px][8], then info words
    const uint32_t chunk_px = it.chunk_px, n_px = it.n_px, n_chunks = it.n_chunks, n_groups = it.n_groups;
    const uint32_t *hinfo_s = reinterpret_cast<const uint32_t *>(hw_s + size_t(chunk_px) * S);
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    const uint32_t *grp = tinfo + it.grp_off;
    const float scale = it.scale, scale_hi = it.scale * 16384.0f;

    // ---- horizontal-stage role of this thread: one output row of the band
    const bool h_active = tid < it.band_rows;
    float hacc[S][C];
#pragma unroll
    for (int j = 0; j < S; j++)
#pragma unroll
        for (int k = 0; k < C; k++) hacc[j][k] = 0.f;
    uint32_t h_next = 0;
    const float *hw = tw + it.hw_off;
    const uint32_t *hinfo = tinfo + it.hinfo_off;
    const uint32_t h_cx0 = it.dst_x, h_cy = it.dst_y + it.band_r0 + tid;

    // cp.async of group g of the current chunk into buffer `buf`: the source rows as 16-byte
    // pieces (a quarter warp writes the 8 rows of one core matrix: 128 contiguous bytes of shared
    // memory, and reads full 32-byte sectors), and the weight-digit tile.  `src_al` is the
    // 16-byte aligned start of the chunk in row 0; the tile's first `sh` columns are padding.
    const uint32_t ld_kr = lane & 7, ld_seg = lane >> 3;
    auto issue_load = [&](uint32_t g, uint32_t buf, const uint8_t *src_al, uint32_t avail) {
        const uint32_t k0 = grp[4 * g], kg = grp[4 * g + 1], b_off = grp[4 * g + 2];
        const uint32_t a_dst = sA_u + buf * kg_max * TC_M + ld_kr * 16;
        const uint32_t y_max = it.src_h - 1;
        for (uint32_t u = warp; u < kg / 4; u += NT / 32) {  // unit = (block of 8 rows, half of the 8 segments)
            const uint32_t rb = u >> 1, seg = (u & 1) * 4 + ld_seg;
            const uint32_t y = min(k0 + rb * 8 + ld_kr, y_max);  // rows past the image carry zero weights
            cp_async16_if(a_dst + (rb * 8 + seg) * 128, src_al + size_t(y) * pitch + seg * 16, seg * 16 < avail);
        }
        const uint32_t b_dst = sB_u + buf * TC_N * kg_max;
        const uint8_t *bsrc = tb + b_off;
        for (uint32_t i = tid; i < kg * (TC_N / 16); i += NT) cp_async16(b_dst + i * 16, bsrc + size_t(i) * 16);
    };

    uint32_t uses[2] = {0, 0};  // completed MMA batches per accumulator buffer (mbarrier phase)

    for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
        const uint32_t cpx0 = chunk * chunk_px;
        const uint32_t npx = min(chunk_px, n_px - cpx0);
        const uint32_t byte0 = (it.px0 + cpx0) * C, al0 = byte0 & ~15u, sh = byte0 - al0;
        const uint8_t *src_col = it.src + al0;
        const uint32_t nbytes = pitch - al0;  // bytes of the row available from the aligned start
        // horizontal table slice of this chunk + the first two groups
        {
            const uint32_t sa_w = smem_u32(hw_s), sa_i = smem_u32(hinfo_s);
            const uint32_t *gw = reinterpret_cast<const uint32_t *>(hw + size_t(cpx0) * S);
            for (uint32_t k = tid; k < npx * S; k += NT) cp_async4(sa_w + 4 * k, gw + k, true);
            for (uint32_t k = tid; k < npx; k += NT) cp_async4(sa_i + 4 * k, hinfo + cpx0 + k, true);
        }
        issue_load(0, 0, src_col, nbytes);
        cp_async_commit();
        if (n_groups > 1) issue_load(1, 1, src_col, nbytes);
        cp_async_commit();

        // ================= vertical stage: tensor cores =================
        for (uint32_t g = 0; g < n_groups; g++) {
            const uint32_t buf = g & 1;
            cp_async_wait<1>();  // this thread's copies of group g have landed
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // -> visible to the tensor core (async proxy)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t kg = grp[4 * g + 1];
            if (tid == 0) {
                const uint32_t a0 = sA_u + buf * kg_max * TC_M, b0 = sB_u + buf * TC_N * kg_max;
                const uint32_t d_tmem = tmem_base + buf * 128;
                for (uint32_t ks = 0; ks < kg / 32; ks++) {
                    const uint64_t da = umma_desc(a0 + ks * 4 * (TC_M / 16) * 128, (TC_M / 16) * 128, 128);
                    const uint64_t db = umma_desc(b0 + ks * 2 * 128, 128, (kg / 16) * 128);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(UMMA_IDESC),
                        "r"(uint32_t(ks > 0))
                        : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[buf]))
                             : "memory");
            }
            mbar_wait(smem_u32(&mbar[buf]), uses[buf] & 1);
            uses[buf]++;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // the MMAs of group g are done: its shared-memory buffer is free for group g + 2
            if (g + 2 < n_groups) issue_load(g + 2, buf, src_col, nbytes);
            cp_async_commit();
            // ---- epilogue: TMEM -> f32 tile.  Warp w reads lanes [32 (w & 3), +32) (= chunk
            // bytes m) and the half (w >> 2) of the group's 32 output rows.
            {
                const uint32_t half = warp >> 2, m = (warp & 3) * 32 + lane;
                const uint32_t taddr = tmem_base + buf * 128 + (((warp & 3) * 32u) << 16) + half * 16;
                uint32_t hi[16], mid[16], lo[16];
                tmem_ld16(taddr, hi);
                tmem_ld16(taddr + 32, mid);
                tmem_ld16(taddr + 64, lo);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t r0 = g * TC_GROUP_ROWS + half * 16;
                float *t = tmp + size_t(m) * r_pad + r0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int j = 4 * q + e;
                        const int ml = int(mid[j]) * 128 + int(lo[j]);
                        v[e] = fmaf(float(int(hi[j])), scale_hi, float(ml) * scale);
                    }
                    if (r0 + 4 * q < r_pad) *reinterpret_cast<float4 *>(t + 4 * q) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        cp_async_wait<0>();
        __syncthreads();
        // ================= horizontal stage: CUDA cores =================
        if (h_active) {
            const float *tcol = tmp + tid;
            for (uint32_t xl = 0; xl < npx; xl++) {
                float v[C];
#pragma unroll
                for (int k = 0; k < C; k++) v[k] = tcol[size_t(sh + xl * C + k) * r_pad];
                const uint32_t info = hinfo_s[xl];
                const float4 w0 = *reinterpret_cast<const float4 *>(hw_s + size_t(xl) * S);
                const float4 w1 = *reinterpret_cast<const float4 *>(hw_s + size_t(xl) * S + 4);
                const float w[S] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int j = 0; j < S; j++)
#pragma unroll
                    for (int k = 0; k < C; k++) hacc[j][k] = fmaf(v[k], w[j], hacc[j][k]);
                const uint32_t fl = (info >> 8) & 0xffu;
                if (fl) {
#pragma unroll
                    for (int j = 0; j < S; j++) {
                        if (fl & (1u << j)) {
                            const uint32_t o = h_next + ((uint32_t(j) - h_next) & (S - 1));
                            uint32_t u[4] = {0, 0, 0, 0};
#pragma unroll
                            for (int k = 0; k < C; k++) { u[k] = round_u8(hacc[j][k]); hacc[j][k] = 0.f; }
                            emit_px<C, FusedTcItem>(it, h_cx0 + o, h_cy, u);
                        }
                    }
                    h_next += __popc(fl);
                }
            }
        }
        __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

template <int C>
void launch_tc_variant(const FusedTcItem *d_items, uint32_t n_items, size_t smem, const uint8_t *d_b, const float *d_w,
                       const uint32_t *d_info, LaunchCtx &lc) {
    auto kern = fused_resample_tc_kernel<C>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    lc.begin("fused_resample_tc_kernel");
    kern<<<n_items, NT, smem, lc.st>>>(d_items, d_b, d_w, d_info);
    lc.end();
}

}  // namespace

int launch_fused_tc(const FusedTcItem *d_items, uint32_t n_items, uint32_t c, size_t smem, const uint8_t *d_b,
                    const float *d_w, const uint32_t *d_info, LaunchCtx &lc) {
    if (n_items == 0) return 0;
    switch (c) {
    case 1: launch_tc_variant<1>(d_items, n_items, smem, d_b, d_w, d_info, lc); return 1;
    case 2: launch_tc_variant<2>(d_items, n_items, smem, d_b, d_w, d_info, lc); return 1;
    case 3: launch_tc_variant<3>(d_items, n_items, smem, d_b, d_w, d_info, lc); return 1;
    case 4: launch_tc_variant<4>(d_items, n_items, smem, d_b, d_w, d_info, lc); return 1;
    }
    return -1;
}

}  // namespace fanlin
