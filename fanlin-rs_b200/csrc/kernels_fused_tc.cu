// Fused separable Lanczos3 resample with the vertical pass on the sm_100a tensor
// cores (tcgen05.mma kind::i8, accumulators in TMEM).  See fused_tc.h.
//
// Per CTA (1 CTA / SM): 10 consumer warps + an MMA-issuing warp + a TMA-issuing warp; one band of <= 192 output rows
// of one image, swept left to right in chunks of 128 source bytes per row.  Per chunk, per
// group of 32 output rows:
//   * the TMA thread fetches the group's source rows with a single TMA tensor copy
//     (cp.async.bulk.tensor.2d; tensor map {row bytes, rows}, box {128 B, kg rows}, 128-byte
//     swizzle) -- the box lands as [row][128 B] with the hardware swizzle, which is the
//     MN-major SWIZZLE_128B operand layout the tensor core reads directly (measured:
//     profiles/microbench/umma_i8_tma128.cu); rows / columns past the image are zero-filled
//     -- and the s8 weight-digit tile with one cp.async.bulk (two shared-memory slots);
//   * the MMA thread issues kg/32 tcgen05.mma (M = 128 bytes of the row, N = 96 = 3
//     digits x 32 output rows, K = 32 source rows) into one of FIVE accumulator regions of
//     TMEM and commits to an mbarrier.  It runs ahead of the consumers as far as TMEM
//     allows: while they execute the horizontal stage of chunk c, the tensor core already
//     computes (almost all of) the vertical pass of chunk c + 1 -- TMEM is the double buffer;
//   * the consumers drain a region (tcgen05.ld 32x32b), recombine the three s32 digit sums
//     into the f32 value of the crate's vertical pass, store it to the tile
//     tmp[element][row] and hand the region back;
// then the consumers run the horizontal stage on the CUDA cores: the scatter of
// kernels_fused.cu (<= 8 live output pixels per row), with one lane per (pair of output rows,
// channel) so that every consumer warp takes part, f32x2 FMAs over the row pair, and a warp
// shuffle that gathers a finished pixel's channels for the epilogue.
#include <cuda.h>

#include "fused_device.cuh"
#include "fused_tc.h"
#include "kernels.h"

namespace fanlin {

namespace {

constexpr int NT = 32 * TC_H_WARPS;  // consumer threads (10 warps: 8 drain TMEM, all run the horizontal stage)
constexpr int NT_ALL = NT + 64;   // + the MMA-issuing warp and the TMA-issuing warp
constexpr uint32_t NR = 5;        // TMEM accumulator regions of 96 columns
constexpr int S = FUSED_SLOTS;
constexpr uint32_t TMEM_COLS = 512;  // NR regions of 96 columns

// Shared-memory matrix descriptor, no swizzle.  Measured on B200
// (profiles/microbench/umma_i8.cu): LBO = byte stride between core matrices along K,
// SBO = along M/N, for both the MN-major A tile and the K-major B tile.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 0) {
    uint64_t d = uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(layout) << 61);  // layout 0 = no swizzle, 2 = 128-byte swizzle
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

__device__ __forceinline__ void ffma2(float2 &acc, float2 a, float w) {
    const float2 b = make_float2(w, w);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(reinterpret_cast<unsigned long long &>(acc))
        : "l"(reinterpret_cast<const unsigned long long &>(a)), "l"(reinterpret_cast<const unsigned long long &>(b)));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

template <int C>
__global__ void __launch_bounds__(NT_ALL, 1) fused_resample_tc_kernel(const FusedTcItem *__restrict__ items,
                                                                  const CUtensorMap *__restrict__ tmaps,
                                                                  const uint8_t *__restrict__ tb,
                                                                  const float *__restrict__ tw,
                                                                  const uint32_t *__restrict__ tinfo) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms need 1024-byte alignment
    __shared__ FusedTcItem it_s;
    __shared__ __align__(8) uint64_t mbar[NR];       // the MMAs into TMEM region r have retired
    __shared__ __align__(8) uint64_t tmem_free[NR];  // the 8 draining warps have emptied region r
    __shared__ __align__(8) uint64_t full[2];  // the copies into shared-memory buffer b have landed
    __shared__ uint32_t tmem_base_s;
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        it_s = items[blockIdx.x];
        for (uint32_t r = 0; r < NR; r++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[r])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&tmem_free[r])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

    if (warp < NT / 32) fill_bars(it, warp, lane, NT / 32);

    // ---- shared-memory carve-up
    const uint32_t r_pad = it.r_pad, kg_max = it.kg_max, grp_rows = it.grp_rows;
    float *tmp = reinterpret_cast<float *>(smem);                      // [128][r_pad]
    uint8_t *sA = smem + size_t(TC_M) * r_pad * 4;                     // 2 x [kg_max rows][128 B], 128-byte swizzle (1024-aligned)
    uint8_t *sB = sA + 2 * size_t(kg_max) * TC_M;                      // 2 x [96][kg_max], core-matrix layout
    float *hw_s0 = reinterpret_cast<float *>(sB + 2 * size_t(TC_N) * kg_max);  // 2 x ([chunk_px][8] weights + [chunk_px] info)
    const uint32_t chunk_px = it.chunk_px, n_px = it.n_px, n_chunks = it.n_chunks, n_groups = it.n_groups;
    const uint32_t htab_words = (chunk_px * (S + 1) + 3) & ~3u;  // each copy stays 16-byte aligned
    const CUtensorMap *tmap = tmaps + blockIdx.x;
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    // group table {k0, kg, b_off, rows} in shared memory: the producer's issue path must not wait on L2
    __shared__ uint32_t grp[4 * 24];
    for (uint32_t k = tid; k < 4 * it.n_groups; k += NT_ALL) grp[k] = tinfo[it.grp_off + k];
    __syncthreads();
    const float scale = it.scale, scale_hi = it.scale * 16384.0f;

    // ---- horizontal-stage role of this thread: one channel of TWO output rows (ra and rb = ra +
    // h_half).  Lanes are (row slot, channel) with the channel fastest, 32 / C slots per warp, so
    // all consumer warps share the stage; the two rows go through one f32x2 FMA per output slot.
    constexpr uint32_t RPW = 32 / C;  // row slots per warp (C = 3 leaves lanes 30 and 31 idle)
    const uint32_t h_half = (it.band_rows + 1) / 2;
    const uint32_t h_ch = lane % C, h_slot = warp * RPW + lane / C;
    const bool h_warp = warp * RPW < h_half;  // warp-uniform: this warp has rows to produce
    const bool h_lane = lane < RPW * C && h_slot < h_half;
    const uint32_t h_ra = min(h_slot, h_half - 1), h_rb = min(h_ra + h_half, it.band_rows - 1);  // clamped: idle lanes read valid tile rows
    // which of the two rows this lane writes out once the channels are gathered: channel 0 -> ra,
    // channel 1 -> rb (a single-channel lane writes both)
    const bool h_emit_a = h_lane && h_ch == 0;
    const bool h_emit_b = h_lane && h_ch == (C > 1 ? 1u : 0u) && h_slot + h_half < it.band_rows;
    float2 hacc[S];  // .x = row ra, .y = row rb
#pragma unroll
    for (int j = 0; j < S; j++) hacc[j] = make_float2(0.f, 0.f);
    uint32_t h_next = 0;
    const float *hw = tw + it.hw_off;
    const uint32_t *hinfo = tinfo + it.hinfo_off;
    const uint32_t h_cx0 = it.dst_x, h_cy = it.dst_y + it.band_r0 + h_ra;

    // Thread 0 only: fetch group g of the chunk whose 16-byte aligned first column is seg0 into
    // shared-memory buffer `buf` (one TMA tensor copy + one bulk copy, completion on full[buf]).
    auto issue_load = [&](uint32_t g, uint32_t buf, uint32_t seg0) {
        const uint32_t k0 = grp[4 * g], kg = grp[4 * g + 1], b_off = grp[4 * g + 2];
        const uint32_t bar = smem_u32(&full[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kg_max * TC_M + kg * TC_N) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         sA_u + buf * kg_max * TC_M),
                     "l"(tmap), "r"(bar), "r"(seg0 * 16), "r"(k0)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sB_u + buf * TC_N * kg_max),
                     "l"(tb + b_off), "r"(kg * TC_N), "r"(bar)
                     : "memory");
    };
    // Thread 0 only: the MMAs of group g (operands in shared-memory buffer buf, accumulators in
    // TMEM buffer buf), committed to mbar[buf].
    auto issue_mma = [&](uint32_t g, uint32_t buf, uint32_t region) {
        const uint32_t kg = grp[4 * g + 1];
        // descriptors advance by a constant per K step: 32 rows x 128 B of A, 2 core matrices of B
        uint64_t da = umma_desc(sA_u + buf * kg_max * TC_M, 16, 1024, 2);  // SBO = 8-row atom stride
        uint64_t db = umma_desc(sB_u + buf * TC_N * kg_max, 128, (kg / 16) * 128);
        const uint32_t d_tmem = tmem_base + region * TC_N;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                     "l"(da), "l"(db), "r"(UMMA_IDESC)
                     : "memory");
        for (uint32_t ks = 1; ks < kg / 32; ks++) {
            da += (32 * 128) >> 4;
            db += (2 * 128) >> 4;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                         "l"(da), "l"(db), "r"(UMMA_IDESC)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[region])) : "memory");
    };

    auto chunk_seg0 = [&](uint32_t chunk) { return ((it.px0 + chunk * chunk_px) * C) >> 4; };
    // Groups are numbered flat across chunks (gg = chunk * n_groups + g): shared-memory slot
    // gg & 1, TMEM region gg % NR.
    const uint32_t total = n_chunks * n_groups;

    if (warp == NT / 32 + 1) {
        // ================= TMA warp: one thread keeps the two shared-memory slots filled =================
        if (lane == 0) {
            uint32_t ld_chunk = 0, ld_g = 0;
            for (uint32_t gg = 0; gg < total; gg++) {
                if (gg >= 2) mbar_wait(smem_u32(&mbar[(gg - 2) % NR]), ((gg - 2) / NR) & 1);  // slot gg & 1 was read by the MMAs of group gg - 2
                issue_load(ld_g, gg & 1, chunk_seg0(ld_chunk));
                if (++ld_g == n_groups) { ld_g = 0; ld_chunk++; }
            }
        }
    } else if (warp == NT / 32) {
        // ================= MMA warp: one thread issues every tcgen05.mma =================
        if (lane == 0) {
            for (uint32_t pg = 0; pg < total; pg++) {
                const uint32_t region = pg % NR, use = pg / NR;
                mbar_wait(smem_u32(&full[pg & 1]), (pg >> 1) & 1);                       // operands have landed
                if (use > 0) mbar_wait(smem_u32(&tmem_free[region]), (use - 1) & 1);    // consumers drained the region's previous contents
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                issue_mma(pg % n_groups, pg & 1, region);
            }
        }
    } else {
    // ================= consumer warps =================
    auto stage_htab = [&](uint32_t chunk) {  // the chunk's slice of the horizontal table -> its shared-memory copy
        const uint32_t cpx0 = chunk * chunk_px, npx = min(chunk_px, n_px - cpx0);
        float *dstw = hw_s0 + (chunk & 1) * htab_words;
        const uint32_t sa_w = smem_u32(dstw), sa_i = smem_u32(dstw + size_t(chunk_px) * S);
        const uint32_t *gw = reinterpret_cast<const uint32_t *>(hw + size_t(cpx0) * S);
        for (uint32_t k = tid; k < npx * S; k += NT) cp_async4(sa_w + 4 * k, gw + k, true);
        for (uint32_t k = tid; k < npx; k += NT) cp_async4(sa_i + 4 * k, hinfo + cpx0 + k, true);
        cp_async_commit();
    };
    stage_htab(0);

    uint32_t gg = 0;
    for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
        const uint32_t cpx0 = chunk * chunk_px;
        const uint32_t npx = min(chunk_px, n_px - cpx0);
        const uint32_t sh = ((it.px0 + cpx0) * C) & 15u;  // padding columns in front of the chunk
        // ================= vertical stage: drain the tensor-core results =================
        for (uint32_t g = 0; g < n_groups; g++, gg++) {
            const uint32_t region = gg % NR;
            // TMEM -> f32 tile.  Warp w < 8 reads lanes [32 (w & 3), +32) (= tile columns m) and the
            // half (w >> 2) of the group's 32 output rows; warps 8 and 9 only join the horizontal stage.
            if (warp < 8) {
                mbar_wait(smem_u32(&mbar[region]), (gg / NR) & 1);  // the MMAs of group gg have retired
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t half = warp >> 2, m = (warp & 3) * 32 + lane;
                const uint32_t taddr = tmem_base + region * TC_N + (((warp & 3) * 32u) << 16) + half * 16;
                uint32_t hi[16], mid[16], lo[16];
                tmem_ld16(taddr, hi);
                tmem_ld16(taddr + 32, mid);
                tmem_ld16(taddr + 64, lo);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_free[region])) : "memory");
                const uint32_t jn = grp[4 * g + 3];  // output rows in this group
                float *t = tmp + size_t(m) * r_pad + g * grp_rows;
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const uint32_t j = half * 16 + e;
                    const int ml = int(mid[e]) * 128 + int(lo[e]);
                    const float v = fmaf(float(int(hi[e])), scale_hi, float(ml) * scale);
                    if (j < jn) t[j] = v;  // lanes = consecutive columns, r_pad odd: conflict-free
                }
            }
        }
        if (chunk + 1 < n_chunks) stage_htab(chunk + 1);  // lands during this chunk's horizontal stage
        else cp_async_commit();
        cp_async_wait<1>();  // this chunk's table slice (committed one chunk ago) has landed
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // consumers only: the tile is complete
        // ================= horizontal stage: CUDA cores =================
        if (h_warp) {
            const float *tcol = tmp + size_t(sh + h_ch) * r_pad;
            const float *hw_s = hw_s0 + (chunk & 1) * htab_words;
            const uint32_t *hinfo_s = reinterpret_cast<const uint32_t *>(hw_s + size_t(chunk_px) * S);
            // one pixel: scatter this lane's channel value (of both rows) into the live slots, then flush completed outputs
            auto step = [&](const float2 v, const float4 &w0, const float4 &w1, uint32_t info) {
                const float w[S] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int j = 0; j < S; j++) ffma2(hacc[j], v, w[j]);
                const uint32_t fl = (info >> 8) & 0xffu;
                if (fl) {  // uniform over the CTA
#pragma unroll
                    for (int j = 0; j < S; j++) {
                        if (fl & (1u << j)) {
                            const uint32_t o = h_next + ((uint32_t(j) - h_next) & (S - 1));
                            const uint32_t mine = round_u8(hacc[j].x) | round_u8(hacc[j].y) << 8;
                            hacc[j] = make_float2(0.f, 0.f);
                            // gather the pixel's channels from the C lanes of this row slot
                            uint32_t u[4] = {0, 0, 0, 0};
                            const uint32_t sel = (C > 1 && h_ch == 1) ? 8u : 0u;
#pragma unroll
                            for (int k = 0; k < C; k++) u[k] = (__shfl_sync(0xffffffffu, mine, int(lane - h_ch) + k) >> sel) & 0xffu;
                            if constexpr (C == 1) {
                                if (h_emit_a) emit_px<C, FusedTcItem>(it, h_cx0 + o, h_cy, u);
                                u[0] = mine >> 8;
                                if (h_emit_b) emit_px<C, FusedTcItem>(it, h_cx0 + o, h_cy + h_half, u);
                            } else {
                                if (h_emit_a || h_emit_b) emit_px<C, FusedTcItem>(it, h_cx0 + o, h_cy + (h_ch ? h_half : 0u), u);
                            }
                        }
                    }
                    h_next += __popc(fl);
                }
            };
            // two pixels per iteration: all shared-memory reads of both are issued before the FMAs
            for (uint32_t xl = 0; xl < npx; xl += 2) {
                const bool two = xl + 1 < npx;
                const float *pa = tcol + size_t(xl * C) * r_pad, *pb = pa + size_t(C) * r_pad;  // within shared memory even past npx
                const float2 va = make_float2(pa[h_ra], pa[h_rb]);
                const float2 vb = make_float2(pb[h_ra], pb[h_rb]);
                const uint32_t ia = hinfo_s[xl], ib = hinfo_s[xl + 1];
                const float4 wa0 = *reinterpret_cast<const float4 *>(hw_s + size_t(xl) * S);
                const float4 wa1 = *reinterpret_cast<const float4 *>(hw_s + size_t(xl) * S + 4);
                const float4 wb0 = *reinterpret_cast<const float4 *>(hw_s + size_t(xl + 1) * S);
                const float4 wb1 = *reinterpret_cast<const float4 *>(hw_s + size_t(xl + 1) * S + 4);
                step(va, wa0, wa1, ia);
                if (two) step(vb, wb0, wb1, ib);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // the tile may be overwritten
    }
    }  // consumer warps
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

template <int C>
void launch_tc_variant(const FusedTcItem *d_items, const void *d_tmaps, uint32_t n_items, size_t smem, const uint8_t *d_b,
                       const float *d_w, const uint32_t *d_info, LaunchCtx &lc) {
    auto kern = fused_resample_tc_kernel<C>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    lc.begin("fused_resample_tc_kernel");
    kern<<<n_items, NT_ALL, smem, lc.st>>>(d_items, static_cast<const CUtensorMap *>(d_tmaps), d_b, d_w, d_info);
    lc.end();
}

}  // namespace

int launch_fused_tc(const FusedTcItem *d_items, const void *d_tmaps, uint32_t n_items, uint32_t c, size_t smem, const uint8_t *d_b,
                    const float *d_w, const uint32_t *d_info, LaunchCtx &lc) {
    if (n_items == 0) return 0;
    switch (c) {
    case 1: launch_tc_variant<1>(d_items, d_tmaps, n_items, smem, d_b, d_w, d_info, lc); return 1;
    case 2: launch_tc_variant<2>(d_items, d_tmaps, n_items, smem, d_b, d_w, d_info, lc); return 1;
    case 3: launch_tc_variant<3>(d_items, d_tmaps, n_items, smem, d_b, d_w, d_info, lc); return 1;
    case 4: launch_tc_variant<4>(d_items, d_tmaps, n_items, smem, d_b, d_w, d_info, lc); return 1;
    }
    return -1;
}

}  // namespace fanlin

// ---- host: tensor maps ---------------------------------------------------------------------
#include <cudaTypedefs.h>

namespace fanlin {

bool encode_row_tile_map(void *out, const void *base, uint32_t pitch, uint32_t rows, uint32_t box_rows) {
    static PFN_cuTensorMapEncodeTiled enc = nullptr;
    if (!enc) {
        cudaDriverEntryPointQueryResult qr;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return false;
        enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    const cuuint64_t gdim[2] = {pitch, rows};
    const cuuint64_t gstr[1] = {pitch};
    const cuuint32_t box[2] = {TC_M, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(static_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace fanlin
