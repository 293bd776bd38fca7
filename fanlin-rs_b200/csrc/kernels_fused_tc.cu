// Fused separable Lanczos3 resample on the sm_100a tensor cores (accumulators in TMEM).  See fused_tc.h.
// Three kernels share the operand pipeline below:
//   fused_resample_tc2_kernel  both passes on the tensor cores (vertical kind::i8, horizontal kind::f16 on
//                              f16 hi / lo tiles); the shipped path where the output ring fits -- second half of the file
//   fused_resample_tc_kernel   vertical pass on the tensor cores, horizontal scatter stage on the CUDA cores
//   blur_v_tc_kernel           vertical Gaussian pass of the blur
// What follows describes fused_resample_tc_kernel; the tc2 kernel has its own header further down.
//
// Per CTA (1 CTA / SM): 8 consumer warps, an MMA-issuing warp and two TMA-issuing warps; one
// band of <= 192 output rows of one image, swept left to right in chunks of 128 source bytes
// per row (a pixel cut by a chunk boundary waits: its leading bytes are carried to the front of
// the next chunk's tile).  Per chunk, per group of <= 32 output rows:
//   * the source TMA thread fetches the group's source rows with one tensor copy
//     (cp.async.bulk.tensor.2d; tensor map {row bytes, rows}, box {128 B, kg rows}, 128-byte
//     swizzle) into one of 2-4 slots -- the box lands as [row][128 B] with the hardware
//     swizzle, which is the MN-major SWIZZLE_128B operand layout the tensor core reads
//     directly (measured: profiles/microbench/umma_i8_tma128.cu); rows / columns past the
//     image are zero-filled -- and the weight TMA thread the s8 digit tile with one
//     cp.async.bulk into one of two slots;
//   * the MMA thread issues kg/32 tcgen05.mma (M = 128 bytes of the row, N = 96 = 3
//     digits x 32 output rows, K = 32 source rows) into one of FIVE accumulator regions of
//     TMEM and commits to an mbarrier, which also frees the operand slots.  It runs ahead of
//     the consumers as far as TMEM allows: while they execute the horizontal stage of chunk
//     c, the tensor core computes the vertical pass of chunk c + 1 -- TMEM is the double buffer;
//   * the consumers drain a region (tcgen05.ld 32x32b), recombine the three s32 digit sums
//     into the f32 value of the crate's vertical pass, store it to the tile
//     tmp[element][row] and hand the region back;
// then the consumers run the horizontal stage on the CUDA cores: a scatter into the <= 8
// unfinished output pixels of a row, one thread per (pair of output rows, channel), f32x2
// FMAs over the row pair, weights per pixel pair from a table staged in shared memory, the
// accumulators a shift register that is written out whenever a pixel's window ends.  Finished
// pixels are staged in shared memory and leave as whole words per row after the chunk.
#include <cuda.h>

#include "fused_device.cuh"
#include "fused_tc.h"
#include "kernels.h"
#include "tc_device.cuh"

namespace fanlin {

namespace {

constexpr int NT = 32 * TC_H_WARPS;  // consumer threads (8 warps: each drains a quarter of TMEM's lanes and runs a share of the horizontal stage)
constexpr int NT_ALL = NT + 96;   // + the MMA-issuing warp and the two TMA-issuing warps (source slabs, weight tiles)
constexpr uint32_t NB = 2;        // shared-memory slots for weight-digit tiles
constexpr uint32_t SLAB = 32 * TC_M;  // one K step of A: 32 source rows x 128 bytes
constexpr uint32_t NA_MAX = 4;    // most shared-memory slots for a group's source rows
constexpr uint32_t NR = 5;        // TMEM accumulator regions of 96 columns
constexpr int S = FUSED_SLOTS;
#ifndef TC_PF_AHEAD
#define TC_PF_AHEAD 4
#endif
constexpr uint32_t TMEM_COLS = 512;  // NR regions of 96 columns

// ---- the operand / MMA pipeline, shared by the resample kernel and the vertical-blur kernel ----
// Groups are numbered flat across chunks (gg = chunk * n_groups + g): source slot gg % n_a,
// weight-tile slot gg % NB, TMEM region gg % NR.
struct TcPipe {
    uint32_t mbar, tmem_free, a_full, b_full;  // shared-memory addresses of the mbarrier arrays (8 bytes per entry)
    uint32_t sA_u, sB_u;                        // source slots [n_a][kg_max rows][128 B]; weight slots [NB][96][kg_max]
    uint32_t kg_max, n_a, n_chunks, n_groups;
    uint32_t x0;                                // byte offset of chunk 0 in a row (16-byte aligned)
    uint32_t tmem_base;
    uint32_t pf_ahead;                          // groups the L2 prefetch runs ahead of the copies (0 = off)
    const uint32_t *grp;                        // {k0, kg, b_off, rows} x n_groups, in shared memory
};

// source TMA thread: one tensor copy per group (kg_max rows x 128 B), n_a groups deep
template <uint32_t NRT = NR>
__device__ __forceinline__ void tc_source_role(const TcPipe &p, const CUtensorMap *tmap) {
    const uint32_t total = p.n_chunks * p.n_groups;
    uint32_t g = 0, chunk = 0, slot = 0;
    // L2 prefetch cursor: the box of group gg + pf_ahead is requested into L2 when group gg's copy is issued, so that
    // the copy itself (issued when a slot frees up) finds its lines in L2
    uint32_t pg = 0, pchunk = 0;
    const uint32_t pf_ahead = p.pf_ahead;
    // The tensor map lives in global memory, written by a host copy into a recycled block: the tensormap proxy must
    // not serve a stale descriptor it cached for an earlier launch's map at the same address.
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
    for (uint32_t k = 0; k < pf_ahead && pchunk < p.n_chunks; k++)
        if (++pg == p.n_groups) { pg = 0; pchunk++; }
    for (uint32_t gg = 0; gg < total; gg++) {
        if (pf_ahead && pchunk < p.n_chunks) {
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(p.x0 + TC_M * pchunk), "r"(p.grp[4 * pg]) : "memory");
            if (++pg == p.n_groups) { pg = 0; pchunk++; }
        }
        if (gg >= p.n_a) mbar_wait_wd(p.mbar + 8 * ((gg - p.n_a) % NRT), ((gg - p.n_a) / NRT) & 1);  // the slot was read by the MMAs of group gg - n_a
        const uint32_t bar = p.a_full + 8 * slot;
#ifdef TC_EXP_SKIP_A  // TIMING EXPERIMENT ONLY (wrong results): the kernel without any source traffic (HBM, L2 -> SM, the TMA writes)
        if (gg >= 2 * p.n_groups) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
            if (++slot == p.n_a) slot = 0;
            if (++g == p.n_groups) { g = 0; chunk++; }
            continue;
        }
#endif
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(p.kg_max * TC_M) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         p.sA_u + slot * p.kg_max * TC_M),
                     "l"(tmap), "r"(bar), "r"(p.x0 + TC_M * chunk), "r"(p.grp[4 * g])
                     : "memory");
        if (++slot == p.n_a) slot = 0;
        if (++g == p.n_groups) { g = 0; chunk++; }
    }
}

// weight TMA thread: one bulk copy per group into slot gg % NB
template <uint32_t NRT = NR>
__device__ __forceinline__ void tc_weight_role(const TcPipe &p, const uint8_t *tb) {
    const uint32_t total = p.n_chunks * p.n_groups;
    uint32_t g = 0;
    for (uint32_t gg = 0; gg < total; gg++) {
        if (gg >= NB) mbar_wait_wd(p.mbar + 8 * ((gg - NB) % NRT), ((gg - NB) / NRT) & 1);  // the slot's previous tile was read by the MMAs of group gg - NB
        const uint32_t kg = p.grp[4 * g + 1], b_off = p.grp[4 * g + 2];
        const uint32_t bar = p.b_full + 8 * (gg % NB);
#ifdef TC_EXP_SKIP_B  // TIMING EXPERIMENT ONLY (wrong results): what the kernel costs without the weight tiles' L2 -> SM traffic
        if (gg >= 2 * p.n_groups) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
            if (++g == p.n_groups) g = 0;
            continue;
        }
#endif
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kg * TC_N) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         p.sB_u + (gg % NB) * TC_N * p.kg_max),
                     "l"(tb + b_off), "r"(kg * TC_N), "r"(bar)
                     : "memory");
        if (++g == p.n_groups) g = 0;
    }
}

// MMA thread: issues every tcgen05.mma; the commit of a group frees its operand slots and hands its TMEM region to the consumers
__device__ __forceinline__ void tc_mma_role(const TcPipe &p) {
    const uint32_t total = p.n_chunks * p.n_groups;
    uint32_t g = 0, slot = 0, suse = 0;
    for (uint32_t gg = 0; gg < total; gg++) {
        const uint32_t region = gg % NR, ruse = gg / NR, bslot = gg % NB;
        const uint32_t kg = p.grp[4 * g + 1];
        mbar_wait_wd(p.b_full + 8 * bslot, (gg / NB) & 1);                        // the weight tile has landed
        mbar_wait_wd(p.a_full + 8 * slot, suse & 1);                              // the source rows have landed
        if (ruse > 0) mbar_wait_wd(p.tmem_free + 8 * region, (ruse - 1) & 1);     // consumers drained the region's previous contents
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // descriptors advance by a constant per K step: 32 rows x 128 B of A, 2 core matrices of B
        uint64_t da = umma_desc(p.sA_u + slot * p.kg_max * TC_M, 16, 1024, 2);  // SBO = 8-row atom stride
        uint64_t db = umma_desc(p.sB_u + bslot * TC_N * p.kg_max, 128, (kg / 16) * 128);
        const uint32_t d_tmem = p.tmem_base + region * TC_N;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                     "l"(da), "l"(db), "r"(UMMA_IDESC)
                     : "memory");
        for (uint32_t ks = 1; ks < kg / 32; ks++) {
            da += SLAB >> 4;
            db += (2 * 128) >> 4;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                         "l"(da), "l"(db), "r"(UMMA_IDESC)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(p.mbar + 8 * region) : "memory");
        if (++slot == p.n_a) { slot = 0; suse++; }
        if (++g == p.n_groups) g = 0;
    }
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
    __shared__ __align__(8) uint64_t a_full[NA_MAX];         // the group's source rows have landed
    __shared__ __align__(8) uint64_t b_full[NB];             // the group's weight-digit tile has landed
    __shared__ uint32_t tmem_base_s;
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        it_s = items[blockIdx.x];
        for (uint32_t r = 0; r < NR; r++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[r])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&tmem_free[r])));
        }
        for (uint32_t r = 0; r < NA_MAX; r++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&a_full[r])));
        for (uint32_t r = 0; r < NB; r++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&b_full[r])));
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
    const uint32_t n_a = it.n_a;
    uint8_t *sA = smem;                                                // n_a x [kg_max rows][128 B], 128-byte swizzle (1024-aligned)
    uint8_t *sB = sA + size_t(n_a) * kg_max * TC_M;                    // NB x [96][kg_max], core-matrix layout
    float *tmp = reinterpret_cast<float *>(sB + NB * size_t(TC_N) * kg_max);  // [128 + C - 1 columns][r_pad]
    float *hw_s0 = tmp + ((size_t(TC_M + C - 1) * r_pad + 3) & ~size_t(3));  // 2 x ([max_pairs][16] weights + [max_pairs] counts), 16-byte aligned
    const uint32_t max_pairs = it.max_pairs, n_chunks = it.n_chunks, n_groups = it.n_groups;
    const uint32_t htab_words = (max_pairs * (2 * S + 1) + 3) & ~3u;  // each copy stays 16-byte aligned
    const uint32_t *crec = tinfo + it.chunk_off;  // per chunk {first pair, pairs | odd << 16, carried columns, tile column of its first pixel}
    // pixels finished during a chunk wait here, [band row][out_stride words], and leave as whole
    // words per row segment after the chunk (scattered 1-byte stores cost one LSU slot per row)
    uint32_t *out_s = reinterpret_cast<uint32_t *>(hw_s0 + 2 * htab_words);
    const uint32_t out_stride = it.out_stride;
    for (uint32_t k = tid; k < it.band_rows * out_stride; k += NT_ALL) out_s[k] = 0xffffffffu;  // alpha of opaque RGBA output stays 255
    const CUtensorMap *tmap = tmaps + blockIdx.x;
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    // group table {k0, kg, b_off, rows} in shared memory: the producer's issue path must not wait on L2
    __shared__ uint32_t grp[4 * 24];
    for (uint32_t k = tid; k < 4 * it.n_groups; k += NT_ALL) grp[k] = tinfo[it.grp_off + k];
    __syncthreads();
    const float scale = it.scale, scale_hi = it.scale * 16384.0f;

    // ---- horizontal-stage role of this thread: one channel of TWO output rows (ra and rb = ra +
    // h_half).  The (row slot, channel) pairs are laid over the consumer threads back to back,
    // channel fastest -- a slot's channels may sit in two warps -- so that 85 slots x 3 channels
    // fill exactly the 8 warps, two per scheduler; the two rows go through one f32x2 FMA per
    // output slot.
    const uint32_t h_half = (it.band_rows + 1) / 2;
    const uint32_t h_ch = tid % C, h_slot = tid / C;
    const bool h_warp = (warp * 32) / C < h_half;  // warp-uniform: this warp has rows to produce
    const bool h_lane = h_slot < h_half;
    const uint32_t h_ra = min(h_slot, h_half - 1), h_rb = min(h_ra + h_half, it.band_rows - 1);  // clamped: idle lanes read valid tile rows
    float2 hacc[S];  // .x = row ra, .y = row rb; slot j = the j-th unfinished output pixel (shift register)
#pragma unroll
    for (int j = 0; j < S; j++) hacc[j] = make_float2(0.f, 0.f);
    const float *hw = tw + it.hw_off;            // [pixel pairs][2][8]
    const uint32_t *hinfo = tinfo + it.hinfo_off;  // [pixel pairs] outputs finished by the pair
    // first canvas byte of the band's rows: row r starts at h_row0 + r * dst_pitch
    const uint32_t h_cout = it.c_out, h_pitch = it.dst_pitch, h_rows = it.band_rows;
    uint8_t *const h_row0 = it.dst + size_t(it.dst_y + it.band_r0) * it.dst_pitch + size_t(it.dst_x) * h_cout;
    const bool h_has_b = h_lane && h_slot + h_half < it.band_rows;
    const uint32_t h_pha = uint32_t(reinterpret_cast<uintptr_t>(h_row0 + size_t(h_ra) * it.dst_pitch));  // low bits: alignment phase
    const uint32_t h_phb = uint32_t(reinterpret_cast<uintptr_t>(h_row0 + size_t(h_rb) * it.dst_pitch));
    const bool h_words = h_cout == 4 && ((reinterpret_cast<uintptr_t>(h_row0) | h_pitch) & 3) == 0;
    const uint32_t out_u = smem_u32(out_s);
    const uint32_t *cpre = tinfo + it.cpre_off;
    const uint32_t h_epi = it.epi, h_fill = it.fill;

    TcPipe pipe;
    pipe.mbar = smem_u32(&mbar[0]); pipe.tmem_free = smem_u32(&tmem_free[0]); pipe.a_full = smem_u32(&a_full[0]); pipe.b_full = smem_u32(&b_full[0]);
    pipe.sA_u = sA_u; pipe.sB_u = sB_u; pipe.kg_max = kg_max; pipe.n_a = n_a; pipe.n_chunks = n_chunks; pipe.n_groups = n_groups;
    pipe.x0 = it.b0; pipe.tmem_base = tmem_base; pipe.grp = grp; pipe.pf_ahead = TC_PF_AHEAD;

    if (warp == NT / 32 + 1) {
        if (elect_one()) tc_source_role(pipe, tmap);
    } else if (warp == NT / 32 + 2) {
        if (elect_one()) tc_weight_role(pipe, tb);
    } else if (warp == NT / 32) {
        if (elect_one()) tc_mma_role(pipe);
    } else {
    // ================= consumer warps =================
    auto stage_htab = [&](uint32_t chunk) {  // the chunk's slice of the horizontal table -> its shared-memory copy
        const uint32_t pair0 = __ldg(crec + 4 * chunk), npairs = __ldg(crec + 4 * chunk + 1) & 0xffffu;
        float *dstw = hw_s0 + (chunk & 1) * htab_words;
        const uint32_t sa_w = smem_u32(dstw), sa_i = smem_u32(dstw + size_t(max_pairs) * 2 * S);
        const uint32_t *gw = reinterpret_cast<const uint32_t *>(hw + size_t(pair0) * 2 * S);
        for (uint32_t k = tid; k < npairs * 4; k += NT) cp_async16(sa_w + 16 * k, gw + 4 * k);
        for (uint32_t k = tid; k < npairs; k += NT) cp_async4(sa_i + 4 * k, hinfo + pair0 + k, true);
        cp_async_commit();
    };
    stage_htab(0);

    uint32_t gg = 0;
    for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
        const uint32_t rec_pairs = __ldg(crec + 4 * chunk + 1), carry = __ldg(crec + 4 * chunk + 2), hstart = __ldg(crec + 4 * chunk + 3);
        const uint32_t o_first = __ldg(cpre + chunk), o_count = __ldg(cpre + chunk + 1) - o_first;  // output pixels finished by this chunk
        // ================= vertical stage: drain the tensor-core results =================
        for (uint32_t g = 0; g < n_groups; g++, gg++) {
            const uint32_t region = gg % NR;
            // TMEM -> f32 tile.  Warp w reads lanes [32 (w & 3), +32) (= tile columns m) and the
            // half (w >> 2) of the group's 32 output rows.
            {
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
                float *t = tmp + size_t(m + carry) * r_pad + g * grp_rows + half * 16;  // behind the columns carried over
                const uint32_t nv = jn > half * 16 ? jn - half * 16 : 0;  // rows of this half that exist
                float2 v2[8];
#pragma unroll
                for (int e = 0; e < 16; e += 2) {  // value = (hi 2^14 + mid 2^7 + lo) 2^-s, two rows per f32x2 op
                    const float2 fh = make_float2(float(int(hi[e])), float(int(hi[e + 1])));
                    const float2 fl = make_float2(float(int(mid[e]) * 128 + int(lo[e])), float(int(mid[e + 1]) * 128 + int(lo[e + 1])));
                    float2 r = make_float2(0.f, 0.f);
                    ffma2(r, fl, scale);
                    ffma2(r, fh, scale_hi);
                    v2[e / 2] = r;
                }
                if (nv >= 16) {  // lanes = consecutive columns, r_pad odd: conflict-free
#pragma unroll
                    for (int e = 0; e < 16; e += 2) { t[e] = v2[e / 2].x; t[e + 1] = v2[e / 2].y; }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; e += 2) {
                        if (uint32_t(e) < nv) t[e] = v2[e / 2].x;
                        if (uint32_t(e + 1) < nv) t[e + 1] = v2[e / 2].y;
                    }
                }
            }
        }
        if (chunk + 1 < n_chunks) stage_htab(chunk + 1);  // lands during this chunk's horizontal stage
        else cp_async_commit();
        cp_async_wait<1>();  // this chunk's table slice (committed one chunk ago) has landed
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // consumers only: the tile is complete
        // ================= horizontal stage: CUDA cores =================
        if (h_warp) {
            const float *hw_s = hw_s0 + (chunk & 1) * htab_words;
            const uint32_t *hinfo_s = reinterpret_cast<const uint32_t *>(hw_s + size_t(max_pairs) * 2 * S);
            const uint32_t n_pairs = rec_pairs & 0xffffu;
            const bool odd = rec_pairs >> 16;  // the last pair has no second pixel: its weights are zero, its address the first pixel's
            // staging byte addresses of this lane's channel in rows ra / rb: the row segment keeps the
            // alignment phase of its canvas address so that staged words are canvas words
            uint32_t sa = out_u + (h_ra * out_stride) * 4 + ((h_pha + o_first * h_cout) & 3u) + h_ch;
            uint32_t sb = out_u + (h_rb * out_stride) * 4 + ((h_phb + o_first * h_cout) & 3u) + h_ch;
            // Software pipeline over pixel pairs, two register sets (A, B) in ping-pong: the shared-memory
            // reads of the next pair are issued before the FMAs of the current one.  (The second pixel
            // of a last odd pair lies inside the tile and has zero weights.)
            struct PairRegs {
                float2 va, vb;
                float4 w0, w1, w2, w3;
                uint32_t cnt;
            };
            // running shared-memory byte addresses of the next pair to load: rows ra / rb of the pair's
            // first pixel (the second is cstep4 further), its 16 weights and its count; they stop
            // advancing at the last pair, so the look-ahead loads never leave the chunk
            uint32_t a_va = smem_u32(tmp + size_t(hstart + h_ch) * r_pad + h_ra), a_vb = smem_u32(tmp + size_t(hstart + h_ch) * r_pad + h_rb);
            uint32_t a_w = smem_u32(hw_s), a_n = smem_u32(hinfo_s);
            const uint32_t cstep4 = C * r_pad * 4, pstep4 = 2 * cstep4;
            uint32_t left = n_pairs - 1;  // pairs after the one the addresses point to
            auto load = [&](PairRegs &P) {
                const uint32_t second = (odd && left == 0) ? 0u : cstep4;
                P.va = make_float2(lds_f32(a_va), lds_f32(a_vb));
                P.vb = make_float2(lds_f32(a_va + second), lds_f32(a_vb + second));
                P.w0 = lds_f32x4(a_w); P.w1 = lds_f32x4(a_w + 16); P.w2 = lds_f32x4(a_w + 32); P.w3 = lds_f32x4(a_w + 48);
                P.cnt = lds_u32(a_n);
                const uint32_t go = left ? 1u : 0u;
                left -= go;
                a_va += go * pstep4; a_vb += go * pstep4; a_w += go * 64; a_n += go * 4;
            };
            auto compute = [&](const PairRegs &P) {
                const float wa[S] = {P.w0.x, P.w0.y, P.w0.z, P.w0.w, P.w1.x, P.w1.y, P.w1.z, P.w1.w};
                const float wb[S] = {P.w2.x, P.w2.y, P.w2.z, P.w2.w, P.w3.x, P.w3.y, P.w3.z, P.w3.w};
#pragma unroll
                for (int j = 0; j < S; j++) ffma2(hacc[j], P.va, wa[j]);
#pragma unroll
                for (int j = 0; j < S; j++) ffma2(hacc[j], P.vb, wb[j]);
#pragma unroll 1
                for (uint32_t n = 0; n < P.cnt; n++) {  // uniform over the CTA: write out slot 0, shift the rest down
                    const uint32_t ua = round_u8(hacc[0].x), ub = round_u8(hacc[0].y);
#pragma unroll
                    for (int j = 0; j + 1 < S; j++) hacc[j] = hacc[j + 1];
                    hacc[S - 1] = make_float2(0.f, 0.f);
                    if constexpr (C == 1 || C == 3) {
                        // Opaque pixels: the overlay onto the fill colour returns the pixel itself (blend_rgba,
                        // alpha 255), so every lane stages its own channel byte and nothing is gathered.
                        if (h_lane) sts8(sa, ua);
                        if (h_has_b) sts8(sb, ub);
                        if (C == 1 && h_epi != EPI_PLAIN) {  // L -> (l, l, l, 255)
                            if (h_lane) { sts8(sa + 1, ua); sts8(sa + 2, ua); }
                            if (h_has_b) { sts8(sb + 1, ub); sts8(sb + 2, ub); }
                        }
                    } else if (h_epi == EPI_PLAIN) {  // LA / RGBA out as they are: every lane stages its own channel byte
                        if (h_lane) sts8(sa, ua);
                        if (h_has_b) sts8(sb, ub);
                    } else {
                        // gather the pixel's channels from the C lanes of this row slot; the channel-0 lane
                        // stages row ra, the channel-1 lane row rb
                        const uint32_t mine = ua | ub << 8;
                        uint32_t u[4] = {0, 0, 0, 0};
                        const uint32_t sel = h_ch == 1 ? 8u : 0u;
#pragma unroll
                        for (int k = 0; k < C; k++) u[k] = (__shfl_sync(0xffffffffu, mine, int(lane - h_ch) + k) >> sel) & 0xffu;
                        if (h_ch == 0 ? h_lane : (h_ch == 1 && h_has_b)) {
                            const uint32_t s0 = (h_ch == 1 ? sb : sa) - h_ch;
                            uint32_t px = to_rgba_packed(u, C);
                            if (h_epi == EPI_BLEND_FILL) px = blend_rgba(h_fill, px);
#pragma unroll
                            for (int k = 0; k < 4; k++) sts8(s0 + k, px >> (8 * k));
                        }
                    }
                    sa += h_cout;
                    sb += h_cout;
                }
            };
            PairRegs A, B;
            load(A);
            for (uint32_t p = 0; p < n_pairs; p += 2) {
                load(B);
                compute(A);
                load(A);
                if (p + 1 < n_pairs) compute(B);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // the tile may be overwritten; the staged pixels are complete
        // a pixel cut by the chunk boundary: its leading columns move to the front of the tile
        if (chunk + 1 < n_chunks) {
            const uint32_t cn = __ldg(crec + 4 * (chunk + 1) + 2);
            if (cn) {
                const float *from = tmp + size_t(carry + TC_M - cn) * r_pad;
                for (uint32_t k = tid; k < cn * r_pad; k += NT) tmp[k] = from[k];
                asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // before the next drain overwrites the source columns
            }
        }
        // ================= write out the pixels finished in this chunk =================
        if (const uint32_t nb = o_count * h_cout; nb && h_words) {
            // RGBA rows at aligned addresses: every staged word is a whole canvas word
            uint32_t *gq = reinterpret_cast<uint32_t *>(h_row0) + o_first + (tid >> 3) * (h_pitch >> 2) + (tid & 7);
            const uint32_t *sq = out_s + (tid >> 3) * out_stride + (tid & 7);
            const uint32_t g_step = (NT / 8) * (h_pitch >> 2), s_step = (NT / 8) * out_stride;
            for (uint32_t r = tid >> 3; r < h_rows; r += NT / 8, gq += g_step, sq += s_step)
                for (uint32_t k = tid & 7; k < o_count; k += 8) gq[k - (tid & 7)] = sq[k - (tid & 7)];
        } else if (nb) {
            for (uint32_t r = tid >> 3; r < h_rows; r += NT / 8) {  // 8 lanes per row segment
                uint8_t *g0 = h_row0 + size_t(r) * h_pitch + size_t(o_first) * h_cout;
                const uint32_t ph = uint32_t(reinterpret_cast<uintptr_t>(g0)) & 3u;
                const uint32_t nw = (ph + nb + 3) >> 2;
                for (uint32_t k = tid & 7; k < nw; k += 8) {
                    const uint32_t wv = out_s[r * out_stride + k];
                    uint8_t *gw = g0 - ph + 4 * k;
                    const int lo = int(ph) - int(4 * k), hi = int(ph + nb) - int(4 * k);  // the word's valid bytes: [lo, hi)
                    if (lo <= 0 && hi >= 4) {
                        *reinterpret_cast<uint32_t *>(gw) = wv;
                    } else {
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (b >= lo && b < hi) gw[b] = uint8_t(wv >> (8 * b));
                    }
                }
            }
        }
    }
    }  // consumer warps
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

// ---- vertical Gaussian pass of the blur: the same pipeline, consumers write the f32 intermediate ----
// One CTA per band of <= 24 groups x 32 rows; per 128-element column chunk and group the consumers
// drain the TMEM region and store value(row, column) to dst[row][column] -- lanes are consecutive
// columns, so every store of a warp is one 128-byte line.  No shared-memory tile, no horizontal stage.
__global__ void __launch_bounds__(NT_ALL, 1) blur_v_tc_kernel(const BlurVTcItem *__restrict__ items, const CUtensorMap *__restrict__ tmaps,
                                                            const uint8_t *__restrict__ tb, const uint32_t *__restrict__ tinfo) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ BlurVTcItem it_s;
    __shared__ __align__(8) uint64_t mbar[NR], tmem_free[NR], a_full[NA_MAX], b_full[NB];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t grp[4 * 24];
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        it_s = items[blockIdx.x];
        for (uint32_t r = 0; r < NR; r++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[r])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&tmem_free[r])));
        }
        for (uint32_t r = 0; r < NA_MAX; r++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&a_full[r])));
        for (uint32_t r = 0; r < NB; r++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&b_full[r])));
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
    const BlurVTcItem &it = it_s;
    for (uint32_t k = tid; k < 4 * it.n_groups; k += NT_ALL) grp[k] = tinfo[it.grp_off + k];
    __syncthreads();
    TcPipe pipe;
    pipe.mbar = smem_u32(&mbar[0]); pipe.tmem_free = smem_u32(&tmem_free[0]); pipe.a_full = smem_u32(&a_full[0]); pipe.b_full = smem_u32(&b_full[0]);
    pipe.sA_u = smem_u32(smem); pipe.sB_u = pipe.sA_u + it.n_a * it.kg_max * TC_M;
    pipe.kg_max = it.kg_max; pipe.n_a = it.n_a; pipe.n_chunks = it.n_chunks; pipe.n_groups = it.n_groups;
    pipe.x0 = 0; pipe.tmem_base = tmem_base_s; pipe.grp = grp; pipe.pf_ahead = 0;
    if (warp == NT / 32 + 1) {
        if (elect_one()) tc_source_role(pipe, tmaps + blockIdx.x);
    } else if (warp == NT / 32 + 2) {
        if (elect_one()) tc_weight_role(pipe, tb);
    } else if (warp == NT / 32) {
        if (elect_one()) tc_mma_role(pipe);
    } else {
        const float scale = it.scale, scale_hi = it.scale * 16384.0f;
        const uint32_t half = warp >> 2, m = (warp & 3) * 32 + lane, n_e = it.n_e;
        uint32_t gg = 0;
        for (uint32_t chunk = 0; chunk < it.n_chunks; chunk++) {
            const uint32_t col = chunk * TC_M + m;
            for (uint32_t g = 0; g < it.n_groups; g++, gg++) {
                const uint32_t region = gg % NR;
                mbar_wait(smem_u32(&mbar[region]), (gg / NR) & 1);  // the MMAs of group gg have retired
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = pipe.tmem_base + region * TC_N + (((warp & 3) * 32u) << 16) + half * 16;
                uint32_t hi[16], mid[16], lo[16];
                tmem_ld16(taddr, hi);
                tmem_ld16(taddr + 32, mid);
                tmem_ld16(taddr + 64, lo);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_free[region])) : "memory");
                const uint32_t jn = grp[4 * g + 3];
                const uint32_t nv = jn > half * 16 ? jn - half * 16 : 0;  // rows of this half that exist
                float *q = it.dst + size_t(it.band_r0 + g * TC_GROUP_ROWS + half * 16) * n_e + col;
                if (col < n_e) {
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const float v = fmaf(float(int(hi[e])), scale_hi, float(int(mid[e]) * 128 + int(lo[e])) * scale);
                        if (uint32_t(e) < nv) q[size_t(e) * n_e] = v;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(TMEM_COLS));
}


// ================= both passes on the tensor cores (fused_resample_tc2_kernel) =================
// The vertical pass is the banded integer contraction above.  Its results no longer go through a
// scatter loop on the CUDA cores: the consumers write them as f16 hi / lo halves (v = hi + lo to
// 2^-14) into the MN-major operand tile T[group row][tile column], and the horizontal pass is a
// second contraction per chunk and per tile of 128 band rows (four groups of 32):
//
//   D2[128 rows x N2] (f32, TMEM) = T_hi . W_hi + T_lo . W_hi + T_hi . W_lo     (kind::f16, K = 128 columns)
//
// W[N2][128] maps the chunk's tile columns (byte b of the row = channel b % c of pixel b / c) to
// accumulator columns (ring position of the output pixel) * c + channel; the host builds it per
// chunk with the crate's weights x 16 split into f16 halves.  A chunk's D2 holds the partial sums
// of the <= N2 / c outputs whose windows meet the chunk; each consumer thread owns one band row
// and half of the ring in registers, adds the partial sums and, when a pixel's window has ended,
// rounds and stores it.  The CUDA cores only drain, convert and store.
constexpr int NT_ALL2 = NT + 128;   // + the vertical MMA warp, the source / vertical-weight TMA warps, the horizontal (weights + MMA) warp
constexpr uint32_t NR2 = 4;          // TMEM regions of the vertical pass (96 columns each, from column 128)
constexpr uint32_t TMEM_V0 = 128;    // columns [0, 128): D2 of the two row tiles

struct Tc2Pipe {
    uint32_t t_ready, d2_full, d2_free, wh_full, wh_free;  // mbarrier arrays (8 bytes per entry)
    uint32_t sT_u, t_bytes;                       // T hi tile; the lo tile follows t_bytes later
    uint32_t sWh_u, n_wh;                         // horizontal weight slots [n_wh][hi N2 x 128 | lo N2 x 128] f16
    uint32_t n_mt;                                // row tiles (1 or 2)
};

// Horizontal thread: fetches the chunk's weight tiles (one bulk copy into slot chunk % n_wh) and issues the 24
// horizontal MMAs of a row tile as soon as the consumers have written it.  A thread of its own: the vertical
// MMA thread never waits behind a row tile that is not ready, and the tensor pipe interleaves the two streams.
template <uint32_t N2>
__device__ __forceinline__ void tc2_h_role(const TcPipe &p, const Tc2Pipe &h, const uint8_t *tb, const uint32_t *hrec) {
    constexpr uint32_t IDESC_H = (1u << 4) | (1u << 15) | ((N2 >> 3) << 17) | ((TC_M >> 4) << 24);  // f16 x f16 -> f32, A MN-major, B K-major
    long long w_wh = 0, w_tr = 0, w_df = 0, w_wf = 0, i_h = 0;
    const long long t_start = clock64();
    auto load_wh = [&](uint32_t ch) {
        const uint32_t slot = ch % h.n_wh, bar = h.wh_full + 8 * slot;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(N2 * 512u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(h.sWh_u + slot * N2 * 512u),
                     "l"(tb + __ldg(hrec + 3 * ch)), "r"(N2 * 512u), "r"(bar)
                     : "memory");
    };
    load_wh(0);
    for (uint32_t ch = 0; ch < p.n_chunks; ch++) {
        const uint32_t slot = ch % h.n_wh;
        if (h.n_wh == 2 && ch + 1 < p.n_chunks) {  // the other slot: free once the horizontal MMAs of chunk ch - 1 have retired
            if (ch >= 1) PWR(w_wf, h.wh_free + 8 * ((ch + 1) & 1), ((ch - 1) >> 1) & 1);
            load_wh(ch + 1);
        }
        PWR(w_wh, h.wh_full + 8 * slot, (ch / h.n_wh) & 1);  // the chunk's weight tiles have landed
        for (uint32_t mt = 0; mt < h.n_mt; mt++) {
            PWR(w_tr, h.t_ready + 8 * mt, ch & 1);                    // the consumers have written the tile's rows
            if (ch > 0) PWR(w_df, h.d2_free + 8 * mt, (ch - 1) & 1);  // ... and read the previous chunk's D2
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef TC2_PROF
            const long long ti0 = clock64();
#endif
            const uint32_t a_hi = h.sT_u + mt * 32768u, a_lo = a_hi + h.t_bytes;
            const uint32_t b_hi = h.sWh_u + slot * N2 * 512u, b_lo = b_hi + N2 * 256u;
            const uint32_t d_tmem = p.tmem_base + mt * N2;
#pragma unroll
            for (int combo = 0; combo < 3; combo++) {
                uint64_t da = umma_desc(combo == 1 ? a_lo : a_hi, 128, 2048);  // LBO: next 8 columns (K), SBO: next 8 rows (M)
                uint64_t db = umma_desc(combo == 2 ? b_lo : b_hi, 128, 2048);  // LBO: next 8 columns (K), SBO: next 8 accumulator columns (N)
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {
                    if (combo == 0 && ks == 0)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da),
                                     "l"(db), "r"(IDESC_H)
                                     : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da),
                                     "l"(db), "r"(IDESC_H)
                                     : "memory");
                    da += 256 >> 4;  // 16 columns = two core matrices along K
                    db += 256 >> 4;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(h.d2_full + 8 * mt) : "memory");
            if (mt + 1 == h.n_mt) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(h.wh_free + 8 * slot) : "memory");
#ifdef TC2_PROF
            i_h += clock64() - ti0;
#endif
        }
        if (h.n_wh == 1 && ch + 1 < p.n_chunks) {  // the only slot: its tiles were read once the chunk's MMAs have retired
            PWR(w_wf, h.wh_free, ch & 1);
            load_wh(ch + 1);
        }
    }
#ifdef TC2_PROF
    if (blockIdx.x == 300)
        printf("horizontal thread: total %lld clk; waits: weights landed %lld, T tile %lld, D2 drained %lld, weight slot free %lld; issue %lld\n", clock64() - t_start, w_wh,
               w_tr, w_df, w_wf, i_h);
#else
    (void)t_start; (void)w_wh; (void)w_tr; (void)w_df; (void)w_wf; (void)i_h;
#endif
}

// Vertical MMA thread: tc_mma_role with NR2 regions behind the D2 columns
__device__ __forceinline__ void tc2_v_role(const TcPipe &p) {
    const uint32_t total = p.n_chunks * p.n_groups;
    long long w_b = 0, w_a = 0, w_tf = 0, i_v = 0;
    const long long t_start = clock64();
    uint32_t g = 0, slot = 0, suse = 0;
    for (uint32_t gg = 0; gg < total; gg++) {
        const uint32_t region = gg % NR2, ruse = gg / NR2, bslot = gg % NB;
        const uint32_t kg = p.grp[4 * g + 1];
        PWR(w_b, p.b_full + 8 * bslot, (gg / NB) & 1);
        PWR(w_a, p.a_full + 8 * slot, suse & 1);
        if (ruse > 0) PWR(w_tf, p.tmem_free + 8 * region, (ruse - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef TC2_PROF
        const long long tv0 = clock64();
#endif
        uint64_t da = umma_desc(p.sA_u + slot * p.kg_max * TC_M, 16, 1024, 2);
        uint64_t db = umma_desc(p.sB_u + bslot * TC_N * p.kg_max, 128, (kg / 16) * 128);
        const uint32_t d_tmem = p.tmem_base + TMEM_V0 + region * TC_N;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                     "l"(da), "l"(db), "r"(UMMA_IDESC)
                     : "memory");
        for (uint32_t ks = 1; ks < kg / 32; ks++) {
            da += SLAB >> 4;
            db += (2 * 128) >> 4;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                         "l"(da), "l"(db), "r"(UMMA_IDESC)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(p.mbar + 8 * region) : "memory");
#ifdef TC2_PROF
        i_v += clock64() - tv0;
#endif
        if (++slot == p.n_a) { slot = 0; suse++; }
        if (++g == p.n_groups) g = 0;
    }
#ifdef TC2_PROF
    if (blockIdx.x == 300)
        printf("vertical MMA thread: total %lld clk; waits: weights %lld, source rows %lld, TMEM region %lld; issue %lld\n", clock64() - t_start, w_b, w_a, w_tf, i_v);
#else
    (void)t_start; (void)w_b; (void)w_a; (void)w_tf; (void)i_v;
#endif
}

// INV: inverse on load (a template parameter: the consumers' conversion loop is what bounds them, and a runtime test in
// it cost the plain kernel 17 %)
template <int C, bool INV = false>
__global__ void __launch_bounds__(NT_ALL2, 1) fused_resample_tc2_kernel(const FusedTcItem *__restrict__ items, const CUtensorMap *__restrict__ tmaps,
                                                                    const uint8_t *__restrict__ tb, const uint32_t *__restrict__ tinfo) {
    constexpr uint32_t N2 = C == 3 ? 48 : 64, RP = N2 / C, HP = RP / 2;  // accumulator columns, ring of outputs, ring positions per thread
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ FusedTcItem it_s;
    __shared__ __align__(8) uint64_t mbar[NR2], tmem_free[NR2], a_full[NA_MAX], b_full[NB];
    __shared__ __align__(8) uint64_t t_ready[2], d2_full[2], d2_free[2], wh_full[2], wh_free[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t grp[4 * 8];
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
#ifdef TC2_PROF
    const long long t_entry = clock64();
    unsigned long long g_entry;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_entry));
    long long t_roles = 0, t_p1 = 0, t_p2 = 0;
#endif
    if (tid == 0) {
        it_s = items[blockIdx.x];
        for (uint32_t r = 0; r < NR2; r++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[r])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&tmem_free[r])));
        }
        for (uint32_t r = 0; r < NA_MAX; r++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&a_full[r])));
        for (uint32_t r = 0; r < NB; r++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&b_full[r])));
        for (uint32_t r = 0; r < 2; r++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&t_ready[r])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&d2_full[r])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&d2_free[r])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&wh_full[r])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&wh_free[r])));
        }
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
#ifdef TC2_PROF
    t_p1 = clock64();
#endif
    if (warp < NT / 32) fill_bars(it, warp, lane, NT / 32);
#ifdef TC2_PROF
    t_p2 = clock64();
#endif

    // ---- shared-memory carve-up: T hi | T lo | source slots | vertical weight slots | horizontal weight slots.
    // The horizontal MMAs always read 128 rows per tile: behind a tile of fewer groups they read into
    // whatever follows (T lo, the source slots) -- rows nobody drains.
    const uint32_t kg_max = it.kg_max, n_a = it.n_a, n_groups = it.n_groups, n_chunks = it.n_chunks;
    const uint32_t t_bytes = n_groups * 32u * 256u;
    uint8_t *sT = smem;
    uint8_t *sA = sT + 2 * size_t(t_bytes);
    uint8_t *sB = sA + size_t(n_a) * kg_max * TC_M;
    uint8_t *sWh = sB + NB * size_t(TC_N) * kg_max;
    // pixels finished by a chunk wait here, [band row][out_stride words], and leave as whole words per row segment
    uint32_t *out_s = reinterpret_cast<uint32_t *>(sWh + size_t(it.n_wh) * N2 * 512u);
    for (uint32_t k = tid; k < 4 * n_groups; k += NT_ALL2) grp[k] = tinfo[it.grp_off + k];
    __syncthreads();

    TcPipe pipe;
    pipe.mbar = smem_u32(&mbar[0]); pipe.tmem_free = smem_u32(&tmem_free[0]); pipe.a_full = smem_u32(&a_full[0]); pipe.b_full = smem_u32(&b_full[0]);
    pipe.sA_u = smem_u32(sA); pipe.sB_u = smem_u32(sB); pipe.kg_max = kg_max; pipe.n_a = n_a; pipe.n_chunks = n_chunks; pipe.n_groups = n_groups;
    pipe.x0 = it.b0; pipe.tmem_base = tmem_base; pipe.grp = grp; pipe.pf_ahead = TC_PF_AHEAD;
    Tc2Pipe hp;
    hp.t_ready = smem_u32(&t_ready[0]); hp.d2_full = smem_u32(&d2_full[0]); hp.d2_free = smem_u32(&d2_free[0]); hp.wh_full = smem_u32(&wh_full[0]); hp.wh_free = smem_u32(&wh_free[0]);
    hp.sT_u = smem_u32(sT); hp.t_bytes = t_bytes; hp.sWh_u = smem_u32(sWh); hp.n_wh = it.n_wh; hp.n_mt = (n_groups + 3) / 4;
    const uint32_t *hrec = tinfo + it.hrec_off;

#ifdef TC2_PROF
    t_roles = clock64();
#endif
    if (warp == NT / 32 + 1) {
        if (elect_one()) tc_source_role<NR2>(pipe, tmaps + blockIdx.x);
    } else if (warp == NT / 32 + 2) {
        if (elect_one()) tc_weight_role<NR2>(pipe, tb);
    } else if (warp == NT / 32 + 3) {
        if (elect_one()) tc2_h_role<N2>(pipe, hp, tb, hrec);
    } else if (warp == NT / 32) {
        if (elect_one()) tc2_v_role(pipe);
    } else {
        // ================= consumer warps =================
        const float scale = it.scale, scale_hi = it.scale * 16384.0f;
        const uint32_t q = warp & 3, half = warp >> 2, m = q * 32 + lane;
        // inverse on load applies to the colour channels of this thread's tile column (byte b0 + 128 chunk + m of the row)
        const bool inv_on = INV && !((C == 2 || C == 4) && ((it.b0 + m) % C) == C - 1);
        const uint32_t grp_rows = it.grp_rows, n_mt = hp.n_mt;
        const uint32_t h_cout = it.c_out, h_pitch = it.dst_pitch, h_rows = it.band_rows, h_epi = it.epi, h_fill = it.fill;
        const uint32_t out_stride = it.out_stride, out_u = smem_u32(out_s);
        uint8_t *const h_row0 = it.dst + size_t(it.dst_y + it.band_r0) * it.dst_pitch + size_t(it.dst_x) * h_cout;  // first canvas byte of the band
        const uint32_t h_ph0 = uint32_t(reinterpret_cast<uintptr_t>(h_row0));
        const bool h_words = h_cout == 4 && ((reinterpret_cast<uintptr_t>(h_row0) | h_pitch) & 3) == 0;
        // the pixels a chunk finished, staged by drain_d2, leave as whole words per row segment
        auto write_out = [&](uint32_t chunk) {
            const uint32_t o_first = __ldg(hrec + 3 * chunk + 1), o_count = __ldg(hrec + 3 * chunk + 2);
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // every row's pixels are staged
            if (const uint32_t nb = o_count * h_cout; nb && h_words) {
                uint32_t *gq = reinterpret_cast<uint32_t *>(h_row0) + o_first + (tid >> 3) * (h_pitch >> 2) + (tid & 7);
                const uint32_t *sq = out_s + (tid >> 3) * out_stride + (tid & 7);
                const uint32_t g_step = (NT / 8) * (h_pitch >> 2), s_step = (NT / 8) * out_stride;
                for (uint32_t r = tid >> 3; r < h_rows; r += NT / 8, gq += g_step, sq += s_step)
                    for (uint32_t k = tid & 7; k < o_count; k += 8) gq[k - (tid & 7)] = sq[k - (tid & 7)];
            } else if (nb) {
                // whole words, 8 lanes per row segment (the staged words are the canvas words: a segment keeps its alignment phase) ...
                for (uint32_t r = tid >> 3; r < h_rows; r += NT / 8) {
                    uint8_t *g0 = h_row0 + size_t(r) * h_pitch + size_t(o_first) * h_cout;
                    const uint32_t ph = uint32_t(reinterpret_cast<uintptr_t>(g0)) & 3u;
                    const uint32_t k_a = ph ? 1u : 0u, k_b = (ph + nb) >> 2;  // words [k_a, k_b) of the segment are whole
                    for (uint32_t k = k_a + (tid & 7); k < k_b; k += 8)
                        *reinterpret_cast<uint32_t *>(g0 - ph + 4 * k) = out_s[r * out_stride + k];
                }
                // ... then the partial words at the two ends, one row per thread (in the loop above they put every warp through a
                // byte-store path for the sake of two of its lanes)
                for (uint32_t r = tid; r < h_rows; r += NT) {
                    uint8_t *g0 = h_row0 + size_t(r) * h_pitch + size_t(o_first) * h_cout;
                    const uint32_t ph = uint32_t(reinterpret_cast<uintptr_t>(g0)) & 3u;
                    const uint32_t k_t = (ph + nb - 1) >> 2;  // the word that holds the last byte
#pragma unroll
                    for (uint32_t e = 0; e < 2; e++) {
                        const uint32_t k = e ? k_t : 0u;
                        const int lo = int(ph) - int(4 * k), hi = int(ph + nb) - int(4 * k);  // the word's valid bytes: [lo, hi)
                        if ((e && k_t == 0) || (lo <= 0 && hi >= 4)) continue;
                        const uint32_t wv = out_s[r * out_stride + k];
                        uint8_t *gw = g0 - ph + 4 * k;
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (b >= lo && b < hi) gw[b] = uint8_t(wv >> (8 * b));
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // the staging buffer may be overwritten
        };
        float acc[2][HP][C];  // partial sums of this thread's band row (one per row tile) for its half of the output ring
#pragma unroll
        for (int t = 0; t < 2; t++)
#pragma unroll
            for (int k = 0; k < int(HP); k++)
#pragma unroll
                for (int c = 0; c < C; c++) acc[t][k][c] = 0.f;

        // D2 of (chunk, row tile) -> registers; pixels whose window ended in the chunk are rounded and stored
        long long w_v = 0, w_d2 = 0, t_d2 = 0;
        const long long t_start = clock64();
        auto drain_d2 = [&](uint32_t chunk, uint32_t mt, float (&a)[HP][C]) {
#ifdef TC2_PROF
            const long long td0 = clock64();
#endif
            PW(w_d2, hp.d2_full + 8 * mt, chunk & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[N2 / 2];
            const uint32_t taddr = tmem_base + mt * N2 + ((q * 32u) << 16) + half * (N2 / 2);
            if constexpr (N2 == 48) {
                tmem_ld16(taddr, v);
                tmem_ld8(taddr + 16, v + 16);
            } else {
                tmem_ld16(taddr, v);
                tmem_ld16(taddr + 16, v + 16);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hp.d2_free + 8 * mt) : "memory");
            const uint32_t g = mt * 4 + q;
            const bool row_ok = g < n_groups && lane < grp[4 * min(g, n_groups - 1) + 3];
            const uint32_t row = min(g * grp_rows + lane, h_rows - 1);
            const uint32_t fin_first = __ldg(hrec + 3 * chunk + 1), n_fin = __ldg(hrec + 3 * chunk + 2);
            // staging byte address of this row's first finished pixel: the row segment keeps the alignment phase
            // of its canvas address so that staged words are canvas words
            const uint32_t sp = out_u + row * out_stride * 4 + ((h_ph0 + row * h_pitch + fin_first * h_cout) & 3u);
#pragma unroll
            for (int k = 0; k < int(HP); k++) {
#pragma unroll
                for (int c = 0; c < C; c++) a[k][c] += __uint_as_float(v[k * C + c]);
                const uint32_t d = (half * HP + k - fin_first) & (RP - 1);  // ring distance from the first output finished here
                if (d < n_fin) {  // uniform over the warp
                    if (row_ok) {
                        uint32_t u[4] = {0, 0, 0, 0};
#pragma unroll
                        for (int c = 0; c < C; c++) u[c] = round_u8(a[k][c] * (1.0f / TC2_WSCALE));
                        const uint32_t sq = sp + d * h_cout;
                        if (h_epi == EPI_PLAIN || (h_epi & EPI_GRAY)) {  // (gray canvas: C = 1, the luma byte)
#pragma unroll
                            for (int c = 0; c < C; c++) sts8(sq + c, u[c]);
                        } else {
                            uint32_t px = to_rgba_packed(u, C);
                            if ((h_epi & EPI_MASK) == EPI_BLEND_FILL) px = blend_rgba(h_fill, px);
                            if (h_epi & EPI_RGB8) {  // to_rgb8 of the result: the alpha byte stays behind (h_cout = 3)
#pragma unroll
                                for (int c = 0; c < 3; c++) sts8(sq + c, px >> (8 * c));
                            } else if (h_words) {
                                asm volatile("st.shared.b32 [%0], %1;" ::"r"(sq), "r"(px) : "memory");
                            } else {
#pragma unroll
                                for (int c = 0; c < 4; c++) sts8(sq + c, px >> (8 * c));
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < C; c++) a[k][c] = 0.f;
                }
            }
#ifdef TC2_PROF
            t_d2 += clock64() - td0;
#endif
        };

        uint32_t gg = 0;
        for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
            for (uint32_t g = 0; g < n_groups; g++, gg++) {
                // the tile's rows are about to be overwritten: its horizontal MMAs of the previous chunk must have
                // finished -- which is also when their D2 can be drained
                if ((g & 3) == 0 && chunk > 0) {
                    if ((g >> 2) == 0) drain_d2(chunk - 1, 0, acc[0]);
                    else drain_d2(chunk - 1, 1, acc[1]);
                    if ((g >> 2) + 1 == n_mt) write_out(chunk - 1);
                }
                const uint32_t region = gg % NR2;
                PW(w_v, smem_u32(&mbar[region]), (gg / NR2) & 1);  // the vertical MMAs of group gg have retired
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + TMEM_V0 + region * TC_N + ((q * 32u) << 16) + half * 16;
                uint32_t hi[16], mid[16], lo[16];
                tmem_ld16(taddr, hi);
                tmem_ld16(taddr + 32, mid);
                tmem_ld16(taddr + 64, lo);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_free[region])) : "memory");
                uint32_t ph[8], pl[8];  // rows 2e, 2e + 1 of this half: f16x2 of the high and of the low halves
#pragma unroll
                for (int e = 0; e < 16; e += 2) {  // value = (hi 2^14 + mid 2^7 + lo) 2^-s, two rows per f32x2 op
                    const float2 fh = make_float2(float(int(hi[e])), float(int(hi[e + 1])));
                    const float2 fl = make_float2(float(int(mid[e]) * 128 + int(lo[e])), float(int(mid[e + 1]) * 128 + int(lo[e + 1])));
                    float2 r = make_float2(0.f, 0.f);
                    ffma2(r, fl, scale);
                    ffma2(r, fh, scale_hi);
                    if (INV && inv_on) {  // inverse on load: 255 sum q - sum q x (colour channels; alpha columns pass)
                        const uint32_t ri = g * grp_rows + half * 16 + e, rmax = h_rows - 1;  // (rows past the band are never stored)
                        r.x = __uint_as_float(__ldg(tinfo + it.inv_off + min(ri, rmax))) - r.x;
                        r.y = __uint_as_float(__ldg(tinfo + it.inv_off + min(ri + 1, rmax))) - r.y;
                    }
                    const uint32_t h2 = pack_f16x2(r.x, r.y);
                    const float2 back = unpack_f16x2(h2);
                    ph[e / 2] = h2;
                    pl[e / 2] = pack_f16x2(r.x - back.x, r.y - back.y);
                }
                // T[row][column m]: core matrices of 8 rows x 8 columns, 16 bytes = 8 rows of one column; a warp's
                // 32 columns are 512 contiguous bytes per block of 8 rows
                const uint32_t t0 = hp.sT_u + (g * 4 + half * 2) * 2048u + m * 16u;
                sts128(t0, ph[0], ph[1], ph[2], ph[3]);
                sts128(t0 + 2048, ph[4], ph[5], ph[6], ph[7]);
                sts128(t0 + t_bytes, pl[0], pl[1], pl[2], pl[3]);
                sts128(t0 + t_bytes + 2048, pl[4], pl[5], pl[6], pl[7]);
                if ((g & 3) == 3 || g + 1 == n_groups) {  // the row tile is complete: hand it to the tensor core
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hp.t_ready + 8 * (g >> 2)) : "memory");
                }
            }
        }
        drain_d2(n_chunks - 1, 0, acc[0]);
        if (n_mt > 1) drain_d2(n_chunks - 1, 1, acc[1]);
        write_out(n_chunks - 1);
#ifdef TC2_PROF
        if (blockIdx.x == 300 && lane == 0 && (warp == 0 || warp == 5))
            printf("consumer warp %u: total %lld clk; waits: vertical results %lld, D2 %lld; D2 drain + emit (incl. wait) %lld\n", warp, clock64() - t_start, w_v, w_d2, t_d2);
#else
        (void)t_start; (void)w_v; (void)w_d2; (void)t_d2;
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
#ifdef TC2_PROF
    if (tid == 0 && (blockIdx.x == 300 || blockIdx.x == 301 || blockIdx.x == 450 || blockIdx.x == 10)) {
        unsigned long long g_exit;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_exit));
        printf("CTA %u: entry at %llu ns, lifetime %llu ns = %lld clk, prologue %lld clk (init %lld, bars %lld, tables %lld)\n", blockIdx.x, g_entry, g_exit - g_entry, clock64() - t_entry, t_roles - t_entry, t_p1 - t_entry, t_p2 - t_p1, t_roles - t_p2);
    }
#endif
}

template <int C, bool INV>
void launch_tc2_variant(const FusedTcItem *d_items, const void *d_tmaps, uint32_t n_items, size_t smem, const uint8_t *d_b, const uint32_t *d_info,
                        LaunchCtx &lc) {
    auto kern = fused_resample_tc2_kernel<C, INV>;
    ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);  // a failure surfaces as the launch error
    lc.begin("fused_resample_tc2_kernel");
    kern<<<n_items, NT_ALL2, smem, lc.st>>>(d_items, static_cast<const CUtensorMap *>(d_tmaps), d_b, d_info);
    lc.end();
}

template <int C>
void launch_tc_variant(const FusedTcItem *d_items, const void *d_tmaps, uint32_t n_items, size_t smem, const uint8_t *d_b,
                       const float *d_w, const uint32_t *d_info, LaunchCtx &lc) {
    auto kern = fused_resample_tc_kernel<C>;
    ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
    lc.begin("fused_resample_tc_kernel");
    kern<<<n_items, NT_ALL, smem, lc.st>>>(d_items, static_cast<const CUtensorMap *>(d_tmaps), d_b, d_w, d_info);
    lc.end();
}

}  // namespace

int launch_fused_tc(const FusedTcItem *d_items, const void *d_tmaps, uint32_t n_items, uint32_t c, size_t smem, const uint8_t *d_b,
                    const float *d_w, const uint32_t *d_info, LaunchCtx &lc) {
    if (n_items == 0) return 0;
    if (c & 16) return launch_fused_tc3(d_items, d_tmaps, n_items, c & 39, smem, d_b, d_info, lc);  // ... with the horizontal sums kept in TMEM
    if (c & 8) {  // both passes on the tensor cores; bit 5: inverse on load
        switch (c & 39) {
        case 1: launch_tc2_variant<1, false>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
        case 2: launch_tc2_variant<2, false>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
        case 3: launch_tc2_variant<3, false>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
        case 4: launch_tc2_variant<4, false>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
        case 33: launch_tc2_variant<1, true>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
        case 34: launch_tc2_variant<2, true>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
        case 35: launch_tc2_variant<3, true>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
        case 36: launch_tc2_variant<4, true>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
        }
        return -1;
    }
    switch (c) {
    case 1: launch_tc_variant<1>(d_items, d_tmaps, n_items, smem, d_b, d_w, d_info, lc); return 1;
    case 2: launch_tc_variant<2>(d_items, d_tmaps, n_items, smem, d_b, d_w, d_info, lc); return 1;
    case 3: launch_tc_variant<3>(d_items, d_tmaps, n_items, smem, d_b, d_w, d_info, lc); return 1;
    case 4: launch_tc_variant<4>(d_items, d_tmaps, n_items, smem, d_b, d_w, d_info, lc); return 1;
    }
    return -1;
}

int launch_blur_v_tc(const BlurVTcItem *d_items, const void *d_tmaps, uint32_t n_items, size_t smem, const uint8_t *d_b,
                     const uint32_t *d_info, LaunchCtx &lc) {
    if (n_items == 0) return 0;
    ensure_dynamic_smem(reinterpret_cast<const void *>(blur_v_tc_kernel), smem);
    lc.begin("blur_v_tc_kernel");
    blur_v_tc_kernel<<<n_items, NT_ALL, smem, lc.st>>>(d_items, static_cast<const CUtensorMap *>(d_tmaps), d_b, d_info);
    lc.end();
    return 1;
}

}  // namespace fanlin

// ---- host: shared-memory budget, tensor maps ---------------------------------------------------------------------
#include <cudaTypedefs.h>

#include <algorithm>

namespace fanlin {

size_t fused_tc_smem_limit() {
    static size_t limit = 0;
    if (!limit) {
        size_t stat = 0;
        const void *kerns[8] = {reinterpret_cast<const void *>(fused_resample_tc_kernel<1>), reinterpret_cast<const void *>(fused_resample_tc_kernel<2>),
                                reinterpret_cast<const void *>(fused_resample_tc_kernel<3>), reinterpret_cast<const void *>(fused_resample_tc_kernel<4>),
                                reinterpret_cast<const void *>(fused_resample_tc2_kernel<1, false>), reinterpret_cast<const void *>(fused_resample_tc2_kernel<2, false>),
                                reinterpret_cast<const void *>(fused_resample_tc2_kernel<3, false>), reinterpret_cast<const void *>(fused_resample_tc2_kernel<4, false>)};
        bool ok = true;
        for (const void *k : kerns) {
            cudaFuncAttributes fa{};
            ok = ok && cudaFuncGetAttributes(&fa, k) == cudaSuccess;
            stat = std::max(stat, fa.sharedSizeBytes);
        }
        int dev = 0, optin = 0;
        ok = ok && cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess;
        if (!ok) { cudaGetLastError(); return 232448 - 4096; }  // no device (host-only tests): a conservative figure
        limit = size_t(optin) - stat;
    }
    return limit;
}

bool encode_row_tile_map(void *out, const void *base, uint32_t pitch, uint32_t rows, uint32_t box_rows, uint32_t width_bytes) {
    static PFN_cuTensorMapEncodeTiled enc = nullptr;
    if (!enc) {
        cudaDriverEntryPointQueryResult qr;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return false;
        enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    const cuuint64_t gdim[2] = {width_bytes ? width_bytes : pitch, rows};
    const cuuint64_t gstr[1] = {pitch};
    const cuuint32_t box[2] = {TC_M, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(static_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace fanlin
