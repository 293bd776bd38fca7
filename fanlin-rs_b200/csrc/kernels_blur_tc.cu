// Gaussian blur with BOTH passes on the sm_100a tensor cores and no intermediate in HBM (see blur.h, BlurTcItem).
// imageops::blur as called at reference src/handler.rs:250-255: vertical pass u8 -> f32, horizontal pass f32 -> u8,
// 2 R + 1 taps, windows truncated at the borders and renormalised.
//
// One CTA per band of <= 128 rows of one image (1 CTA / SM, 12 warps), swept in chunks of 128 bytes per row:
//   warp 9   source TMA thread: ONE tensor copy per chunk -- the band's rows and their halo, box_rows x 128 B, 128-byte
//            swizzle (two slots); rows / columns outside the image arrive as zeros;
//   warp 10  vertical-weight TMA thread: the s8 digit tile of the next group (two slots); every interior group shares
//            one tile, so after the first two groups of a band nothing is fetched any more -- the slot is re-armed
//            with a plain arrive;
//   warp 8   vertical MMA thread: kg / 32 tcgen05.mma kind::i8 per 32-row group (A = the group's window inside the
//            box, MN-major as it lies in the image; B = its digit tile) into one of two 96-column TMEM regions;
//   warps 0-7 consumers: drain a region, recombine the three digit sums exactly, split the f32 value into f16 hi / lo
//            and store them as the MN-major operand tile T[128 rows][128 bytes] (two buffers); between two groups they
//            drain one finished sub-step of the PREVIOUS chunk's horizontal pass from the ring;
//   warp 11  horizontal thread: holds the Toeplitz tile W[n_win][32] (f16 hi | lo, weights x 16) resident; per chunk and
//            per 32 bytes of T: D[128 rows x n_win] += T_hi W_hi + T_lo W_hi + T_hi W_lo (kind::f16, f32 accumulators)
//            into the ring of 256 TMEM columns at column (output byte mod 256), split in two where the window wraps.
// Sub-step jj covers input bytes [32 jj, 32 jj + 32) and output bytes [32 jj - r_pad, 32 jj + 32 + r_pad); when it has
// retired, output bytes [32 jj - r_pad, + 32) have all their taps.  The horizontal border renormalisation of the crate
// (weights divided by the sum of the taps inside the image) is one factor per column applied at the drain, with zeros
// outside the image -- the approach of kernels_blur.cu; the vertical pass needs none, its tiles come from the crate's
// own tap tables.
#include <cuda.h>

#include "blur.h"
#include "fused_device.cuh"
#include "fused_tc.h"
#include "kernels.h"
#include "tc_device.cuh"

namespace fanlin {

namespace {

constexpr int BT_NT = 256;          // consumer threads
constexpr int BT_NT_ALL = BT_NT + 128;
constexpr uint32_t BT_NRV = 2;      // TMEM regions of the vertical pass (96 columns each)
constexpr uint32_t BT_RING0 = 256;  // the ring: TMEM columns [256, 512); the vertical regions: [0, 192)
constexpr uint32_t BT_T_BYTES = 32768;  // one half (hi or lo) of a T buffer: 128 rows x 128 columns f16
constexpr uint32_t BT_STAGE_WARP = 1184;  // output staging per consumer warp: 16 bytes in front + 32 rows x 36 bytes, rounded to 16
constexpr uint32_t BT_SBO_W = 512;  // horizontal tile: 8 accumulator columns further = 4 core matrices of 128 bytes

__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

__global__ void __launch_bounds__(BT_NT_ALL, 1) blur_tc2_kernel(const BlurTcItem *__restrict__ items, const CUtensorMap *__restrict__ tmaps,
                                                                const uint8_t *__restrict__ tb, const uint32_t *__restrict__ tinfo,
                                                                const float *__restrict__ tw) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ BlurTcItem it_s;
    __shared__ __align__(8) uint64_t v_full[BT_NRV], v_free[BT_NRV], a_full[2], a_free[2], b_full[2], t_ready[2], d2_full[4], d2_free[4], wh_full[1];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t grp[16];
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        it_s = items[blockIdx.x];
        auto init = [](uint64_t *b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); };
        for (uint32_t r = 0; r < BT_NRV; r++) { init(&v_full[r], 1); init(&v_free[r], 8); }
        for (uint32_t r = 0; r < 2; r++) { init(&a_full[r], 1); init(&a_free[r], 1); init(&b_full[r], 1); init(&t_ready[r], 8); }
        for (uint32_t r = 0; r < 4; r++) { init(&d2_full[r], 1); init(&d2_free[r], 4); }  // a sub-step is drained by the four warps of its parity
        init(&wh_full[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const BlurTcItem &it = it_s;
    const uint32_t tmem_base = tmem_base_s;
    for (uint32_t k = tid; k < 4 * it.n_groups; k += BT_NT_ALL) grp[k] = tinfo[it.grp_off + k];
    __syncthreads();

    const uint32_t n_groups = it.n_groups, n_chunks = it.n_chunks, kg_max = it.kg_max;
    const uint32_t box_bytes = it.box_rows * TC_M;
    const uint32_t sT_u = smem_u32(smem);                      // 2 x (hi 32 KB | lo 32 KB)
    const uint32_t sA_u = sT_u + 4 * BT_T_BYTES;               // 2 x box
    const uint32_t sB_u = sA_u + 2 * box_bytes;                // 2 x [96][kg_max]
    const uint32_t sWh_u = sB_u + 2 * TC_N * kg_max;           // hi [n_win][32] | lo [n_win][32]
    const uint32_t stage_u = sWh_u + it.n_win * 128u;          // output staging 8 warps x [32 rows][9 words]
    const uint32_t total = n_chunks * n_groups;

    if (warp == BT_NT / 32 + 1) {
        // ================= source TMA thread =================
        if (elect_one()) {
            const CUtensorMap *tmap = tmaps + blockIdx.x;
            asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
            for (uint32_t ch = 0; ch < n_chunks; ch++) {
                const uint32_t slot = ch & 1, bar = smem_u32(&a_full[slot]);
                if (ch >= 2) mbar_wait_wd(smem_u32(&a_free[slot]), ((ch >> 1) - 1) & 1);  // the vertical MMAs of chunk ch - 2 have retired
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sA_u + slot * box_bytes),
                             "l"(tmap), "r"(bar), "r"(TC_M * ch), "r"(it.box_row0)
                             : "memory");
            }
        }
    } else if (warp == BT_NT / 32 + 2) {
        // ================= vertical-weight TMA thread =================
        if (elect_one()) {
            uint32_t have[2] = {0xffffffffu, 0xffffffffu};
            uint32_t g = 0;
            for (uint32_t gg = 0; gg < total; gg++) {
                const uint32_t slot = gg & 1, bar = smem_u32(&b_full[slot]);
                if (gg >= 2) mbar_wait_wd(smem_u32(&v_full[(gg - 2) % BT_NRV]), ((gg - 2) / BT_NRV) & 1);  // the slot's tile was read by the MMAs of group gg - 2
                const uint32_t kg = grp[4 * g + 1], b_off = grp[4 * g + 2];
                if (have[slot] == b_off) {
                    mbar_arrive(bar);  // the tile is there already: just complete the phase the MMA thread waits for
                } else {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kg * TC_N) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sB_u + slot * TC_N * kg_max),
                                 "l"(tb + b_off), "r"(kg * TC_N), "r"(bar)
                                 : "memory");
                    have[slot] = b_off;
                }
                if (++g == n_groups) g = 0;
            }
        }
    } else if (warp == BT_NT / 32) {
        // ================= vertical MMA thread =================
        if (elect_one()) {
            uint32_t g = 0, ch = 0;
            for (uint32_t gg = 0; gg < total; gg++) {
                const uint32_t region = gg % BT_NRV, ruse = gg / BT_NRV, bslot = gg & 1, aslot = ch & 1;
                const uint32_t a_off = grp[4 * g], kg = grp[4 * g + 1];
                mbar_wait_wd(smem_u32(&b_full[bslot]), (gg >> 1) & 1);
                if (g == 0) mbar_wait_wd(smem_u32(&a_full[aslot]), (ch >> 1) & 1);
                if (ruse > 0) mbar_wait_wd(smem_u32(&v_free[region]), (ruse - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint64_t da = umma_desc(sA_u + aslot * box_bytes + a_off * TC_M, 16, 1024, 2);  // the group's window: a_off rows into the box
                uint64_t db = umma_desc(sB_u + bslot * TC_N * kg_max, 128, (kg / 16) * 128);
                const uint32_t d_tmem = tmem_base + region * TC_N;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da),
                             "l"(db), "r"(UMMA_IDESC)
                             : "memory");
                for (uint32_t ks = 1; ks < kg / 32; ks++) {
                    da += (32 * TC_M) >> 4;
                    db += (2 * 128) >> 4;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                                 "l"(da), "l"(db), "r"(UMMA_IDESC)
                                 : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&v_full[region])) : "memory");
                if (++g == n_groups) {  // the chunk's box has been read: hand its slot back to the source thread
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&a_free[aslot])) : "memory");
                    g = 0;
                    ch++;
                }
            }
        }
    } else if (warp == BT_NT / 32 + 3) {
        // ================= horizontal thread =================
        if (elect_one()) {
            const uint32_t n_win = it.n_win, r_pad = it.r_pad, slack = it.slack;
            {
                const uint32_t bar = smem_u32(&wh_full[0]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n_win * 128u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sWh_u), "l"(tb + it.hw_off),
                             "r"(n_win * 128u), "r"(bar)
                             : "memory");
                mbar_wait_wd(bar, 0);
            }
            const uint32_t b_hi0 = sWh_u, b_lo0 = sWh_u + n_win * 64u;
            // one piece of a window: accumulator columns [col, col + n) += T[:, 32 j .. + 32) . W[brow .. brow + n)
            auto piece = [&](uint32_t a_hi, uint32_t col, uint32_t n, uint32_t brow) {
                const uint32_t idesc = (1u << 4) | (1u << 15) | ((n >> 3) << 17) | ((TC_M >> 4) << 24);  // f16 x f16 -> f32, A MN-major, B K-major
                const uint32_t d_tmem = tmem_base + BT_RING0 + col;
#pragma unroll
                for (int combo = 0; combo < 3; combo++) {
                    uint64_t da = umma_desc(combo == 1 ? a_hi + BT_T_BYTES : a_hi, 128, 2048);                              // LBO: next 8 columns (K), SBO: next 8 rows (M)
                    uint64_t db = umma_desc((combo == 2 ? b_lo0 : b_hi0) + (brow >> 3) * BT_SBO_W, 128, BT_SBO_W);         // LBO: next 8 columns (K), SBO: next 8 accumulator columns (N)
#pragma unroll
                    for (int ks = 0; ks < 2; ks++) {
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da),
                                     "l"(db), "r"(idesc)
                                     : "memory");
                        da += 256 >> 4;  // 16 columns = two core matrices along K
                        db += 256 >> 4;
                    }
                }
            };
            uint32_t jj = 0;
            for (uint32_t ch = 0; ch < n_chunks; ch++) {
                const uint32_t buf = ch & 1;
                mbar_wait_wd(smem_u32(&t_ready[buf]), (ch >> 1) & 1);  // the consumers have written the chunk's T
                for (uint32_t j = 0; j < 4; j++, jj++) {
                    if (jj >= slack) mbar_wait_wd(smem_u32(&d2_free[(jj - slack) & 3]), ((jj - slack) >> 2) & 1);  // the columns this window re-uses are drained and zero
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = sT_u + buf * 2 * BT_T_BYTES + j * 512u;  // 32 columns = four core matrices along K
                    const uint32_t s = (32u * jj + 4096u - r_pad) & 255u;
                    const uint32_t n1 = min(n_win, 256u - s);
                    piece(a_hi, s, n1, 0);
                    if (n1 < n_win) piece(a_hi, 0, n_win - n1, n1);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&d2_full[jj & 3])) : "memory");
                }
            }
        }
    } else {
        // ================= consumer warps =================
        const float scale = it.scale, scale_hi = it.scale * 16384.0f;
        const uint32_t q = warp & 3, half = warp >> 2, m = q * 32 + lane;
        const uint32_t r_pad = it.r_pad, n_e = it.n_e, band_rows = it.band_rows, dst_pitch = it.dst_pitch;
        const float *corr = tw + it.corr_off;
        const uint32_t row = q * 32 + lane;  // band row this thread drains from the ring
        uint8_t *const dst0 = it.dst + size_t(it.band_r0) * dst_pitch;  // first output byte of the band
#ifdef BT_PROF
        long long p_wv = 0, p_v = 0, p_wh = 0, p_h = 0, p_h1 = 0, p_h2 = 0;
        const long long p_start = clock64();
#define BT_T0 const long long t0_ = clock64()
#define BT_ACC(x) (x) += clock64() - t0_
#else
#define BT_T0
#define BT_ACC(x)
#endif
        // the ring starts at zero
#pragma unroll
        for (int k = 0; k < 8; k++) tmem_st16_zero(tmem_base + BT_RING0 + ((q * 32u) << 16) + half * 128 + 16 * k);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");

        // One finished sub-step of the horizontal pass: 32 columns of the ring -> 32 output bytes of every band row.  The
        // sub-steps alternate between the two warps that share a quarter of TMEM's lanes (warp >> 2 = parity of jj), so a
        // warp owns 32 rows x 32 bytes and needs no barrier with the others.  A thread holds ONE row (TMEM lanes are rows):
        // stored from registers, every store of a warp would touch 32 rows = 32 sectors (measured: 4 k clk per chunk for
        // 4-byte stores, 16 k for the byte stores of rows at odd addresses).  The bytes go through the warp's own staging
        // tile instead, [32 rows][8 words + pad], and leave as aligned words of a few rows per store: a word of a row
        // at an unaligned address is funnel-shifted from two staged words, the partial words at the two ends of the
        // segment leave as bytes.
        const uint32_t my_stage = stage_u + warp * BT_STAGE_WARP + 16u;  // 16 bytes in front: "word -1" of row 0 is readable
        const bool words_ok = ((reinterpret_cast<uintptr_t>(dst0) | dst_pitch) & 3) == 0;  // every row segment starts on a word
        const bool direct_ok = ((reinterpret_cast<uintptr_t>(dst0) | dst_pitch) & 7) == 0;  // every row starts on an 8-byte boundary (c0s is a multiple of 16)
        const bool halves_ok = ((reinterpret_cast<uintptr_t>(dst0) | dst_pitch | n_e) & 1) == 0;  // ... on a 2-byte boundary, and ends on one
        const uint32_t c_lo = it.c_lo, c_hi = it.c_hi;  // output bytes [c_lo, c_hi) have the whole window inside the row: factor 1 / 16
        auto drain_h = [&](uint32_t jj) {
            {
                BT_T0;
                mbar_wait(smem_u32(&d2_full[jj & 3]), (jj >> 2) & 1);
                BT_ACC(p_wh);
            }
            BT_T0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t f = (32u * jj + 4096u - r_pad) & 255u;  // a multiple of 16: each half of the 32 columns is contiguous in the ring
            const uint32_t tbase = tmem_base + BT_RING0 + ((q * 32u) << 16);
            uint32_t v[32];
            tmem_ld16(tbase + f, v);
            tmem_ld16(tbase + ((f + 16u) & 255u), v + 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tmem_st16_zero(tbase + f);
            tmem_st16_zero(tbase + ((f + 16u) & 255u));
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) mbar_arrive(smem_u32(&d2_free[jj & 3]));
#ifdef BT_PROF
            const long long t1_ = clock64();
            p_h1 += t1_ - t0_;
#endif
            const int c0s = int(32u * jj) - int(r_pad);   // first output byte of the sub-step (may be negative, or beyond the row)
            const int lo_b = c0s < 0 ? -c0s : 0, hi_b = min(32, int(n_e) - c0s);  // valid bytes of the segment
            if (hi_b <= lo_b) { BT_ACC(p_h); return; }
            uint32_t w8[8];
            if (c0s >= int(c_lo) && c0s + 32 <= int(c_hi)) {  // interior: no border factor
#pragma unroll
                for (int k = 0; k < 8; k++)
                    w8[k] = round_u8(__uint_as_float(v[4 * k]) * (1.0f / TC2_WSCALE)) | round_u8(__uint_as_float(v[4 * k + 1]) * (1.0f / TC2_WSCALE)) << 8 |
                            round_u8(__uint_as_float(v[4 * k + 2]) * (1.0f / TC2_WSCALE)) << 16 | round_u8(__uint_as_float(v[4 * k + 3]) * (1.0f / TC2_WSCALE)) << 24;
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c0s + 4 * k >= 0) c4 = __ldg(reinterpret_cast<const float4 *>(corr + c0s + 4 * k));  // (the table is padded past the row)
                    w8[k] = round_u8(__uint_as_float(v[4 * k]) * c4.x) | round_u8(__uint_as_float(v[4 * k + 1]) * c4.y) << 8 |
                            round_u8(__uint_as_float(v[4 * k + 2]) * c4.z) << 16 | round_u8(__uint_as_float(v[4 * k + 3]) * c4.w) << 24;
                }
            }
#ifndef BT_NO_DIRECT
            // whole, 8-byte aligned segments of rows on an 8-byte stride leave straight from the registers: a thread holds the
            // 32 bytes of ITS row, i.e. one whole sector -- two 16-byte or four 8-byte stores per thread instead of 8 staging
            // stores, 8 staged loads and 8 word stores per warp
            if (direct_ok && lo_b == 0 && hi_b == 32) {
                uint8_t *g = dst0 + size_t(row) * dst_pitch + c0s;
                if (row < band_rows) {
                    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
                        *reinterpret_cast<uint4 *>(g) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                        *reinterpret_cast<uint4 *>(g + 16) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++) *reinterpret_cast<uint2 *>(g + 8 * k) = make_uint2(w8[2 * k], w8[2 * k + 1]);
                    }
                }
                BT_ACC(p_h);
                return;
            }
#endif
            __syncwarp();  // the previous segment's words have left the staging tile
            const uint32_t sw = my_stage + lane * 36u;
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("st.shared.b32 [%0], %1;" ::"r"(sw + 4 * k), "r"(w8[k]) : "memory");
            __syncwarp();
#ifdef BT_PROF
            const long long t2_ = clock64();
            p_h2 += t2_ - t1_;
#endif
            uint8_t *const seg0 = dst0 + size_t(q * 32u) * dst_pitch + c0s;  // byte 0 of the segment in the warp's first row (not dereferenced outside the row)
            if (words_ok && lo_b == 0 && hi_b == 32) {
                // aligned rows: 8 lanes per row, four rows per store
                const uint32_t k = lane & 7;
#pragma unroll
                for (uint32_t i = 0; i < 8; i++) {
                    const uint32_t rr = 4 * i + (lane >> 3);
                    const uint32_t val = lds_u32(my_stage + rr * 36u + 4u * k);
                    if (q * 32u + rr < band_rows) *reinterpret_cast<uint32_t *>(seg0 + size_t(rr) * dst_pitch + 4 * k) = val;
                }
            } else if (halves_ok) {
                // rows at even addresses, even widths (L8 / RGB rows whose byte count is 2 mod 4): 2-byte stores, two rows per
                // instruction, every lane busy and no partial words.  The general path below took ~3.2 k clk per segment here
                // (~90 store instructions of a few lanes each: 70 % of the kernel's time on a 1618-byte row).
                const uint32_t hk = lane & 15;
#pragma unroll
                for (uint32_t i = 0; i < 16; i++) {
                    const uint32_t rr = 2 * i + (lane >> 4);
                    uint32_t val;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(val) : "r"(my_stage + rr * 36u + 2u * hk));
                    if (q * 32u + rr < band_rows && int(2 * hk) >= lo_b && int(2 * hk) < hi_b)
                        *reinterpret_cast<uint16_t *>(seg0 + size_t(rr) * dst_pitch + 2 * hk) = uint16_t(val);
                }
            } else {
                // rows at any address (odd widths, odd pitches).  Whole words first: the 32 x 9 (row, word) slots -- words 0..8 of the
                // segment as the row's address aligns them -- flattened over the lanes, three per lane with their staged words
                // fetched up front; then the partial words at the two ends of the segment, one row per lane.  (Nine lanes per
                // row and a store down one of four paths per word took ~3.2 k clk per segment.)
#pragma unroll 1
                for (uint32_t base = 0; base < 288; base += 96) {
                    uint32_t wa[3], wb[3], rr[3], kk[3];
#pragma unroll
                    for (uint32_t i = 0; i < 3; i++) {
                        const uint32_t item = base + 32 * i + lane;
                        rr[i] = item / 9;
                        kk[i] = item - 9 * rr[i];
                        const uint32_t s0 = my_stage + rr[i] * 36u + 4u * kk[i];
                        wa[i] = lds_u32(s0 - 4);  // staged words k - 1 and k (k = 8: the pad word, never a whole word of the segment)
                        wb[i] = lds_u32(s0);
                    }
#pragma unroll
                    for (uint32_t i = 0; i < 3; i++) {
                        uint8_t *a = seg0 + size_t(rr[i]) * dst_pitch;
                        const uint32_t ph = uint32_t(reinterpret_cast<uintptr_t>(a)) & 3u;
                        const int b0 = int(4 * kk[i]) - int(ph);  // segment byte of the word's byte 0
                        if (b0 >= lo_b && b0 + 4 <= hi_b && q * 32u + rr[i] < band_rows)
                            *reinterpret_cast<uint32_t *>(a + b0) = ph ? __funnelshift_r(wa[i], wb[i], 8 * (4 - ph)) : wb[i];
                    }
                }
                if (q * 32u + lane < band_rows) {  // the partial words of this lane's row: the one that holds byte lo_b, the one that holds byte hi_b - 1
                    uint8_t *a = seg0 + size_t(lane) * dst_pitch;
                    const uint32_t ph = uint32_t(reinterpret_cast<uintptr_t>(a)) & 3u;
                    const uint32_t k_h = (uint32_t(lo_b) + ph) >> 2, k_t = (uint32_t(hi_b) - 1 + ph) >> 2;
#pragma unroll
                    for (uint32_t e = 0; e < 2; e++) {
                        const uint32_t k = e ? k_t : k_h;
                        const int b0 = int(4 * k) - int(ph);
                        const int va = max(lo_b - b0, 0), vb = min(hi_b - b0, 4);  // valid bytes [va, vb) of the word
                        if ((e && k_t == k_h) || (va == 0 && vb == 4)) continue;
                        const uint32_t s0 = my_stage + lane * 36u + 4u * k;
                        const uint32_t w0 = lds_u32(s0 - 4), w1 = lds_u32(s0);
                        const uint32_t val = ph ? __funnelshift_r(w0, w1, 8 * (4 - ph)) : w1;
#pragma unroll
                        for (int bb = 0; bb < 4; bb++)
                            if (bb >= va && bb < vb) a[b0 + bb] = uint8_t(val >> (8 * bb));
                    }
                }
            }
            BT_ACC(p_h);
        };

        uint32_t gg = 0;
        for (uint32_t ch = 0; ch < n_chunks; ch++) {
            const uint32_t buf = ch & 1;
            for (uint32_t g = 0; g < n_groups; g++, gg++) {
                // ---- vertical results of group g -> T[buf] rows [32 g, + 32), this thread: column m, 16 rows
                const uint32_t region = gg % BT_NRV;
                {
                    BT_T0;
                    mbar_wait(smem_u32(&v_full[region]), (gg / BT_NRV) & 1);
                    BT_ACC(p_wv);
                }
#ifdef BT_PROF
                const long long tv_ = clock64();
#endif
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + region * TC_N + ((q * 32u) << 16) + half * 16;
                uint32_t hi[16], mid[16], lo[16];
                tmem_ld16(taddr, hi);
                tmem_ld16(taddr + 32, mid);
                tmem_ld16(taddr + 64, lo);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (lane == 0) mbar_arrive(smem_u32(&v_free[region]));
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int e = 0; e < 16; e += 2) {  // value = (hi 2^14 + mid 2^7 + lo) 2^-s, two rows per f32x2 op
                    const float2 fh = make_float2(float(int(hi[e])), float(int(hi[e + 1])));
                    const float2 fl = make_float2(float(int(mid[e]) * 128 + int(lo[e])), float(int(mid[e + 1]) * 128 + int(lo[e + 1])));
                    float2 r = make_float2(0.f, 0.f);
                    ffma2(r, fl, scale);
                    ffma2(r, fh, scale_hi);
                    const uint32_t h2 = pack_f16x2(r.x, r.y);
                    const float2 back = unpack_f16x2(h2);
                    ph[e / 2] = h2;
                    pl[e / 2] = pack_f16x2(r.x - back.x, r.y - back.y);
                }
                const uint32_t t0 = sT_u + buf * 2 * BT_T_BYTES + (g * 4 + half * 2) * 2048u + m * 16u;
                sts128(t0, ph[0], ph[1], ph[2], ph[3]);
                sts128(t0 + 2048, ph[4], ph[5], ph[6], ph[7]);
                sts128(t0 + BT_T_BYTES, pl[0], pl[1], pl[2], pl[3]);
                sts128(t0 + BT_T_BYTES + 2048, pl[4], pl[5], pl[6], pl[7]);
                if (g + 1 == n_groups) {  // the chunk's T is complete: hand it to the horizontal thread
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&t_ready[buf]));
                }
#ifdef BT_PROF
                p_v += clock64() - tv_;
#endif
                if (ch > 0 && (g & 1) == half) drain_h(4 * (ch - 1) + g);  // a finished sub-step of the previous chunk (this warp's parity)
            }
            if (ch > 0)
                for (uint32_t j = n_groups; j < 4; j++)
                    if ((j & 1) == half) drain_h(4 * (ch - 1) + j);
        }
        for (uint32_t j = 0; j < 4; j++)
            if ((j & 1) == half) drain_h(4 * (n_chunks - 1) + j);
#ifdef BT_PROF
        if (blockIdx.x == 200 && lane == 0 && (warp == 0 || warp == 5))
            printf("blur consumer warp %u: total %lld clk, %u chunks; wait vertical %lld, vertical drain %lld, wait horizontal %lld, horizontal drain %lld (TMEM part %lld, convert + stage %lld)\n",
                   warp, clock64() - p_start, n_chunks, p_wv, p_v, p_wh, p_h, p_h1, p_h2);
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

}  // namespace

int launch_blur_tc(const BlurTcItem *d_items, const void *d_tmaps, uint32_t n_items, size_t smem, const uint8_t *d_b, const uint32_t *d_info,
                   const float *d_w, LaunchCtx &lc) {
    if (n_items == 0) return 0;
    ensure_dynamic_smem(reinterpret_cast<const void *>(blur_tc2_kernel), smem);
    lc.begin("blur_tc2_kernel");
    blur_tc2_kernel<<<n_items, BT_NT_ALL, smem, lc.st>>>(d_items, static_cast<const CUtensorMap *>(d_tmaps), d_b, d_info, d_w);
    lc.end();
    return 1;
}

}  // namespace fanlin
