// Fused separable Lanczos3 resample, both passes on the sm_100a tensor cores, horizontal sums ACCUMULATED IN TMEM
// across chunks (fused_resample_tc3_kernel).  Same vertical pass and T tiles as fused_resample_tc2_kernel
// (kernels_fused_tc.cu: banded u8 x s8 contraction, three weight digits, f16 hi / lo operand tiles), but the horizontal
// contraction of a chunk adds into a RING of accumulator columns per row tile instead of producing partial sums the
// consumers keep in registers: output pixel o lives at columns ((o - ox0) mod RP) * c + channel, a chunk's MMAs cover
// the window of slots its 128 tile columns touch (two pieces where the window wraps), and the consumers read, round,
// store and zero only the slots whose taps ended in the chunk.  The ring has to hold what ONE chunk touches (not what
// fits a thread's registers), so downscales near 2 (BASELINE C1: 512 -> 200, C3: 3840 -> 1778 RGBA) take this kernel.
//
// Roles (1 CTA / SM, 16 warps; the kernel needs ~100 registers): source TMA thread, vertical-weight TMA thread, vertical MMA
// thread, horizontal thread (weight tiles -- the hi and the lo half fetched separately, each as early as its readers allow --
// and the MMAs, of which only the K steps with any weight for the window piece are issued), 12 consumer warps in three sets of
// one warp per TMEM lane quarter.  In chunk k set (k + t) mod 3 drains what chunk k - 1 finished in row tile t (TMEM -> rounded
// bytes -> the staging tile of (tile, lane quarter) -> aligned words of a few rows per store), and the tile's groups are
// converted to T by the other two sets (t3_owner): a tile goes back to the tensor core after max(drain, two conversions).
// Reference call sites: src/handler.rs:229-248 (resize / resize_to_fill + letterbox), image-0.25.6 imageops::resize.
#include <cuda.h>

#include "fused_device.cuh"
#include "fused_tc.h"
#include "kernels.h"
#include "tc_device.cuh"

namespace fanlin {

namespace {

constexpr int T3_NT = 384;            // consumer threads: 12 warps = 3 sets of one warp per TMEM lane quarter (the kernel needs ~100 registers)
constexpr int T3_NT_ALL = T3_NT + 128;
constexpr uint32_t T3_SETS = T3_NT / 128;
constexpr uint32_t T3_NB = 2;         // vertical weight-tile slots
constexpr uint32_t T3_NA_MAX = 4;
#ifndef T3_PF_AHEAD
#define T3_PF_AHEAD 4
#endif

#ifdef T3_PROF
#define T3W(acc, ...) do { const long long t0_ = clock64(); mbar_wait(__VA_ARGS__); (acc) += clock64() - t0_; } while (0)
#define T3WR(acc, ...) do { const long long t0_ = clock64(); mbar_wait_wd(__VA_ARGS__); (acc) += clock64() - t0_; } while (0)
#else
#define T3W(acc, ...) mbar_wait(__VA_ARGS__)
#define T3WR(acc, ...) mbar_wait_wd(__VA_ARGS__)  // role threads: with the watchdog
#endif

// Which of the three consumer sets converts group g (of 8 at most) of a chunk, cr = chunk mod 3.  In chunk k set cr drains
// what chunk k - 1 finished in row tile 0 and set cr + 1 what it finished in tile 1; a tile's groups go to the OTHER two sets,
// alternating, so that the tile is handed to the tensor core after max(drain, two conversions) and not after their sum.
__device__ __forceinline__ uint32_t t3_owner(uint32_t cr, uint32_t g) {
    const uint32_t b = cr == 2 ? 0u : cr + 1, c = cr == 0 ? 2u : cr - 1;
    return (g & 1) ? c : (g < 4 ? b : cr);
}
// v_full barrier of a group: one per (owner set, vertical region) -- only the owner waits on it, and on every phase of it (a
// set that skipped phases of a barrier shared with the other sets could not tell them apart by parity).  Waiters keep the
// parities in a bit mask.
__device__ __forceinline__ uint32_t t3_vbar(uint32_t owner, uint32_t region) { return owner * 4 + region; }

__device__ __forceinline__ void t3_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

template <int C, bool INV = false>  // INV: inverse on load, a template parameter (see fused_resample_tc2_kernel)
__global__ void __launch_bounds__(T3_NT_ALL, 1) fused_resample_tc3_kernel(const FusedTcItem *__restrict__ items, const CUtensorMap *__restrict__ tmaps,
                                                                          const uint8_t *__restrict__ tb, const uint32_t *__restrict__ tinfo) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ FusedTcItem it_s;
    // v_full: one barrier per (consumer set, vertical region), see t3_vbar
    __shared__ __align__(8) uint64_t v_full[12], v_free[4], a_full[T3_NA_MAX], b_full[T3_NB];
    __shared__ __align__(8) uint64_t t_ready[2], d2_full[2], d2_free[2], wh_full[2], wh_free[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t grp[4 * 8];
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        it_s = items[blockIdx.x];
        auto init = [](uint64_t *b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); };
        for (uint32_t r = 0; r < 12; r++) init(&v_full[r], 1);
        for (uint32_t r = 0; r < 4; r++) init(&v_free[r], 4);  // a group's vertical results are read by the four warps of ONE set
        for (uint32_t r = 0; r < T3_NA_MAX; r++) init(&a_full[r], 1);
        for (uint32_t r = 0; r < T3_NB; r++) init(&b_full[r], 1);
        for (uint32_t r = 0; r < 2; r++) {
            init(&t_ready[r], T3_NT / 32);
            init(&d2_full[r], 1);
            init(&d2_free[r], 4);  // what a chunk finished in a tile is drained by the four warps of one set
            init(&wh_full[r], 1);
            init(&wh_free[r], 1);
        }
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
    const FusedTcItem &it = it_s;
    const uint32_t tmem_base = tmem_base_s;
    if (warp < T3_NT / 32) fill_bars(it, warp, lane, T3_NT / 32);

    // ---- shared memory: T hi | T lo | source slots | vertical weight slots | horizontal weight slots | output staging per consumer warp
    const uint32_t kg_max = it.kg_max, n_a = it.n_a, n_groups = it.n_groups, n_chunks = it.n_chunks;
    const uint32_t t_bytes = n_groups * 32u * 256u;
    const uint32_t n_mt = (n_groups + 3) / 4, n_vr = it.n_vr, ring_cols = it.ring_cols;
    const uint32_t v0 = n_mt * ring_cols;  // first TMEM column of the vertical regions
    const uint32_t sT_u = smem_u32(smem);
    const uint32_t sA_u = sT_u + 2 * t_bytes;
    const uint32_t sB_u = sA_u + n_a * kg_max * TC_M;
    const uint32_t sWh_u = sB_u + T3_NB * TC_N * kg_max;
    const uint32_t n_wh = it.n_wh, wh_bytes = it.wh_bytes;
    const uint32_t stage_warp = (16u + 32u * it.stage_stride * 4u + 15u) & ~15u;
    const uint32_t stage_u = sWh_u + n_wh * wh_bytes;
    for (uint32_t k = tid; k < 4 * n_groups; k += T3_NT_ALL) grp[k] = tinfo[it.grp_off + k];
    __syncthreads();
    const uint32_t total = n_chunks * n_groups;
    const uint32_t *hrec = tinfo + it.hrec_off;

    if (warp == T3_NT / 32 + 1) {
        // ================= source TMA thread: one tensor copy per group, n_a groups deep =================
        if (elect_one()) {
            const CUtensorMap *tmap = tmaps + blockIdx.x;
            asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
            uint32_t g = 0, chunk = 0, slot = 0, pg = 0, pchunk = 0;
            uint32_t wg = 0, wcr = 0, wreg = 0, wph = 0;  // the group whose MMAs free the slot being refilled (group gg - n_a): index in its chunk, chunk mod 3, region; parities
            for (uint32_t k = 0; k < T3_PF_AHEAD && pchunk < n_chunks; k++)
                if (++pg == n_groups) { pg = 0; pchunk++; }
            for (uint32_t gg = 0; gg < total; gg++) {
                if (pchunk < n_chunks) {
                    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(it.b0 + TC_M * pchunk), "r"(grp[4 * pg]) : "memory");
                    if (++pg == n_groups) { pg = 0; pchunk++; }
                }
                if (gg >= n_a) {
                    const uint32_t vb = t3_vbar(t3_owner(wcr, wg), wreg);
                    mbar_wait_wd(smem_u32(&v_full[vb]), (wph >> vb) & 1);
                    wph ^= 1u << vb;
                    if (++wreg == n_vr) wreg = 0;
                    if (++wg == n_groups) { wg = 0; if (++wcr == 3) wcr = 0; }
                }
                const uint32_t bar = smem_u32(&a_full[slot]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kg_max * TC_M) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sA_u + slot * kg_max * TC_M),
                             "l"(tmap), "r"(bar), "r"(it.b0 + TC_M * chunk), "r"(grp[4 * g])
                             : "memory");
                if (++slot == n_a) slot = 0;
                if (++g == n_groups) { g = 0; chunk++; }
            }
        }
    } else if (warp == T3_NT / 32 + 2) {
        // ================= vertical-weight TMA thread =================
        if (elect_one()) {
            uint32_t g = 0, wg = 0, wcr = 0, wreg = 0, wph = 0;
            for (uint32_t gg = 0; gg < total; gg++) {
                if (gg >= T3_NB) {
                    const uint32_t vb = t3_vbar(t3_owner(wcr, wg), wreg);
                    mbar_wait_wd(smem_u32(&v_full[vb]), (wph >> vb) & 1);  // the slot's previous tile was read by the MMAs of group gg - 2
                    wph ^= 1u << vb;
                    if (++wreg == n_vr) wreg = 0;
                    if (++wg == n_groups) { wg = 0; if (++wcr == 3) wcr = 0; }
                }
                const uint32_t kg = grp[4 * g + 1], b_off = grp[4 * g + 2];
                const uint32_t bar = smem_u32(&b_full[gg & 1]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kg * TC_N) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sB_u + (gg & 1) * TC_N * kg_max),
                             "l"(tb + b_off), "r"(kg * TC_N), "r"(bar)
                             : "memory");
                if (++g == n_groups) g = 0;
            }
        }
    } else if (warp == T3_NT / 32) {
        // ================= vertical MMA thread =================
        if (elect_one()) {
            uint32_t g = 0, cr = 0, slot = 0, suse = 0, region = 0, ruse = 0;
            long long w_b = 0, w_a = 0, w_r = 0;
            const long long t_start = clock64();
            for (uint32_t gg = 0; gg < total; gg++) {
                const uint32_t bslot = gg & 1, kg = grp[4 * g + 1];
                T3WR(w_b, smem_u32(&b_full[bslot]), (gg >> 1) & 1);
                T3WR(w_a, smem_u32(&a_full[slot]), suse & 1);
                if (ruse > 0) T3WR(w_r, smem_u32(&v_free[region]), (ruse - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint64_t da = umma_desc(sA_u + slot * kg_max * TC_M, 16, 1024, 2);
                uint64_t db = umma_desc(sB_u + bslot * TC_N * kg_max, 128, (kg / 16) * 128);
                const uint32_t d_tmem = tmem_base + v0 + region * TC_N;
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
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&v_full[t3_vbar(t3_owner(cr, g), region)])) : "memory");
                if (++slot == n_a) { slot = 0; suse++; }
                if (++region == n_vr) { region = 0; ruse++; }
                if (++g == n_groups) { g = 0; if (++cr == 3) cr = 0; }
            }
#ifdef T3_PROF
            if (blockIdx.x == 300) printf("tc3 vertical MMA thread: total %lld clk, %u chunks x %u groups; waits: weights %lld, source rows %lld, TMEM region %lld\n", clock64() - t_start, n_chunks, n_groups, w_b, w_a, w_r);
#else
            (void)t_start; (void)w_b; (void)w_a; (void)w_r;
#endif
        }
    } else if (warp == T3_NT / 32 + 3) {
        // ================= horizontal thread: weight tiles + MMAs into the rings =================
        if (elect_one()) {
            auto load_wh = [&](uint32_t ch) {
                const uint32_t slot = n_wh == 2 ? (ch & 1) : 0u, bar = smem_u32(&wh_full[slot]);
                const uint32_t bytes = __ldg(hrec + 8 * ch + 4) * 512u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sWh_u + slot * wh_bytes),
                             "l"(tb + __ldg(hrec + 8 * ch)), "r"(bytes), "r"(bar)
                             : "memory");
            };
            long long w_wh = 0, w_tr = 0, w_df = 0, w_wf = 0;
            const long long t_start = clock64();
            // window piece: accumulator columns [col, col + n) += T . W[brow .. brow + n), combos [c_a, c_b) of hi x hi, lo x hi, hi x lo
            // K steps [k_lo, k_hi) only: the others hold no weight for this piece (host record word 6)
            auto piece = [&](uint32_t mt, uint32_t col, uint32_t n, uint32_t brow, uint32_t b_hi0, uint32_t b_lo0, int c_a, int c_b, uint32_t k_lo, uint32_t k_hi) {
                const uint32_t idesc = (1u << 4) | (1u << 15) | ((n >> 3) << 17) | ((TC_M >> 4) << 24);  // f16 x f16 -> f32, A MN-major, B K-major
                const uint32_t d_tmem = tmem_base + mt * ring_cols + col;
                const uint32_t a_hi = sT_u + mt * 32768u, a_lo = a_hi + t_bytes;
#pragma unroll
                for (int combo = 0; combo < 3; combo++) {
                    if (combo < c_a || combo >= c_b) continue;
                    // (a loop with a uniform trip count: a predicate on the instruction itself broke the back-to-back issue -- C3 + 4 %)
                    uint64_t da = umma_desc(combo == 1 ? a_lo : a_hi, 128, 2048) + k_lo * (256 >> 4);
                    uint64_t db = umma_desc((combo == 2 ? b_lo0 : b_hi0) + (brow >> 3) * 2048u, 128, 2048) + k_lo * (256 >> 4);
#pragma unroll 1
                    for (uint32_t ks = k_lo; ks < k_hi; ks++) {
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                                     "l"(da), "l"(db), "r"(idesc)
                                     : "memory");
                        da += 256 >> 4;
                        db += 256 >> 4;
                    }
                }
            };
            auto commit = [](uint64_t *bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); };
            if (n_wh == 1) {
                // ONE slot for the chunk's weight tiles (they are up to 48 KB), but its hi and lo halves have lives of their own:
                // the hi tile is read by the hi x hi and lo x hi series, the lo tile by hi x lo only.  The last row tile's series are
                // issued hi x hi, lo x hi, hi x lo; when the first two have retired the next chunk's hi tile is fetched under the
                // hi x lo series, and its lo tile under the next chunk's first series -- with the whole pair fetched only once a
                // chunk's MMAs had retired, the copy (~1.2 k clk per chunk) was exposed: 134 k of a C3 CTA's 915 k clk.
                auto load_half = [&](uint32_t ch, uint32_t half) {  // wh_full[0 / 1], wh_free[0 / 1]: hi / lo half
                    const uint32_t bar = smem_u32(&wh_full[half]), n_total = __ldg(hrec + 8 * ch + 4), bytes = n_total * 256u;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                    // (fixed places for the halves: the tiles of consecutive chunks differ in size, and the next hi tile lands while this lo tile is read)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sWh_u + half * (wh_bytes >> 1)),
                                 "l"(tb + __ldg(hrec + 8 * ch) + half * bytes), "r"(bytes), "r"(bar)
                                 : "memory");
                };
                load_half(0, 0);
                load_half(0, 1);
                for (uint32_t ch = 0; ch < n_chunks; ch++) {
                    const uint32_t w0 = __ldg(hrec + 8 * ch + 3), n_total = __ldg(hrec + 8 * ch + 4), kr = __ldg(hrec + 8 * ch + 6);
                    const uint32_t n1 = min(n_total, ring_cols - w0);
                    const uint32_t b_hi0 = sWh_u, b_lo0 = b_hi0 + (wh_bytes >> 1);
                    for (uint32_t mt = 0; mt < n_mt; mt++) {
                        const bool last = mt + 1 == n_mt;
                        T3WR(w_tr, smem_u32(&t_ready[mt]), ch & 1);                     // the consumers have written the tile's rows
                        if (ch > 0) T3WR(w_df, smem_u32(&d2_free[mt]), (ch - 1) & 1);   // ... and drained and zeroed what the previous chunk finished
                        if (mt == 0) T3WR(w_wh, smem_u32(&wh_full[0]), ch & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        piece(mt, w0, n1, 0, b_hi0, b_lo0, 0, 2, kr & 255u, (kr >> 8) & 255u);
                        if (n1 < n_total) piece(mt, 0, n_total - n1, n1, b_hi0, b_lo0, 0, 2, (kr >> 16) & 255u, kr >> 24);
                        if (last) commit(&wh_free[0]);
                        if (mt == 0) { T3WR(w_wh, smem_u32(&wh_full[1]), ch & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
                        piece(mt, w0, n1, 0, b_hi0, b_lo0, 2, 3, kr & 255u, (kr >> 8) & 255u);
                        if (n1 < n_total) piece(mt, 0, n_total - n1, n1, b_hi0, b_lo0, 2, 3, (kr >> 16) & 255u, kr >> 24);
                        commit(&d2_full[mt]);
                        if (last) commit(&wh_free[1]);
                    }
                    if (ch + 1 < n_chunks) {
                        T3WR(w_wf, smem_u32(&wh_free[0]), ch & 1);  // hi x hi and lo x hi of the last tile have retired (hi x lo is still running)
                        load_half(ch + 1, 0);
                        T3WR(w_wf, smem_u32(&wh_free[1]), ch & 1);
                        load_half(ch + 1, 1);
                    }
                }
            } else {
            load_wh(0);
            for (uint32_t ch = 0; ch < n_chunks; ch++) {
                const uint32_t slot = ch & 1;
                if (ch + 1 < n_chunks) {  // the other slot: free once the horizontal MMAs of chunk ch - 1 have retired
                    if (ch >= 1) T3WR(w_wf, smem_u32(&wh_free[(ch + 1) & 1]), ((ch - 1) >> 1) & 1);
                    load_wh(ch + 1);
                }
                T3WR(w_wh, smem_u32(&wh_full[slot]), (ch >> 1) & 1);
                const uint32_t w0 = __ldg(hrec + 8 * ch + 3), n_total = __ldg(hrec + 8 * ch + 4), kr = __ldg(hrec + 8 * ch + 6);
                const uint32_t n1 = min(n_total, ring_cols - w0);
                const uint32_t b_hi0 = sWh_u + slot * wh_bytes, b_lo0 = b_hi0 + n_total * 256u;
                for (uint32_t mt = 0; mt < n_mt; mt++) {
                    T3WR(w_tr, smem_u32(&t_ready[mt]), ch & 1);                     // the consumers have written the tile's rows
                    if (ch > 0) T3WR(w_df, smem_u32(&d2_free[mt]), (ch - 1) & 1);   // ... and drained and zeroed what the previous chunk finished
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    piece(mt, w0, n1, 0, b_hi0, b_lo0, 0, 3, kr & 255u, (kr >> 8) & 255u);
                    if (n1 < n_total) piece(mt, 0, n_total - n1, n1, b_hi0, b_lo0, 0, 3, (kr >> 16) & 255u, kr >> 24);
                    commit(&d2_full[mt]);
                    if (mt + 1 == n_mt) commit(&wh_free[slot]);
                }
            }
            }
#ifdef T3_PROF
            if (blockIdx.x == 300) printf("tc3 horizontal thread: total %lld clk, n_wh %u, n_a %u, n_vr %u, ring %u; waits: weight tile landed %lld, T tile %lld, ring drained %lld, weight slot free %lld\n", clock64() - t_start, n_wh, n_a, n_vr, ring_cols, w_wh, w_tr, w_df, w_wf);
#else
            (void)t_start; (void)w_wh; (void)w_tr; (void)w_df; (void)w_wf;
#endif
        }
    } else {
        // ================= consumer warps =================
        const float scale = it.scale, scale_hi = it.scale * 16384.0f;
        // Work split (t3_owner): what chunk k - 1 finished in row tile t is drained by the four warps of set (k + t) mod 3, a
        // tile's groups (both halves of their 32 rows) are converted by the other two sets.  (Eight warps -- two per scheduler -- left the consumers latency-bound at ~0.35 IPC each and
        // the tensor pipe waiting for them; the kernel needs ~100 registers, so twelve fit.)
        const uint32_t q = warp & 3, set = warp >> 2, m = q * 32 + lane;
        // inverse on load applies to the colour channels of this thread's tile column (byte b0 + 128 chunk + m of the row)
        const bool inv_on = INV && !((C == 2 || C == 4) && ((it.b0 + m) % C) == C - 1);
        const uint32_t grp_rows = it.grp_rows;
        const uint32_t h_cout = it.c_out, h_pitch = it.dst_pitch, h_rows = it.band_rows, h_epi = it.epi, h_fill = it.fill;
        const uint32_t RP = ring_cols / C;
        uint8_t *const h_row0 = it.dst + size_t(it.dst_y + it.band_r0) * it.dst_pitch + size_t(it.dst_x) * h_cout;  // first canvas byte of the band
        const uint32_t stride4 = it.stage_stride * 4u;
        // the rings start at zero: warp (q, set) clears its lanes of the tiles' columns, a third of them each
        for (uint32_t c0 = set * 16; c0 < v0; c0 += 16 * T3_SETS) tmem_st16_zero(tmem_base + ((q * 32u) << 16) + c0);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        const bool words_ok = h_cout == 4 && ((reinterpret_cast<uintptr_t>(h_row0) | h_pitch) & 3) == 0;  // every row segment starts on a word

        // the pixels chunk `chunk` finished in row tile `mt`: ring -> registers -> rounded bytes -> staging -> canvas
        long long w_v = 0, w_d2 = 0, t_dr = 0, t_v = 0;
        long long t_ld = 0, t_zs = 0, t_wo = 0, t_cv = 0, t_z = 0, t_sg = 0;  // T3_PROF: inside the ring drain -- TMEM loads, the wait for the zeroing stores, write-out
        const long long t_start = clock64();
        // (fin_first, n_fin, fin_slot: the chunk's record words 1, 2, 5 -- fetched a chunk ahead by the caller: a load from
        // global memory inside the drain cost ~0.8 k of its ~3.2 k clk)
        auto drain_ring = [&](uint32_t chunk, uint32_t mt, uint32_t drainer, uint32_t fin_first, uint32_t n_fin, uint32_t fin_slot) {
            T3W(w_d2, smem_u32(&d2_full[mt]), chunk & 1);  // every warp: the tile's T rows may be overwritten from here on
            if (drainer != set) return;  // ... the sets take turns at draining (a drain is latency-bound: splitting one does not shorten it)
            // staging tile of (row tile, lane quarter): its previous user -- another set, a chunk ago -- finished its write-out
            // before it arrived at t_ready for the chunk whose horizontal MMAs this drain waited for; 16 bytes in front: "word -1" of row 0 is readable
            const uint32_t my_stage = stage_u + (mt * 4 + q) * stage_warp + 16u;
#ifdef T3_PROF
            const long long td0 = clock64();
            struct Acc { long long &a; long long t0; __device__ ~Acc() { a += clock64() - t0; } } acc_{t_dr, td0};
#endif
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            constexpr uint32_t PB = 16 / C;
            const uint32_t p0 = 0, p1 = n_fin;
            const uint32_t tbase = tmem_base + mt * ring_cols + ((q * 32u) << 16);
            const uint32_t sw = my_stage + lane * stride4;
            __syncwarp();  // the previous chunk's words have left the staging tile
            // Batches of PB = 16 / C pixels: ONE tcgen05.ld of 16 columns (what it reads past the batch is ignored) and one
            // to four stores that zero exactly the batch's columns, a batch ending where the ring wraps.  Per-pixel loads and stores cost an issue slot each next to a tensor core
            // that keeps the shared-memory and TMEM ports busy (measured: 6.4 k clk per chunk for 15 RGBA pixels).
            uint32_t slot = fin_slot + p0;
            if (slot >= RP) slot -= RP;
            for (uint32_t pb = p0; pb < p1;) {
                const uint32_t np = min(min(PB, p1 - pb), RP - slot);
                const uint32_t ta = tbase + slot * C;
                uint32_t v[16];
#ifdef T3_PROF
                const long long tl0 = clock64();
#endif
                tmem_ld16(ta, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#ifdef T3_PROF
                t_ld += clock64() - tl0;
#endif
                uint32_t px[PB];
#pragma unroll
                for (uint32_t i = 0; i < PB; i++) {
                    uint32_t u[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        u[c] = round_u8(__uint_as_float(v[i * C + c]) * (1.0f / TC2_WSCALE));
                    }
                    if (h_epi == EPI_PLAIN || (h_epi & EPI_GRAY)) {  // (gray canvas: C = 1, the luma byte)
                        px[i] = u[0] | u[1] << 8 | u[2] << 16 | u[3] << 24;  // the pixel's c bytes, low byte first
                    } else {
                        px[i] = to_rgba_packed(u, C);
                        if ((h_epi & EPI_MASK) == EPI_BLEND_FILL) px[i] = blend_rgba(h_fill, px[i]);
                    }
                }
#ifdef T3_PROF
                const long long tc1 = clock64();
                t_cv += tc1 - tl0;
#endif
                // zero exactly the columns of the np pixels read: the neighbouring columns belong to live outputs, or to the
                // pixels the other warp of this lane quarter drains
                {
                    const uint32_t z = 0;
                    if (np == PB) {
                        if constexpr (PB * C == 16) {
                            tmem_st16_zero(ta);
                        } else {  // C == 3: 15 columns = 8 + 4 + 2 + 1
                            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(ta), "r"(z) : "memory");
                            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%1,%1,%1};" ::"r"(ta + 8), "r"(z) : "memory");
                            asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%1};" ::"r"(ta + 12), "r"(z) : "memory");
                            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(ta + 14), "r"(z) : "memory");
                        }
                    } else {
                        for (uint32_t i = 0; i < np; i++) {
                            const uint32_t tp = ta + i * C;
                            if constexpr (C == 4) asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%1,%1,%1};" ::"r"(tp), "r"(z) : "memory");
                            else if constexpr (C == 3) {
                                asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%1};" ::"r"(tp), "r"(z) : "memory");
                                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tp + 2), "r"(z) : "memory");
                            } else if constexpr (C == 2) asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%1};" ::"r"(tp), "r"(z) : "memory");
                            else asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tp), "r"(z) : "memory");
                        }
                    }
                }
#ifdef T3_PROF
                const long long tc2 = clock64();
                t_z += tc2 - tc1;
#endif
                const uint32_t sp = sw + (pb - p0) * h_cout;  // staged at its byte offset within this warp's segment
                if (h_cout == 4) {
                    if constexpr (PB == 4) {
                        if (np == 4 && ((pb - p0) & 3) == 0) {
                            sts128(sp, px[0], px[1], px[2], px[3]);
                        } else {
#pragma unroll
                            for (uint32_t i = 0; i < PB; i++)
                                if (i < np) asm volatile("st.shared.b32 [%0], %1;" ::"r"(sp + 4 * i), "r"(px[i]) : "memory");
                        }
                    } else {
#pragma unroll
                        for (uint32_t i = 0; i < PB; i++)
                            if (i < np) asm volatile("st.shared.b32 [%0], %1;" ::"r"(sp + 4 * i), "r"(px[i]) : "memory");
                    }
                } else {  // plain output of 1..3 channels, or RGB8 (EPI_RGB8: the packed pixel without its alpha byte): byte by byte
#pragma unroll
                    for (uint32_t i = 0; i < PB; i++)
                        if (i < np) {
#pragma unroll
                            for (uint32_t c = 0; c < 3; c++)
                                if (c < h_cout) sts8(sp + i * h_cout + c, px[i] >> (8 * c));
                        }
                }
#ifdef T3_PROF
                t_sg += clock64() - tc2;
#endif
                pb += np;
                slot += np;
                if (slot >= RP) slot -= RP;
            }
#ifdef T3_PROF
            const long long tz0 = clock64();
#endif
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) t3_arrive(smem_u32(&d2_free[mt]));
#ifdef T3_PROF
            t_zs += clock64() - tz0;
            const long long tw0 = clock64();
            struct AccW { long long &a; long long t0; __device__ ~AccW() { a += clock64() - t0; } } accw_{t_wo, tw0};
#endif
            if (p1 <= p0) return;
            __syncwarp();
            // ---- write out: this group's rows x nb bytes at canvas byte (fin_first + p0) * c_out of each row
            const uint32_t nb = (p1 - p0) * h_cout;
            const uint32_t gq = mt * 4 + q;
            uint8_t *const seg0 = h_row0 + size_t(gq) * grp_rows * h_pitch + size_t(fin_first + p0) * h_cout;  // row = group * grp_rows + lane
            const uint32_t rows_here = gq < n_groups ? grp[4 * gq + 3] : 0u;
            if (words_ok) {
                // every row segment starts on a word: nw words per row, lpr lanes per row (a power of two), 32 / lpr rows per store,
                // four stores in flight
                const uint32_t nw = nb >> 2;
                uint32_t sh = 0;
                while ((1u << sh) < nw && sh < 5) sh++;
                const uint32_t lpr = 1u << sh, rpi = 32u >> sh, k0 = lane & (lpr - 1), r0 = lane >> sh;
                for (uint32_t k = k0; k < nw; k += lpr) {
                    for (uint32_t rb = r0; rb < rows_here; rb += 4 * rpi) {
                        uint32_t val[4];
#pragma unroll
                        for (uint32_t i = 0; i < 4; i++) val[i] = lds_u32(my_stage + min(rb + i * rpi, 31u) * stride4 + 4u * k);
#pragma unroll
                        for (uint32_t i = 0; i < 4; i++)
                            if (rb + i * rpi < rows_here) *reinterpret_cast<uint32_t *>(seg0 + size_t(rb + i * rpi) * h_pitch + 4u * k) = val[i];
                    }
                }
            } else {
                // Row segments at any byte address (three-byte pixels, L8 / LA, odd pitches).  Whole words first: the (row, word)
                // pairs flattened over the lanes, four per lane with their staged words fetched up front; then the partial
                // words at the two ends of every segment, one row per lane.  (One row per iteration -- 21 words of a 25-pixel
                // RGB segment on 32 lanes, a chain of load, shift and a store down one of four paths -- took 11 k clk per
                // drain: a C1-sized crop request ran four times slower than the letterboxed one.)
                const uint32_t nw_max = (nb + 6) >> 2;  // words a segment can span, whatever its alignment
                const uint32_t items = rows_here * nw_max;
                const float inv = 1.0f / float(nw_max);
                for (uint32_t base = 0; base < items; base += 128) {
                    uint32_t wa[4], wb[4], rr[4], kk[4];
#pragma unroll
                    for (uint32_t i = 0; i < 4; i++) {
                        const uint32_t item = min(base + 32 * i + lane, items - 1);
                        rr[i] = uint32_t((float(item) + 0.5f) * inv);
                        kk[i] = item - rr[i] * nw_max;
                        const uint32_t s0 = my_stage + rr[i] * stride4 + 4u * kk[i];
                        wa[i] = lds_u32(s0 - 4);
                        wb[i] = lds_u32(s0);
                    }
#pragma unroll
                    for (uint32_t i = 0; i < 4; i++) {
                        uint8_t *a = seg0 + size_t(rr[i]) * h_pitch;
                        const uint32_t ph = uint32_t(reinterpret_cast<uintptr_t>(a)) & 3u;
                        const int b0 = int(4 * kk[i]) - int(ph);  // segment byte of the word's byte 0
                        if (base + 32 * i + lane < items && b0 >= 0 && b0 + 4 <= int(nb))
                            *reinterpret_cast<uint32_t *>(a + b0) = ph ? __funnelshift_r(wa[i], wb[i], 8 * (4 - ph)) : wb[i];
                    }
                }
                if (lane < rows_here) {  // the partial words: the head (ph bytes missing in front) and the tail of this lane's row
                    uint8_t *a = seg0 + size_t(lane) * h_pitch;
                    const uint32_t ph = uint32_t(reinterpret_cast<uintptr_t>(a)) & 3u;
                    const uint32_t nw = (ph + nb + 3) >> 2;
                    const uint32_t s0 = my_stage + lane * stride4;
#pragma unroll
                    for (uint32_t e = 0; e < 2; e++) {
                        const uint32_t k = e ? nw - 1 : 0u;
                        const int b0 = int(4 * k) - int(ph);
                        const int va = max(-b0, 0), vb = min(int(nb) - b0, 4);
                        if ((e && nw < 2) || (va == 0 && vb == 4)) continue;  // (a one-word segment is the head; whole words went above)
                        const uint32_t w0 = lds_u32(s0 + 4u * k - 4), w1 = lds_u32(s0 + 4u * k);
                        const uint32_t val = ph ? __funnelshift_r(w0, w1, 8 * (4 - ph)) : w1;
#pragma unroll
                        for (int bb = 0; bb < 4; bb++)
                            if (bb >= va && bb < vb) a[b0 + bb] = uint8_t(val >> (8 * bb));
                    }
                }
            }
        };

        uint32_t gg = 0, region = 0, vph = 0, cr = 0;  // vph: parities of this set's v_full barriers; cr = chunk mod 3
        uint32_t rc0 = 0, rc1 = 0, rc2 = 0, rn0 = 0, rn1 = 0, rn2 = 0;  // record of the chunk being drained (chunk - 1) / of this chunk
        for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
            rc0 = rn0; rc1 = rn1; rc2 = rn2;
            rn0 = __ldg(hrec + 8 * chunk + 1); rn1 = __ldg(hrec + 8 * chunk + 2); rn2 = __ldg(hrec + 8 * chunk + 5);
            for (uint32_t g = 0; g < n_groups; g++, gg++) {
                // the tile's rows are about to be overwritten: its horizontal MMAs of the previous chunk must have retired --
                // which is also when the pixels they finished can be drained
                if ((g & 3) == 0 && chunk > 0) drain_ring(chunk - 1, g >> 2, (g >> 2) ? (cr == 2 ? 0u : cr + 1) : cr, rc0, rc1, rc2);
                const bool mine = t3_owner(cr, g) == set;
                if (mine) {
                const uint32_t vb = t3_vbar(set, region);
                T3W(w_v, smem_u32(&v_full[vb]), (vph >> vb) & 1);  // the vertical MMAs of group gg have retired
                vph ^= 1u << vb;
#ifdef T3_PROF
                const long long tv0 = clock64();
#endif
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (uint32_t half = 0; half < 2; half++) {  // rows 0-15, 16-31 of the group
                const uint32_t taddr = tmem_base + v0 + region * TC_N + ((q * 32u) << 16) + half * 16;
                uint32_t hi[16], mid[16], lo[16];
                tmem_ld16(taddr, hi);
                tmem_ld16(taddr + 32, mid);
                tmem_ld16(taddr + 64, lo);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (half == 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (lane == 0) t3_arrive(smem_u32(&v_free[region]));
                }
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int e = 0; e < 16; e += 2) {  // value = (hi 2^14 + mid 2^7 + lo) 2^-s, two rows per f32x2 op
                    const float2 fh = make_float2(float(int(hi[e])), float(int(hi[e + 1])));
                    const float2 fl = make_float2(float(int(mid[e]) * 128 + int(lo[e])), float(int(mid[e + 1]) * 128 + int(lo[e + 1])));
                    float2 r = make_float2(0.f, 0.f);
                    ffma2(r, fl, scale);
                    ffma2(r, fh, scale_hi);
                    if (INV && inv_on) {  // inverse on load: 255 sum q - sum q x (colour channels; alpha columns pass)
                        const uint32_t ri = g * grp_rows + half * 16 + e, rmax = it.band_rows - 1;  // (rows past the band are never stored)
                        r.x = __uint_as_float(__ldg(tinfo + it.inv_off + min(ri, rmax))) - r.x;
                        r.y = __uint_as_float(__ldg(tinfo + it.inv_off + min(ri + 1, rmax))) - r.y;
                    }
                    const uint32_t h2 = pack_f16x2(r.x, r.y);
                    const float2 back = unpack_f16x2(h2);
                    ph[e / 2] = h2;
                    pl[e / 2] = pack_f16x2(r.x - back.x, r.y - back.y);
                }
                const uint32_t t0 = sT_u + (g * 4 + half * 2) * 2048u + m * 16u;
                sts128(t0, ph[0], ph[1], ph[2], ph[3]);
                sts128(t0 + 2048, ph[4], ph[5], ph[6], ph[7]);
                sts128(t0 + t_bytes, pl[0], pl[1], pl[2], pl[3]);
                sts128(t0 + t_bytes + 2048, pl[4], pl[5], pl[6], pl[7]);
                }
#ifdef T3_PROF
                t_v += clock64() - tv0;
#endif
                }
                if (++region == n_vr) region = 0;
                if ((g & 3) == 3 || g + 1 == n_groups) {  // the row tile is complete: hand it to the tensor core
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) t3_arrive(smem_u32(&t_ready[g >> 2]));
                }
            }
            if (++cr == 3) cr = 0;
        }
        drain_ring(n_chunks - 1, 0, cr, rn0, rn1, rn2);  // (cr = n_chunks mod 3 here)
        if (n_mt > 1) drain_ring(n_chunks - 1, 1, cr == 2 ? 0u : cr + 1, rn0, rn1, rn2);
#ifdef T3_PROF
        if (blockIdx.x == 300 && lane == 0 && (warp == 0 || warp == 5 || warp == 10))
            printf("tc3 consumer warp %u: total %lld clk; waits: vertical results %lld, ring ready %lld; vertical drain %lld, ring drain + write-out %lld (TMEM loads %lld, load + convert %lld, zeroing %lld, staging %lld, wait for zeroing stores %lld, write-out %lld)\n", warp, clock64() - t_start, w_v, w_d2, t_v, t_dr, t_ld, t_cv, t_z, t_sg, t_zs, t_wo);
#else
        (void)t_start; (void)w_v; (void)w_d2; (void)t_dr; (void)t_v; (void)t_ld; (void)t_zs; (void)t_wo; (void)t_cv; (void)t_z; (void)t_sg;
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

template <int C, bool INV>
void launch_tc3_variant(const FusedTcItem *d_items, const void *d_tmaps, uint32_t n_items, size_t smem, const uint8_t *d_b, const uint32_t *d_info,
                        LaunchCtx &lc) {
    auto kern = fused_resample_tc3_kernel<C, INV>;
    ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
    lc.begin("fused_resample_tc3_kernel");
    kern<<<n_items, T3_NT_ALL, smem, lc.st>>>(d_items, static_cast<const CUtensorMap *>(d_tmaps), d_b, d_info);
    lc.end();
}

}  // namespace

int launch_fused_tc3(const FusedTcItem *d_items, const void *d_tmaps, uint32_t n_items, uint32_t c, size_t smem, const uint8_t *d_b,
                     const uint32_t *d_info, LaunchCtx &lc) {
    if (n_items == 0) return 0;
    switch (c) {  // bit 5: inverse on load
    case 1: launch_tc3_variant<1, false>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
    case 2: launch_tc3_variant<2, false>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
    case 3: launch_tc3_variant<3, false>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
    case 4: launch_tc3_variant<4, false>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
    case 33: launch_tc3_variant<1, true>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
    case 34: launch_tc3_variant<2, true>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
    case 35: launch_tc3_variant<3, true>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
    case 36: launch_tc3_variant<4, true>(d_items, d_tmaps, n_items, smem, d_b, d_info, lc); return 1;
    }
    return -1;
}

}  // namespace fanlin
