// Host side of the tensor-core blur (kernels_blur_tc.cu): bands of 128 rows, their 32-row groups and
// vertical weight-digit tiles (deduplicated: every interior group shares one), the Toeplitz tile of the
// horizontal pass and the per-column border factors.
#include "blur.h"
#include "fused_tc.h"

#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>

namespace fanlin {

namespace {
constexpr uint32_t BT_ROWS = 128;   // band rows = M of the horizontal MMAs
constexpr uint32_t BT_KSUB = 32;    // bytes of T per horizontal sub-step
constexpr uint32_t BT_RING = 256;   // TMEM columns of the output ring

struct BandGeom {
    uint32_t r0, rows, grp_off, n_groups, box_row0, box_rows, kg_max;
};
struct Geom {
    bool ok = false;
    float scale = 1.f;
    uint32_t hw_off = 0, n_win = 0, r_pad = 0, slack = 0, corr_off = 0, n_chunks = 0;
    std::vector<BandGeom> bands;
};
}  // namespace

struct BlurTcCache {
    std::map<std::tuple<const AxisTable *, const AxisTable *, uint32_t>, Geom> geoms;  // (vtab, htab, c)
    std::map<std::string, uint32_t> tiles;                                              // tile bytes -> offset in the arena
};
BlurTcCache *blur_tc_cache_new() { return new BlurTcCache(); }
void blur_tc_cache_free(BlurTcCache *c) { delete c; }

size_t blur_tc_smem_bytes(uint32_t box_rows, uint32_t kg_max, uint32_t n_win) {
    return 2 * size_t(65536)                      // T hi | lo, two buffers
           + 2 * size_t(box_rows) * TC_M          // source boxes
           + 2 * size_t(TC_N) * kg_max            // vertical weight slots
           + size_t(n_win) * 128                  // horizontal tile hi | lo
           + 8 * 1184;                            // output staging, per consumer warp
    // (the kernel's static shared memory is a multiple of 1024 bytes, so the dynamic part starts 1024-aligned)
}

static uint32_t window_cols(const StagePlan &s) { return BT_KSUB + 2 * ((blur_radius(s.sigma) * s.c + 15) & ~15u); }

bool blur_tc_eligible(const StagePlan &s, uint32_t pitch, const uint8_t *src) {
    if (!blur_eligible(s)) return false;
    if (pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(src) & 15)) return false;  // TMA
    const uint32_t radius = blur_radius(s.sigma), n_win = window_cols(s);
    if (n_win + BT_KSUB > BT_RING) return false;              // the ring must hold a window and one drained sub-step
    const uint32_t kg = (32 + 2 * radius + 31) / 32 * 32;
    if (96 + kg > TC_KG_MAX) return false;                     // one box holds the band's rows and their halo
    // the box holds the rows the band needs (to a multiple of 8: the swizzle atom); the last group's K steps may read up
    // to 31 rows past it -- zero weights, integer arithmetic, and the bytes behind a box are the next slot or the weight slots
    if (blur_tc_smem_bytes(96 + ((32 + 2 * radius + 7) & ~7u), kg, n_win) > fused_tc_smem_limit()) return false;
    if (s.in_h < 1 || s.in_w < 1) return false;
    return true;
}

// q = round(w * 2^sh) must fit three signed base-128 digits (as fused_tc.cpp)
static int digit_shift(const AxisTable &vt) {
    float maxw = 0.f;
    for (float w : vt.weights) maxw = std::max(maxw, std::fabs(w));
    int sh = 30;
    while (sh > 0 && std::ldexp(double(maxw), sh) > 2080000.0) sh--;
    return sh;
}

static uint32_t intern_tile(BlurTcCache *cache, FusedTcTables *tct, const std::vector<uint8_t> &tile) {
    std::string key(reinterpret_cast<const char *>(tile.data()), tile.size());
    auto it = cache->tiles.find(key);
    if (it != cache->tiles.end()) return it->second;
    const size_t off = (tct->b.size() + 127) & ~size_t(127);
    tct->b.resize(off + tile.size(), 0);
    std::memcpy(&tct->b[off], tile.data(), tile.size());
    cache->tiles.emplace(std::move(key), uint32_t(off));
    return uint32_t(off);
}

static bool build_geom(const StagePlan &s, BlurTcCache *cache, BlurTables *bt, FusedTables *tabs, FusedTcTables *tct, Geom *g) {
    const AxisTable &vt = *s.vtab;
    const uint32_t C = s.c, H = s.in_h, n_e = s.in_w * C, radius = blur_radius(s.sigma);
    const int sh = digit_shift(vt);
    g->scale = std::ldexp(1.0f, -sh);
    g->r_pad = (radius * C + 15) & ~15u;
    g->n_win = BT_KSUB + 2 * g->r_pad;
    g->slack = std::min<uint32_t>(3, (BT_RING - g->n_win) / BT_KSUB);
    if (g->slack < 1) return false;
    g->n_chunks = (n_e + g->r_pad + TC_M - 1) / TC_M;
    // vertical: bands of 128 rows, groups of 32; the group's window starts a multiple of 32 rows into the band's box
    for (uint32_t r0 = 0; r0 < H; r0 += BT_ROWS) {
        BandGeom b{};
        b.r0 = r0;
        b.rows = std::min(BT_ROWS, H - r0);
        b.n_groups = (b.rows + 31) / 32;
        b.box_row0 = vt.entries[r0].left;
        b.grp_off = uint32_t(tabs->info.size());
        tabs->info.resize(tabs->info.size() + 4 * size_t(b.n_groups), 0u);
        for (uint32_t gi = 0; gi < b.n_groups; gi++) {
            const uint32_t ra = r0 + 32 * gi, rb = std::min(r0 + b.rows, ra + 32);
            const uint32_t k0 = vt.entries[ra].left;
            uint32_t k1 = 0;
            for (uint32_t r = ra; r < rb; r++) k1 = std::max(k1, vt.entries[r].left + vt.entries[r].count);
            const uint32_t a_off = (k0 - b.box_row0) / 32 * 32, base = b.box_row0 + a_off;
            const uint32_t kg = (k1 - base + 31) / 32 * 32;
            if (a_off + kg > TC_KG_MAX) return false;
            b.box_rows = std::max(b.box_rows, a_off + ((k1 - base + 7) & ~7u));  // rows actually needed; see blur_tc_eligible
            b.kg_max = std::max(b.kg_max, kg);
            std::vector<uint8_t> tile(size_t(TC_N) * kg, 0);
            int8_t *t8 = reinterpret_cast<int8_t *>(tile.data());
            for (uint32_t r = ra; r < rb; r++) {
                const TapEntry &e = vt.entries[r];
                for (uint32_t t = 0; t < e.count; t++) {
                    const long q = std::lround(std::ldexp(double(vt.weights[e.woff + t]), sh));
                    const long lo = ((q + 64) & 127) - 64;
                    const long q1 = (q - lo) / 128;
                    const long mid = ((q1 + 64) & 127) - 64;
                    const long hi = (q1 - mid) / 128;
                    if (hi < -128 || hi > 127) return false;
                    const uint32_t k = e.left + t - base, j = r - ra;
                    const long dig[3] = {hi, mid, lo};
                    for (uint32_t d = 0; d < 3; d++) {
                        const uint32_t n = d * TC_GROUP_ROWS + j;  // B row: digit-major
                        t8[(size_t(n / 8) * (kg / 16) + k / 16) * 128 + (n % 8) * 16 + k % 16] = int8_t(dig[d]);
                    }
                }
            }
            uint32_t *rec = &tabs->info[b.grp_off + 4 * size_t(gi)];
            rec[0] = a_off; rec[1] = kg; rec[2] = intern_tile(cache, tct, tile); rec[3] = rb - ra;
        }
        g->bands.push_back(b);
    }
    // horizontal: W[n][k], n = output byte (window start + n), k = input byte of the sub-step; in - out = k + r_pad - n
    BlurItem bi{};
    blur_build(s, bt, &tabs->w, &bi);  // interior weights u[2 R + 1] and the per-column border factors
    {
        std::vector<uint8_t> tile(size_t(g->n_win) * 128, 0);
        uint16_t *hi = reinterpret_cast<uint16_t *>(tile.data()), *lo = hi + size_t(g->n_win) * BT_KSUB;
        for (uint32_t n = 0; n < g->n_win; n++)
            for (uint32_t k = 0; k < BT_KSUB; k++) {
                const int d = int(k + g->r_pad) - int(n);
                if (d % int(C) != 0) continue;
                const int tap = d / int(C) + int(radius);
                if (tap < 0 || tap > int(2 * radius)) continue;
                const float w = tabs->w[bi.u_off + uint32_t(tap)] * TC2_WSCALE;
                const __half wh = __float2half_rn(w);
                const size_t at = (size_t(n / 8) * (BT_KSUB / 8) + k / 8) * 64 + (n % 8) * 8 + k % 8;  // K-major core matrices, in f16 elements
                hi[at] = __half_as_ushort(wh);
                lo[at] = __half_as_ushort(__float2half_rn(w - __half2float(wh)));
            }
        g->hw_off = intern_tile(cache, tct, tile);
    }
    // per byte column: border factor of its pixel / 16 (the weights are stored x 16); zero beyond the row
    tabs->w.resize((tabs->w.size() + 3) & ~size_t(3), 0.0f);
    g->corr_off = uint32_t(tabs->w.size());
    tabs->w.resize(tabs->w.size() + size_t(g->n_chunks) * TC_M, 0.0f);
    for (uint32_t cb = 0; cb < n_e; cb++) tabs->w[g->corr_off + cb] = tabs->w[bi.corrh_off + cb / C] * (1.0f / TC2_WSCALE);
    g->ok = true;
    return true;
}

int blur_tc_build(const StagePlan &s, const uint8_t *src, uint32_t src_pitch, uint8_t *dst, uint32_t dst_pitch, BlurTcCache *cache,
                  BlurTables *bt, FusedTables *tabs, FusedTcTables *tct, std::vector<BlurTcItem> *items) {
    const auto key = std::make_tuple(s.vtab.get(), s.htab.get(), s.c);
    auto it = cache->geoms.find(key);
    if (it == cache->geoms.end()) {
        Geom g;
        if (!build_geom(s, cache, bt, tabs, tct, &g)) g.ok = false;
        it = cache->geoms.emplace(key, std::move(g)).first;
    }
    const Geom &g = it->second;
    if (!g.ok) return FANLIN_EINVAL;
    for (const BandGeom &b : g.bands) {
        BlurTcItem f{};
        f.src = src; f.dst = dst; f.src_pitch = src_pitch; f.dst_pitch = dst_pitch; f.src_h = s.in_h;
        f.n_e = s.in_w * s.c; f.n_chunks = g.n_chunks;
        f.band_r0 = b.r0; f.band_rows = b.rows;
        f.grp_off = b.grp_off; f.n_groups = b.n_groups;
        f.box_row0 = b.box_row0; f.box_rows = b.box_rows; f.kg_max = b.kg_max;
        f.scale = g.scale;
        f.hw_off = g.hw_off; f.n_win = g.n_win; f.r_pad = g.r_pad; f.slack = g.slack; f.corr_off = g.corr_off;
        const uint32_t rb = blur_radius(s.sigma) * s.c;  // output bytes [rb, n_e - rb) have their whole window inside the row
        f.c_lo = std::min(rb, f.n_e); f.c_hi = f.n_e > rb ? f.n_e - rb : 0;
        items->push_back(f);
    }
    return FANLIN_OK;
}

}  // namespace fanlin
