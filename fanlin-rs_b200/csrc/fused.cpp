// Host side of the fused separable resample: band / chunk decomposition and the
// scatter tables (see fused.h).
#include "fused.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <tuple>

namespace fanlin {

namespace {

struct BandTab {
    uint32_t r0, rows, y0, n_y, vw_off, vinfo_off, vinfo_stride;
};
struct GeomTab {
    bool ok = false;
    uint32_t px0 = 0, n_px = 0, hw_off = 0, hinfo_off = 0;
    std::vector<BandTab> bands;
};
using GeomKey = std::tuple<const AxisTable *, const AxisTable *, uint32_t, uint32_t, uint32_t, uint32_t>;


// Scatter view of outputs [o0, o0+n) of `t` over sources [s0, s0+n_s): weights
// [n_s][8] (slot = (o - slot_base) % 8) and live/flush masks.  Returns false when
// more than 8 outputs are live at one source index.
}  // namespace

bool fused_scatter(const AxisTable &t, uint32_t o0, uint32_t n, uint32_t slot_base, uint32_t s0, uint32_t n_s, float scale,
                   float *w, uint32_t *info) {
    for (uint32_t o = o0; o < o0 + n; o++) {
        const TapEntry &e = t.entries[o];
        const uint32_t slot = (o - slot_base) % FUSED_SLOTS;
        for (uint32_t k = 0; k < e.count; k++) {
            const uint32_t s = e.left + k;
            if (s < s0 || s >= s0 + n_s) return false;
            if (info[s - s0] & (1u << slot)) return false;  // slot still occupied: > 8 live outputs
            info[s - s0] |= 1u << slot;
            if (w) w[size_t(s - s0) * FUSED_SLOTS + slot] = t.weights[e.woff + k] * scale;
        }
        info[e.left + e.count - 1 - s0] |= 1u << (8 + slot);
    }
    return true;
}

struct FusedCache {
    std::map<GeomKey, GeomTab> geoms;
};
FusedCache *fused_cache_new() { return new FusedCache(); }
void fused_cache_free(FusedCache *c) { delete c; }

bool fused_eligible(const StagePlan &s, const fanlin_job &job) {
    if (!s.present || !s.separable || !s.vtab || !s.htab) return false;
    if (s.n_rows == 0 || s.n_cols == 0) return false;
    if (s.v_kind == KIND_GAUSSIAN) return false;  // blur has its own kernel
    const uint32_t pitch = s.src_is_input ? (job.src_pitch ? job.src_pitch : job.src_w * job.src_channels) : s.in_w * s.c_mem;
    if (pitch % 4 != 0) return false;
    if ((uint64_t(s.in_w) * s.c) % 4 != 0) return false;  // whole 4-element words per row
    if (s.src_is_input && (reinterpret_cast<uintptr_t>(job.src) & 3)) return false;
    // quick bound on live outputs per source index: taps / stride between outputs
    if (s.vtab->max_taps > 8 * std::max(1u, s.in_h / std::max(1u, s.v_out)) + 8) return false;
    if (s.htab->max_taps > 8 * std::max(1u, s.in_w / std::max(1u, s.h_out)) + 8) return false;
    return true;
}

static const GeomTab &geom_of(const StagePlan &s, FusedCache *cache, FusedTables *tabs) {
    const GeomKey key(s.vtab.get(), s.htab.get(), s.oy0 | (s.n_rows << 16), s.ox0 | (s.n_cols << 16), s.c, s.c_mem);
    auto it = cache->geoms.find(key);
    if (it == cache->geoms.end()) {
        GeomTab g;
        const float vscale = std::ldexp(1.0f, FUSED_V_SCALE_LOG2), hscale = std::ldexp(1.0f, FUSED_H_SCALE_LOG2);
        // horizontal: one table for the whole produced rect
        g.px0 = s.sx0 / 4 * 4;
        g.n_px = s.sx0 + s.n_sx - g.px0;
        g.hw_off = uint32_t(tabs->w.size());
        g.hinfo_off = uint32_t(tabs->info.size());
        tabs->w.resize(tabs->w.size() + size_t(g.n_px) * FUSED_SLOTS, 0.0f);
        tabs->info.resize(tabs->info.size() + g.n_px, 0u);
        bool ok = fused_scatter(*s.htab, s.ox0, s.n_cols, s.ox0, g.px0, g.n_px, hscale, &tabs->w[g.hw_off], &tabs->info[g.hinfo_off]);
        // vertical: per band, weights for the band and masks per warp sub-band
        const uint32_t max_band = fused_max_band(s.c, s.c_mem);
        const uint32_t n_bands = (s.n_rows + max_band - 1) / max_band;
        const uint32_t band_rows = (s.n_rows + n_bands - 1) / n_bands;
        for (uint32_t b = 0; b < n_bands && ok; b++) {
            BandTab bt{};
            bt.r0 = b * band_rows;
            bt.rows = std::min(band_rows, s.n_rows - bt.r0);
            const uint32_t o_first = s.oy0 + bt.r0, o_last = o_first + bt.rows - 1;
            bt.y0 = s.vtab->entries[o_first].left;
            uint32_t y1 = 0;
            for (uint32_t o = o_first; o <= o_last; o++) y1 = std::max(y1, s.vtab->entries[o].left + s.vtab->entries[o].count);
            bt.n_y = y1 - bt.y0;
            std::vector<uint32_t> band_info(bt.n_y, 0u);
            ok = fused_scatter(*s.vtab, o_first, bt.rows, o_first, bt.y0, bt.n_y, 0.f, nullptr, band_info.data());  // <= 8 live?
            if (!ok) break;
            const uint32_t rq = bt.rows / FUSED_WARPS, rrem = bt.rows % FUSED_WARPS;  // balanced split
            uint32_t max_n = 0;
            struct WarpRange { uint32_t ra, rb, ya, yb; } wr[FUSED_WARPS];
            for (uint32_t w = 0; w < FUSED_WARPS; w++) {
                wr[w].ra = w * rq + std::min(w, rrem);
                wr[w].rb = wr[w].ra + rq + (w < rrem ? 1 : 0);
                wr[w].ya = wr[w].yb = 0;
                if (wr[w].rb > wr[w].ra) {
                    wr[w].ya = s.vtab->entries[o_first + wr[w].ra].left;
                    for (uint32_t r = wr[w].ra; r < wr[w].rb; r++)
                        wr[w].yb = std::max(wr[w].yb, s.vtab->entries[o_first + r].left + s.vtab->entries[o_first + r].count);
                    max_n = std::max(max_n, wr[w].yb - wr[w].ya);
                }
            }
            bt.vinfo_stride = 4 + max_n;
            bt.vinfo_off = uint32_t(tabs->info.size());
            tabs->info.resize(tabs->info.size() + size_t(bt.vinfo_stride) * FUSED_WARPS, 0u);
            bt.vw_off = uint32_t(tabs->w.size());
            for (uint32_t w = 0; w < FUSED_WARPS && ok; w++) {
                // per warp: weights [rows streamed][8], zero for outputs of other sub-bands, so the
                // kernel can run all 8 slots unconditionally
                const uint32_t n = wr[w].yb - wr[w].ya;
                const size_t woff = tabs->w.size();
                tabs->w.resize(woff + size_t(n) * FUSED_SLOTS, 0.0f);
                uint32_t *blk = &tabs->info[bt.vinfo_off + size_t(w) * bt.vinfo_stride];
                blk[0] = wr[w].rb > wr[w].ra ? wr[w].ya - bt.y0 : 0;  // first source row, relative to the band window
                blk[1] = n;                                            // rows to stream
                blk[2] = wr[w].ra;
                blk[3] = uint32_t(woff);
                if (n)
                    ok = fused_scatter(*s.vtab, o_first + wr[w].ra, wr[w].rb - wr[w].ra, o_first, wr[w].ya, n, vscale,
                                 &tabs->w[woff], blk + 4);
            }
            g.bands.push_back(bt);
        }
        g.ok = ok;
        it = cache->geoms.emplace(key, std::move(g)).first;
    }
    return it->second;
}

bool fused_geometry_ok(const StagePlan &s, FusedCache *cache, FusedTables *tabs) { return geom_of(s, cache, tabs).ok; }

int fused_build(const StagePlan &s, const uint8_t *src, uint32_t src_pitch, uint8_t *dst, FusedCache *cache,
                FusedTables *tabs, std::vector<FusedItem> *items) {
    const GeomTab &g = geom_of(s, cache, tabs);
    if (!g.ok) return FANLIN_EINVAL;
    const uint32_t chunk_px = fused_chunk_px(s.c);
    for (size_t b = 0; b < g.bands.size(); b++) {
        const BandTab &bt = g.bands[b];
        FusedItem f{};
        f.src = src; f.dst = dst; f.src_pitch = src_pitch;
        f.c_mem = s.c_mem; f.c = s.c; f.color_op = s.color_op;
        f.px0 = g.px0; f.n_px = g.n_px;
        f.chunk_px = chunk_px; f.n_chunks = (g.n_px + chunk_px - 1) / chunk_px;
        f.band_r0 = bt.r0; f.band_rows = bt.rows; f.y0 = bt.y0; f.n_y = bt.n_y;
        f.vw_off = bt.vw_off; f.vinfo_off = bt.vinfo_off; f.vinfo_stride = bt.vinfo_stride;
        f.hw_off = g.hw_off; f.hinfo_off = g.hinfo_off;
        f.n_cols = s.n_cols;
        f.dst_pitch = s.canvas_pitch ? s.canvas_pitch : s.canvas_w * s.c_out; f.c_out = s.c_out; f.canvas_w = s.canvas_w; f.canvas_h = s.canvas_h;
        f.dst_x = s.dst_x; f.dst_y = s.dst_y; f.epi = s.epi; f.fill = s.fill;
        f.first_band = b == 0; f.last_band = b + 1 == g.bands.size();
        items->push_back(f);
    }
    return FANLIN_OK;
}

}  // namespace fanlin
