// Host side of the fast blur path: interior weights and the border renormalisation factors,
// both read off the crate-recipe tap tables of plan.cpp.
#include "blur.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace fanlin {

// Radius of the crate's Gaussian window for this sigma: left = floor(o + 0.5 - 2 sigma), right = ceil(o + 0.5 + 2 sigma)
// (plan.cpp AxisGeom::window), i.e. ceil(2 sigma - 0.5) taps on either side -- NOT floor(2 sigma): sigma = 10.3 has 43
// taps, not 41.  Query::blur() only yields integers, the C ABI takes any float.
uint32_t blur_radius(float sigma) {
    const float r = std::ceil(2.0f * sigma - 0.5f);
    return r < 0.f ? 0u : uint32_t(r);
}

// The fast kernels assume a symmetric window of 2 R + 1 taps truncated at the borders; checked against the tables
// the exact path would use, so a sigma whose f32 window arithmetic lands elsewhere takes the exact path instead.
static bool table_is_symmetric(const AxisTable &t, uint32_t n, uint32_t radius) {
    if (t.entries.size() != n) return false;
    for (uint32_t o = 0; o < n; o++) {
        const TapEntry &e = t.entries[o];
        const uint32_t l = o > radius ? o - radius : 0u, r = std::min(n, o + radius + 1);
        if (e.left != l || e.left + e.count != r) return false;
    }
    return true;
}

bool blur_eligible(const StagePlan &s) {
    if (!s.present || !s.separable || s.v_kind != KIND_GAUSSIAN || !s.vtab || !s.htab) return false;
    if (s.color_op != COLOR_NONE || s.c_mem != s.c || s.epi != EPI_PLAIN) return false;
    if (s.n_rows != s.in_h || s.n_cols != s.in_w) return false;
    const uint32_t radius = blur_radius(s.sigma);
    if (radius < 1 || radius > 64) return false;
    if (!table_is_symmetric(*s.vtab, s.in_h, radius) || !table_is_symmetric(*s.htab, s.in_w, radius)) return false;
    return blur_v_smem(radius, (2 * radius + 1 + 7) & ~7u) <= 200 * 1024 && blur_h_smem(radius, (2 * radius + 1 + 7) & ~7u, s.c) <= 200 * 1024;
}

static uint32_t sigma_bits(float s) {
    uint32_t b;
    std::memcpy(&b, &s, 4);
    return b;
}

void blur_build(const StagePlan &s, BlurTables *bt, std::vector<float> *w, BlurItem *item) {
    const uint32_t radius = blur_radius(s.sigma), taps = 2 * radius + 1, taps_pad = (taps + 7) & ~7u;
    const uint32_t sb = sigma_bits(s.sigma);
    auto iu = bt->u.find(sb);
    if (iu == bt->u.end()) {
        // interior weights: the centre output of an axis long enough to hold the whole window
        auto t = build_axis_table(KIND_GAUSSIAN, s.sigma, 4 * radius + 4, 4 * radius + 4);
        const TapEntry &e = t->entries[2 * radius + 2];
        const uint32_t off = uint32_t(w->size());
        w->resize(w->size() + taps_pad, 0.0f);
        for (uint32_t k = 0; k < std::min(taps, e.count); k++) (*w)[off + k] = t->weights[e.woff + k];
        iu = bt->u.emplace(sb, off).first;
    }
    const uint32_t u_off = iu->second;
    const float u_centre = (*w)[u_off + radius];
    auto corr_of = [&](const std::shared_ptr<const AxisTable> &t, uint32_t n) {
        auto key = std::make_tuple(sb, n);
        auto it = bt->corr.find(key);
        if (it != bt->corr.end()) return it->second;
        const uint32_t off = uint32_t(w->size());
        w->resize(w->size() + ((n + 3) & ~3u), 1.0f);
        for (uint32_t j = 0; j < n; j++) {
            const TapEntry &e = t->entries[j];
            (*w)[off + j] = t->weights[e.woff + (j - e.left)] / u_centre;  // tap at source index j is always inside
        }
        bt->corr.emplace(key, off);
        return off;
    };
    item->w = s.in_w; item->h = s.in_h; item->c = s.c;
    item->radius = radius; item->taps_pad = taps_pad;
    item->u_off = u_off;
    item->corrv_off = corr_of(s.vtab, s.in_h);
    item->corrh_off = corr_of(s.htab, s.in_w);
}

}  // namespace fanlin
