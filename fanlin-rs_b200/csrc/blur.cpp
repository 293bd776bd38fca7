// Host side of the fast blur path: interior weights and the border renormalisation factors,
// both read off the crate-recipe tap tables of plan.cpp.
#include "blur.h"

#include <algorithm>
#include <cstring>

namespace fanlin {

bool blur_eligible(const StagePlan &s) {
    if (!s.present || !s.separable || s.v_kind != KIND_GAUSSIAN || !s.vtab || !s.htab) return false;
    if (s.color_op != COLOR_NONE || s.c_mem != s.c || s.epi != EPI_PLAIN) return false;
    if (s.n_rows != s.in_h || s.n_cols != s.in_w) return false;
    const uint32_t radius = uint32_t(2.0f * s.sigma);
    if (radius < 1 || radius > 64) return false;
    return blur_v_smem(radius, (2 * radius + 1 + 7) & ~7u) <= 200 * 1024 && blur_h_smem(radius, (2 * radius + 1 + 7) & ~7u, s.c) <= 200 * 1024;
}

static uint32_t sigma_bits(float s) {
    uint32_t b;
    std::memcpy(&b, &s, 4);
    return b;
}

void blur_build(const StagePlan &s, BlurTables *bt, std::vector<float> *w, BlurItem *item) {
    const uint32_t radius = uint32_t(2.0f * s.sigma), taps = 2 * radius + 1, taps_pad = (taps + 7) & ~7u;
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
