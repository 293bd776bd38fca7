// Fast Gaussian blur path (kernels_blur.cu): per-job descriptor and host-side tables.
#pragma once
#include <map>
#include <tuple>
#include <vector>

#include "plan.h"

namespace fanlin {

struct BlurItem {
    const uint8_t *src;
    uint8_t *dst;
    float *tmp;            // f32 intermediate [h][w * c]
    uint32_t w, h, c, src_pitch;
    uint32_t radius;       // ceil(2 sigma - 0.5): taps on either side of the centre (blur_radius)
    uint32_t taps_pad;     // 2 radius + 1 rounded up to a multiple of 8 (zero weights beyond)
    uint32_t u_off;        // float offset of the interior weights u[taps_pad]
    uint32_t corrv_off;    // float offset of the per-row border factors [h]
    uint32_t corrh_off;    // float offset of the per-column border factors [w]
    uint32_t aligned4;     // rows can be read as 4-byte words
};

// Interior weights and border factors, deduplicated per (sigma, axis length).
struct BlurTables {
    std::map<std::tuple<uint32_t, uint32_t>, uint32_t> corr;  // (sigma bits, n) -> float offset
    std::map<uint32_t, uint32_t> u;                            // sigma bits -> float offset
};

uint32_t blur_radius(float sigma);
// True when stage `s` (a Gaussian stage) can take the fast blur kernels.
bool blur_eligible(const StagePlan &s);
// Appends tables as needed to `w` and fills the geometry/table fields of `item`.
void blur_build(const StagePlan &s, BlurTables *bt, std::vector<float> *w, BlurItem *item);

size_t blur_v_smem(uint32_t radius, uint32_t taps_pad);
size_t blur_h_smem(uint32_t radius, uint32_t taps_pad, uint32_t c);

}  // namespace fanlin
