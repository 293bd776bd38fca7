// Host-side planning: geometry of a job and the per-axis filter tables.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "common.h"

namespace fanlin {

void set_error(const std::string &msg);
const char *get_error();

// Filter taps of one axis (n_in -> n_out), built with the f32 recipe of
// image-0.25.6 imageops/sample.rs.
struct AxisTable {
    uint32_t kind = 0, n_in = 0, n_out = 0;
    float sigma = 0.f;
    uint32_t max_taps = 0;
    std::vector<TapEntry> entries;  // woff relative to `weights`
    std::vector<float> weights;
};

using TableKey = std::tuple<uint32_t, uint32_t, uint32_t, uint32_t>;  // kind, sigma bits, n_in, n_out
TableKey table_key(uint32_t kind, float sigma, uint32_t n_in, uint32_t n_out);
std::shared_ptr<const AxisTable> build_axis_table(uint32_t kind, float sigma, uint32_t n_in, uint32_t n_out);

// The same taps seen from the other end of both axes: entry n_out - 1 - o reads source [n_in - left - count, n_in - left) with
// the weights of entry o reversed -- the table of an axis that is stored mirrored (EXIF flips, cached per table).
std::shared_ptr<const AxisTable> mirror_axis_table(const std::shared_ptr<const AxisTable> &t);

// resize_dimensions of image-0.25.6 math/utils.rs.
void resize_dimensions(uint32_t w, uint32_t h, uint32_t nw, uint32_t nh, bool fill, uint32_t *ow, uint32_t *oh);

// One device stage of a job, host view.
struct StagePlan {
    bool present = false;
    bool separable = false;  // false: compose only (colour op / letterbox / to_rgba8 / copy)
    bool src_is_input = true;  // else reads the previous stage's canvas
    uint32_t in_pitch = 0;     // bytes per row of that canvas when it is not tightly packed (0: in_w * c_mem)
    uint32_t in_w = 0, in_h = 0, c_mem = 0, c = 0, color_op = COLOR_NONE;
    uint32_t s_in = SAMPLE_U8, s_out = SAMPLE_U8;  // subpixel type read / written (u16 / f32 stages take the kernels of kernels_deep.cu)
    uint32_t v_kind = 0, h_kind = 0;
    float sigma = 0.f;
    uint32_t v_out = 0, h_out = 0;  // full filtered size (table n_out) per axis
    uint32_t oy0 = 0, n_rows = 0, ox0 = 0, n_cols = 0;
    uint32_t sx0 = 0, n_sx = 0, sy0 = 0, n_sy = 0;
    uint32_t canvas_w = 0, canvas_h = 0, c_out = 0, dst_x = 0, dst_y = 0, epi = EPI_PLAIN, fill = 0;
    uint32_t canvas_pitch = 0;  // bytes per canvas row when the canvas is scratch with padded rows (0: canvas_w * c_out)
    uint32_t min_bands = 1;     // tensor-core resample: cut the image into at least this many bands (CTAs) -- set for small batches, where
                                // one CTA per image would leave the GPU to a handful of SMs (latency of a single request)
    std::shared_ptr<const AxisTable> vtab, htab;
};

// EXIF orientation (apply_orientation, handler.rs:221-223): the stored image is turned -- and the
// colour op applied, it commutes with a permutation of pixels -- by a pass of its own into
// scratch; stages a / b are planned for `job`, the request as it looks behind that pass.
struct OrientPlan {
    bool present = false;
    uint32_t orient = 0;                    // 2..8
    uint32_t c_mem = 0, c = 0, color_op = 0;  // the stored image's channels, after the colour op, the op
    uint32_t sample = SAMPLE_U8;            // subpixel type of the stored and of the oriented image
    fanlin_job job{};                       // oriented size, c channels, 16-byte aligned pitch, no colour flags; src unset
};

struct JobPlan {
    fanlin_plan pub{};
    OrientPlan pre;
    uint32_t post_c_in = 0;  // FANLIN_TO_RGB8: channels of the final image before to_rgb8 (0: no conversion pass)
    uint32_t post_s_in = SAMPLE_U8;  // ... and its subpixel type
    uint32_t post_c_out = 3;         // channels that pass writes: 3 (to_rgb8), or 4 = (l, l, l, 255) behind a gray canvas (EPI_GRAY)
    bool post_ycbcr = false;         // FANLIN_TO_YCBCR: that pass writes the planes Y, Cb, Cr of to_rgb8() of the final image
    StagePlan a;  // colour op + resample + letterbox (+ to_rgba8), or compose
    StagePlan b;  // blur
};

// EXIF orientation applied AFTER the resample (fast paths): the Lanczos3 stage `a` of a job planned for the oriented image,
// restated on the image AS STORED -- the oriented axes' tables on the stored axes they run along, mirrored where the
// axis is stored mirrored, the produced rectangle mapped likewise, plain output.  Orienting its (small) output with the
// same EXIF value gives the oriented stage's pixels up to the order of the f32 additions: a separable filter commutes
// with flips and transposition, and the pass over the stored source at full resolution disappears.
StagePlan stored_axes_stage(const StagePlan &a, const fanlin_job &stored, uint32_t exif);

// Returns FANLIN_OK or FANLIN_EINVAL (message via set_error).  with_tables: also
// build the axis tables (not needed for fanlin_plan_job).
int plan_job(const fanlin_job &job, JobPlan *out, bool with_tables);

}  // namespace fanlin
