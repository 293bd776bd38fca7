// Host-side planning for the pixel-transform stage: which device stages a job
// needs, their geometry, and the per-axis filter tables.
//
// Sequencing follows reference src/handler.rs:224-255 (stills) and :329-355 (GIF
// frames); geometry follows image-0.25.6 math/utils.rs (resize_dimensions),
// dynimage.rs (resize, resize_to_fill, crop) and imageops/sample.rs (tap windows
// and weights) as restated in SURVEY.md Appendix A.
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

namespace fanlin {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
const char *get_error() { return g_error.c_str(); }

// ---- resize_dimensions ----------------------------------------------------

static uint64_t round_u64(double v) {
    double r = std::round(v);  // half away from zero, like f64::round
    if (!(r > 0.0)) return 0;
    if (r >= 18446744073709551616.0) return UINT64_MAX;
    return static_cast<uint64_t>(r);
}

void resize_dimensions(uint32_t w, uint32_t h, uint32_t nw, uint32_t nh, bool fill, uint32_t *ow, uint32_t *oh) {
    const double wr = double(nw) / double(w), hr = double(nh) / double(h);
    const double ratio = fill ? std::fmax(wr, hr) : std::fmin(wr, hr);
    uint64_t a = std::max<uint64_t>(round_u64(double(w) * ratio), 1);
    uint64_t b = std::max<uint64_t>(round_u64(double(h) * ratio), 1);
    if (a > UINT32_MAX) {
        const double r2 = double(UINT32_MAX) / double(w);
        *ow = UINT32_MAX;
        *oh = (uint32_t)std::max<uint64_t>(std::min<uint64_t>(round_u64(double(h) * r2), UINT32_MAX), 1);
    } else if (b > UINT32_MAX) {
        const double r2 = double(UINT32_MAX) / double(h);
        *ow = (uint32_t)std::max<uint64_t>(std::min<uint64_t>(round_u64(double(w) * r2), UINT32_MAX), 1);
        *oh = UINT32_MAX;
    } else {
        *ow = (uint32_t)a;
        *oh = (uint32_t)b;
    }
}

// ---- axis tables ------------------------------------------------------------

namespace {

struct FilterFn {
    uint32_t kind;
    float sigma;
    float support() const {
        if (kind == KIND_NEAREST) return 0.0f;
        if (kind == KIND_LANCZOS3) return 3.0f;
        return 2.0f * sigma;
    }
    static float sinc(float t) {
        const float a = t * 3.14159265358979323846f;
        return t == 0.0f ? 1.0f : sinf(a) / a;
    }
    float operator()(float x) const {
        switch (kind) {
        case KIND_NEAREST: return 1.0f;
        case KIND_LANCZOS3: return fabsf(x) < 3.0f ? sinc(x) * sinc(x / 3.0f) : 0.0f;
        default: {
            const float norm = 1.0f / (sqrtf(2.0f * 3.14159265358979323846f) * sigma);
            return norm * expf(-(x * x) / (2.0f * (sigma * sigma)));
        }
        }
    }
};

struct AxisGeom {
    float ratio, sratio, src_support;
    AxisGeom(const FilterFn &f, uint32_t n_in, uint32_t n_out) {
        ratio = float(n_in) / float(n_out);
        sratio = ratio < 1.0f ? 1.0f : ratio;
        src_support = f.support() * sratio;
    }
    // [left, right) of output o and the tap centre
    void window(uint32_t n_in, uint32_t o, int64_t *left, int64_t *right, float *centre) const {
        const float c = (float(o) + 0.5f) * ratio;
        int64_t l = (int64_t)floorf(c - src_support);
        l = std::min<int64_t>(std::max<int64_t>(l, 0), int64_t(n_in) - 1);
        int64_t r = (int64_t)ceilf(c + src_support);
        r = std::min<int64_t>(std::max<int64_t>(r, l + 1), int64_t(n_in));
        *left = l;
        *right = r;
        *centre = c - 0.5f;
    }
};

std::mutex g_table_mu;
std::map<TableKey, std::shared_ptr<const AxisTable>> g_table_cache;
size_t g_table_cache_floats = 0;

}  // namespace

TableKey table_key(uint32_t kind, float sigma, uint32_t n_in, uint32_t n_out) {
    uint32_t bits;
    std::memcpy(&bits, &sigma, 4);
    return TableKey(kind, kind == KIND_GAUSSIAN ? bits : 0u, n_in, n_out);
}

std::shared_ptr<const AxisTable> build_axis_table(uint32_t kind, float sigma, uint32_t n_in, uint32_t n_out) {
    const TableKey key = table_key(kind, sigma, n_in, n_out);
    {
        std::lock_guard<std::mutex> lk(g_table_mu);
        auto it = g_table_cache.find(key);
        if (it != g_table_cache.end()) return it->second;
    }
    auto t = std::make_shared<AxisTable>();
    t->kind = kind;
    t->sigma = sigma;
    t->n_in = n_in;
    t->n_out = n_out;
    const FilterFn f{kind, sigma};
    const AxisGeom g(f, n_in, n_out);
    t->entries.resize(n_out);
    t->weights.reserve(size_t(n_out) * size_t(2.0f * g.src_support + 3.0f));
    for (uint32_t o = 0; o < n_out; o++) {
        int64_t l, r;
        float c;
        g.window(n_in, o, &l, &r, &c);
        const size_t base = t->weights.size();
        float sum = 0.0f;
        for (int64_t i = l; i < r; i++) {
            const float w = f((float(i) - c) / g.sratio);
            t->weights.push_back(w);
            sum += w;
        }
        for (size_t k = base; k < t->weights.size(); k++) t->weights[k] /= sum;
        t->entries[o] = TapEntry{uint32_t(l), uint32_t(r - l), uint32_t(base)};
        t->max_taps = std::max(t->max_taps, uint32_t(r - l));
    }
    std::lock_guard<std::mutex> lk(g_table_mu);
    if (g_table_cache_floats > (64u << 20)) {  // bound the cache (256 MB of weights)
        g_table_cache.clear();
        g_table_cache_floats = 0;
    }
    g_table_cache_floats += t->weights.size() + 3 * t->entries.size();
    g_table_cache[key] = t;
    return t;
}

std::shared_ptr<const AxisTable> mirror_axis_table(const std::shared_ptr<const AxisTable> &t) {
    static std::mutex mu;
    static std::map<const AxisTable *, std::pair<std::shared_ptr<const AxisTable>, std::shared_ptr<const AxisTable>>> cache;  // original (kept alive) -> mirrored
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(t.get());
    if (it != cache.end()) return it->second.second;
    if (cache.size() > 4096) cache.clear();
    auto m = std::make_shared<AxisTable>();
    m->kind = t->kind; m->n_in = t->n_in; m->n_out = t->n_out; m->sigma = t->sigma; m->max_taps = t->max_taps;
    m->entries.resize(t->entries.size());
    m->weights.reserve(t->weights.size());
    const uint32_t n_out = uint32_t(t->entries.size());
    for (uint32_t o2 = 0; o2 < n_out; o2++) {
        const TapEntry &e = t->entries[n_out - 1 - o2];
        m->entries[o2] = TapEntry{t->n_in - e.left - e.count, e.count, uint32_t(m->weights.size())};
        for (uint32_t k = 0; k < e.count; k++) m->weights.push_back(t->weights[e.woff + e.count - 1 - k]);
    }
    cache[t.get()] = std::make_pair(t, std::shared_ptr<const AxisTable>(m));
    return cache[t.get()].second;
}

StagePlan stored_axes_stage(const StagePlan &a, const fanlin_job &stored, uint32_t e) {
    StagePlan in;
    in.present = true; in.separable = true; in.src_is_input = true;
    in.in_w = stored.src_w; in.in_h = stored.src_h; in.c_mem = stored.src_channels; in.c = stored.src_channels; in.color_op = COLOR_NONE;
    if (stored.flags & FANLIN_GRAYSCALE) {
        if (stored.src_channels >= 3) { in.color_op = COLOR_GRAY; in.c = stored.src_channels - 2; }
    } else if (stored.flags & FANLIN_INVERSE) {
        in.color_op = COLOR_INVERT;
    }
    in.v_kind = a.v_kind; in.h_kind = a.h_kind;
    const bool swap = e >= 5;
    // oriented (xo, yo) = stored (sx, sy) (Orientation::from_exif + apply_orientation, handler.rs:221-223):
    //   e < 5: sx = xo, mirrored for 2 and 3; sy = yo, mirrored for 3 and 4
    //   e >= 5: sx = yo, mirrored for 7 and 8; sy = xo, mirrored for 6 and 7
    const bool mx = swap ? (e == 7 || e == 8) : (e == 2 || e == 3);  // the stored x axis runs against its oriented axis
    const bool my = swap ? (e == 6 || e == 7) : (e == 3 || e == 4);
    const std::shared_ptr<const AxisTable> &tx = swap ? a.vtab : a.htab, &ty = swap ? a.htab : a.vtab;  // oriented tables of the stored x / y axes
    const uint32_t x_out = swap ? a.v_out : a.h_out, y_out = swap ? a.h_out : a.v_out;
    const uint32_t x0 = swap ? a.oy0 : a.ox0, xn = swap ? a.n_rows : a.n_cols, y0 = swap ? a.ox0 : a.oy0, yn = swap ? a.n_cols : a.n_rows;
    in.htab = tx ? (mx ? mirror_axis_table(tx) : tx) : nullptr;
    in.vtab = ty ? (my ? mirror_axis_table(ty) : ty) : nullptr;
    in.h_out = x_out; in.v_out = y_out;
    in.ox0 = mx ? x_out - x0 - xn : x0; in.n_cols = xn;
    in.oy0 = my ? y_out - y0 - yn : y0; in.n_rows = yn;
    auto window = [](const AxisTable &t, uint32_t o0, uint32_t n, uint32_t *s0, uint32_t *ns) {
        uint32_t lo = ~0u, hi = 0;
        for (uint32_t o = o0; o < o0 + n; o++) { lo = std::min(lo, t.entries[o].left); hi = std::max(hi, t.entries[o].left + t.entries[o].count); }
        *s0 = lo; *ns = hi - lo;
    };
    if (in.htab) window(*in.htab, in.ox0, in.n_cols, &in.sx0, &in.n_sx);
    if (in.vtab) window(*in.vtab, in.oy0, in.n_rows, &in.sy0, &in.n_sy);
    in.canvas_w = in.n_cols; in.canvas_h = in.n_rows; in.c_out = in.c; in.dst_x = in.dst_y = 0; in.epi = EPI_PLAIN; in.fill = 0;
    in.min_bands = a.min_bands;
    return in;
}

static void axis_window(uint32_t kind, float sigma, uint32_t n_in, uint32_t n_out, uint32_t o0, uint32_t n,
                        uint32_t *s0, uint32_t *s1) {
    const FilterFn f{kind, sigma};
    const AxisGeom g(f, n_in, n_out);
    int64_t l0, r0, l1, r1;
    float c;
    g.window(n_in, o0, &l0, &r0, &c);
    g.window(n_in, o0 + n - 1, &l1, &r1, &c);
    *s0 = uint32_t(l0);
    *s1 = uint32_t(std::max(r0, r1));
}

// ---- job planning -----------------------------------------------------------

static uint32_t absdiff(uint32_t a, uint32_t b) { return a > b ? a - b : b - a; }

int plan_job(const fanlin_job &job, JobPlan *out, bool with_tables) {
    if (job.orientation > 8) { set_error("fanlin: orientation must be an EXIF value 0..8"); return FANLIN_EINVAL; }
    if (((job.flags & FANLIN_TO_RGBA8) != 0) + ((job.flags & FANLIN_TO_RGB8) != 0) + ((job.flags & FANLIN_TO_YCBCR) != 0) > 1) { set_error("fanlin: TO_RGBA8, TO_RGB8 and TO_YCBCR exclude each other"); return FANLIN_EINVAL; }
    if (job.orientation >= 2) {
        // Orientation::from_exif + apply_orientation: 5..8 swap width and height.  Plan the request as it
        // looks behind the orientation pass, which also applies the colour op.
        if (job.src_w == 0 || job.src_h == 0 || job.src_channels < 1 || job.src_channels > 4 || job.src_sample > SAMPLE_F32) {
            set_error("fanlin: bad source image");
            return FANLIN_EINVAL;
        }
        if (reinterpret_cast<uintptr_t>(job.src) % sample_bytes(job.src_sample) != 0) { set_error("fanlin: src / dst not aligned to the subpixel size"); return FANLIN_EINVAL; }
        OrientPlan pre;
        pre.present = true;
        pre.orient = job.orientation;
        pre.c_mem = pre.c = job.src_channels;
        pre.sample = job.src_sample;
        pre.color_op = COLOR_NONE;
        if (job.flags & FANLIN_GRAYSCALE) {  // (Rgb32F / Rgba32F keep their channels: the luma is replicated)
            if (job.src_channels >= 3) { pre.color_op = COLOR_GRAY; pre.c = job.src_sample == SAMPLE_F32 ? job.src_channels : job.src_channels - 2; }
        } else if (job.flags & FANLIN_INVERSE) {
            pre.color_op = COLOR_INVERT;
        }
        fanlin_job &ej = pre.job;
        ej = job;
        ej.orientation = 0;
        ej.flags &= ~uint32_t(FANLIN_GRAYSCALE | FANLIN_INVERSE);
        if (job.orientation >= 5) { ej.src_w = job.src_h; ej.src_h = job.src_w; }
        ej.src_channels = pre.c;
        ej.src_pitch = (ej.src_w * pre.c * sample_bytes(pre.sample) + 15u) & ~15u;
        ej.src = nullptr;
        const int rc = plan_job(ej, out, with_tables);
        if (rc != FANLIN_OK) return rc;
        out->pre = pre;
        out->pub.stages |= pre.color_op != COLOR_NONE ? 1u : 0u;
        out->pub.algorithmic_bytes = uint64_t(out->pub.src_x1 - out->pub.src_x0) * (out->pub.src_y1 - out->pub.src_y0) * job.src_channels * sample_bytes(pre.sample) + out->pub.out_bytes;
        return FANLIN_OK;
    }
    JobPlan p;
    const uint32_t W = job.src_w, H = job.src_h, c0 = job.src_channels;
    if (W == 0 || H == 0) { set_error("fanlin: empty source image"); return FANLIN_EINVAL; }
    if (c0 < 1 || c0 > 4) { set_error("fanlin: src_channels must be 1..4 (L/La/Rgb/Rgba)"); return FANLIN_EINVAL; }
    const uint32_t s0 = job.src_sample;
    if (s0 > SAMPLE_F32) { set_error("fanlin: src_sample must be 0 (u8), 1 (u16) or 2 (f32)"); return FANLIN_EINVAL; }
    if (s0 == SAMPLE_F32 && c0 < 3) { set_error("fanlin: f32 images are Rgb32F or Rgba32F (src_channels 3 or 4)"); return FANLIN_EINVAL; }
    const uint32_t bp0 = sample_bytes(s0);
    if (job.src_pitch != 0 && (job.src_pitch < W * c0 * bp0 || job.src_pitch % bp0 != 0)) { set_error("fanlin: src_pitch smaller than a row (or not a multiple of the subpixel size)"); return FANLIN_EINVAL; }
    if (s0 != SAMPLE_U8 && (reinterpret_cast<uintptr_t>(job.src) % bp0 != 0 || reinterpret_cast<uintptr_t>(job.dst) % bp0 != 0)) { set_error("fanlin: src / dst not aligned to the subpixel size"); return FANLIN_EINVAL; }
    if (job.filter != FANLIN_FILTER_NEAREST && job.filter != FANLIN_FILTER_LANCZOS3) {
        set_error("fanlin: unknown filter");
        return FANLIN_EINVAL;
    }
    if (!(job.blur_sigma >= 0.0f) || job.blur_sigma > 1000.0f) { set_error("fanlin: bad blur_sigma"); return FANLIN_EINVAL; }
    if (uint64_t(W) * H * c0 > (uint64_t(1) << 33)) { set_error("fanlin: source image too large"); return FANLIN_EINVAL; }

    // handler.rs:224-228 -- grayscale wins over inverse
    uint32_t op = COLOR_NONE, c1 = c0;
    if (job.flags & FANLIN_GRAYSCALE) {
        if (c0 >= 3) { op = COLOR_GRAY; c1 = s0 == SAMPLE_F32 ? c0 : c0 - 2; }  // grayscale_with_type: Rgb32F / Rgba32F keep their type
    } else if (job.flags & FANLIN_INVERSE) {
        op = COLOR_INVERT;
    }
    const uint32_t kind = job.filter == FANLIN_FILTER_NEAREST ? KIND_NEAREST : KIND_LANCZOS3;
    const bool want_rgba = (job.flags & FANLIN_TO_RGBA8) != 0;
    const bool blur = job.blur_sigma > 0.0f;
    const uint32_t fill = uint32_t(job.fill_rgb[0]) | uint32_t(job.fill_rgb[1]) << 8 | uint32_t(job.fill_rgb[2]) << 16 | 0xff000000u;

    uint32_t cur_w = W, cur_h = H;           // image after (optional) resize + crop
    uint32_t full_w = W, full_h = H;         // filtered size before crop
    uint32_t rx = 0, ry = 0;                 // crop origin
    bool resample = false, letterbox = false;
    uint32_t ov_x = 0, ov_y = 0;
    if (job.flags & FANLIN_HAS_DIMS) {
        const uint32_t rw = job.req_w, rh = job.req_h;
        if (rw == 0 || rh == 0 || rw > 65535 || rh > 65535) { set_error("fanlin: bad requested dimensions"); return FANLIN_EINVAL; }
        if (rw != W || rh != H) {  // handler.rs:231
            if (job.flags & FANLIN_CROP) {  // DynamicImage::resize_to_fill
                resize_dimensions(W, H, rw, rh, true, &full_w, &full_h);
                if (uint64_t(rw) * full_h > uint64_t(full_w) * rh) { rx = 0; ry = full_h > rh ? (full_h - rh) / 2 : 0; }
                else { rx = full_w > rw ? (full_w - rw) / 2 : 0; ry = 0; }
                rx = std::min(rx, full_w);
                ry = std::min(ry, full_h);
                cur_w = std::min(rw, full_w - rx);
                cur_h = std::min(rh, full_h - ry);
            } else {  // DynamicImage::resize
                resize_dimensions(W, H, rw, rh, false, &full_w, &full_h);
                cur_w = full_w;
                cur_h = full_h;
            }
            if (full_w > 65535 || full_h > 65535) { set_error("fanlin: resized dimensions too large"); return FANLIN_EINVAL; }
            resample = full_w != W || full_h != H;  // imageops::resize copies when the size is unchanged
        }
        if (rw > cur_w || rh > cur_h) {  // handler.rs:238
            letterbox = true;
            ov_x = absdiff(rw, cur_w) / 2;
            ov_y = absdiff(rh, cur_h) / 2;
        }
    }

    fanlin_plan &pub = p.pub;
    pub.resized_w = resample ? full_w : 0;
    pub.resized_h = resample ? full_h : 0;
    pub.crop_x = rx;
    pub.crop_y = ry;
    pub.overlay_x = ov_x;
    pub.overlay_y = ov_y;
    pub.stages = (op != COLOR_NONE ? 1u : 0u) | (resample ? 2u : 0u) | (letterbox ? 4u : 0u) | (blur ? 8u : 0u) | (want_rgba ? 16u : 0u);

    uint32_t img_w = cur_w, img_h = cur_h, img_c = c1, img_s = s0;  // running image description
    const bool subrect = !resample && (cur_w != W || cur_h != H);
    StagePlan &a = p.a;
    const bool need_a = resample || letterbox || subrect || !blur;
    if (need_a) {
        a.present = true;
        a.separable = resample;
        a.src_is_input = true;
        a.in_w = W; a.in_h = H; a.c_mem = c0; a.c = c1; a.color_op = op;
        a.s_in = a.s_out = s0;
        a.v_kind = a.h_kind = kind;
        a.v_out = full_h; a.h_out = full_w;
        a.oy0 = ry; a.n_rows = cur_h; a.ox0 = rx; a.n_cols = cur_w;
        if (letterbox) {
            a.canvas_w = job.req_w; a.canvas_h = job.req_h; a.c_out = 4; a.epi = EPI_BLEND_FILL;
            a.dst_x = ov_x; a.dst_y = ov_y; a.fill = fill;
            a.n_cols = std::min(a.n_cols, a.canvas_w - std::min(a.dst_x, a.canvas_w));
            a.n_rows = std::min(a.n_rows, a.canvas_h - std::min(a.dst_y, a.canvas_h));
            img_w = a.canvas_w; img_h = a.canvas_h; img_c = 4;
            a.s_out = img_s = SAMPLE_U8;  // the canvas is Rgba<u8> whatever the image (handler.rs:240)
        } else {
            a.canvas_w = cur_w; a.canvas_h = cur_h; a.dst_x = a.dst_y = 0;
            if (want_rgba && !blur) { a.epi = EPI_TO_RGBA; a.c_out = 4; img_c = 4; a.s_out = img_s = SAMPLE_U8; }
            else { a.epi = EPI_PLAIN; a.c_out = c1; }
        }
        if (resample) {
            axis_window(kind, 0.f, H, full_h, a.oy0, std::max(a.n_rows, 1u), &a.sy0, &a.n_sy);
            a.n_sy -= a.sy0;
            axis_window(kind, 0.f, W, full_w, a.ox0, std::max(a.n_cols, 1u), &a.sx0, &a.n_sx);
            a.n_sx -= a.sx0;
            if (with_tables) {
                a.vtab = build_axis_table(kind, 0.f, H, full_h);
                a.htab = build_axis_table(kind, 0.f, W, full_w);
            }
        } else {
            a.sx0 = a.ox0; a.n_sx = a.n_cols; a.sy0 = a.oy0; a.n_sy = a.n_rows;
        }
        pub.src_x0 = a.sx0; pub.src_x1 = a.sx0 + a.n_sx; pub.src_y0 = a.sy0; pub.src_y1 = a.sy0 + a.n_sy;
    } else {
        pub.src_x0 = 0; pub.src_x1 = W; pub.src_y0 = 0; pub.src_y1 = H;
    }
    if (blur) {  // handler.rs:250-255 -- runs last, on the letterboxed canvas
        StagePlan &b = p.b;
        b.present = true;
        b.separable = true;
        b.src_is_input = !a.present;
        b.in_w = img_w; b.in_h = img_h;
        if (a.present) { b.c_mem = img_c; b.c = img_c; b.color_op = COLOR_NONE; }
        else { b.c_mem = c0; b.c = c1; b.color_op = op; img_c = c1; }
        b.s_in = b.s_out = img_s;
        b.v_kind = b.h_kind = KIND_GAUSSIAN;
        b.sigma = job.blur_sigma;
        b.v_out = img_h; b.h_out = img_w;
        b.oy0 = 0; b.n_rows = img_h; b.ox0 = 0; b.n_cols = img_w;
        b.sx0 = 0; b.n_sx = img_w; b.sy0 = 0; b.n_sy = img_h;
        b.canvas_w = img_w; b.canvas_h = img_h; b.dst_x = b.dst_y = 0;
        if (want_rgba) { b.epi = EPI_TO_RGBA; b.c_out = 4; img_c = 4; b.s_out = img_s = SAMPLE_U8; }
        else { b.epi = EPI_PLAIN; b.c_out = b.c; }
        if (with_tables) {
            b.vtab = build_axis_table(KIND_GAUSSIAN, b.sigma, img_h, img_h);
            b.htab = build_axis_table(KIND_GAUSSIAN, b.sigma, img_w, img_w);
        }
    }
    if ((job.flags & FANLIN_TO_RGB8) && (img_c != 3 || img_s != SAMPLE_U8)) {  // DynamicImage::to_rgb8 as a last pass over the (small) output
        p.post_c_in = img_c;
        p.post_s_in = img_s;
        img_c = 3;
        img_s = SAMPLE_U8;
        pub.stages |= 32u;
    }
    if (job.flags & FANLIN_TO_YCBCR) {  // the JPEG encoder's rgb_to_ycbcr of to_rgb8() of the result, as three planes
        p.post_c_in = img_c;
        p.post_s_in = img_s;
        p.post_ycbcr = true;
        img_c = 3;
        img_s = SAMPLE_U8;
        pub.stages |= 64u;
    }
    pub.out_w = img_w;
    pub.out_h = img_h;
    pub.out_channels = img_c;
    pub.out_sample = img_s;
    pub.out_bytes = uint64_t(img_w) * img_h * img_c * sample_bytes(img_s);
    pub.algorithmic_bytes = uint64_t(pub.src_x1 - pub.src_x0) * (pub.src_y1 - pub.src_y0) * c0 * bp0 + pub.out_bytes;
    *out = std::move(p);
    return FANLIN_OK;
}

}  // namespace fanlin
