// The stage for DynamicImage variants with 16-bit and f32 subpixels (ImageLuma16 .. ImageRgba16 from 16-bit PNG / TIFF /
// PNM, ImageRgb32F / ImageRgba32F from HDR / EXR: what DynamicImage::from_decoder yields at reference src/handler.rs:219;
// SURVEY.md 8f rank 4).  In the reference they run through the same generic code as the u8 variants (:224-255):
// image-0.25.6 sample.rs is generic over the subpixel S, so these kernels keep the crate's operation order -- vertical
// pass into an unclamped f32 intermediate, horizontal pass, every tap a separately rounded multiply and add -- and are
// bit-identical to the CPU path:
//   sample.rs    t += (sample as f32) * w;  store NumCast::from(FloatNearest(clamp(t, S::MIN, S::MAX))):
//                [0, 65535] + round half away for u16, [0.0, 1.0] and no rounding for f32
//   color.rs     rgb_to_luma in S::Larger: u32 for u16, f64 for f32 (Rgb32F / Rgba32F keep their pixel type in
//                DynamicImage::grayscale: the luma is replicated); Invert: MAX - c, alpha kept
//   dynimage.rs  GenericImageView for DynamicImage (what overlay reads through), to_rgba8, to_rgb8:
//                FromPrimitive<u16> for u8 = (c + 128) / 257, FromPrimitive<f32> for u8 = round(clamp(c, 0, 1) * 255)
// The letterbox canvas is Rgba<u8> (handler.rs:240), so a blur behind it is a u8 stage and takes the u8 kernels.
// One thread per element; the tensor-core kernels stay u8-only (the north_star's path).
#include "device_common.cuh"
#include "kernels.h"

#include <algorithm>

namespace fanlin {

namespace {

constexpr int TX = 128;

// Pixel (x, y) of the stage input with the colour op applied: v[0..d.c) in the subpixel's own scale.
__device__ __forceinline__ void load_px_deep(const StageDesc &d, uint32_t x, uint32_t y, float v[4]) {
    const uint8_t *p = d.src + size_t(y) * d.src_pitch + size_t(x) * d.c_mem * sample_bytes(d.s_in);
    const uint32_t cm = d.c_mem;
    if (d.s_in == SAMPLE_F32) {
        const float *q = reinterpret_cast<const float *>(p);
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (uint32_t k = 0; k < cm; k++) s[k] = q[k];
        if (d.color_op == COLOR_GRAY) {  // c_mem is 3 or 4: ((2126 r + 7152 g) + 722 b) / 10000 in f64, `as f32`, replicated
            const double l = __dadd_rn(__dadd_rn(__dmul_rn(2126.0, double(s[0])), __dmul_rn(7152.0, double(s[1]))), __dmul_rn(722.0, double(s[2])));
            const float lf = __double2float_rn(__ddiv_rn(l, 10000.0));
            v[0] = v[1] = v[2] = lf;
            v[3] = s[3];
        } else if (d.color_op == COLOR_INVERT) {
            const uint32_t ncol = cm == 4 ? 3u : cm;
            for (uint32_t k = 0; k < cm; k++) v[k] = k < ncol ? __fsub_rn(1.0f, s[k]) : s[k];
        } else {
            for (uint32_t k = 0; k < cm; k++) v[k] = s[k];
        }
        return;
    }
    uint32_t s[4] = {0, 0, 0, 0};
    uint32_t mx;
    if (d.s_in == SAMPLE_U16) {
        const uint16_t *q = reinterpret_cast<const uint16_t *>(p);
        for (uint32_t k = 0; k < cm; k++) s[k] = q[k];
        mx = 65535u;
    } else {
        for (uint32_t k = 0; k < cm; k++) s[k] = p[k];
        mx = 255u;
    }
    if (d.color_op == COLOR_GRAY) {  // c_mem is 3 or 4; u32 holds 10000 * 65535
        v[0] = float((2126u * s[0] + 7152u * s[1] + 722u * s[2]) / 10000u);
        if (cm == 4) v[1] = float(s[3]);
    } else if (d.color_op == COLOR_INVERT) {
        const uint32_t ncol = (cm == 2 || cm == 4) ? cm - 1 : cm;
        for (uint32_t k = 0; k < cm; k++) v[k] = float(k < ncol ? mx - s[k] : s[k]);
    } else {
        for (uint32_t k = 0; k < cm; k++) v[k] = float(s[k]);
    }
}

// NumCast::from(FloatNearest(clamp(t, S::DEFAULT_MIN_VALUE, S::DEFAULT_MAX_VALUE))) of horizontal_sample.
__device__ __forceinline__ float finish_sample(uint32_t s, float t) {
    if (s == SAMPLE_F32) return t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    const float mx = s == SAMPLE_U16 ? 65535.0f : 255.0f;
    t = t < 0.0f ? 0.0f : (t > mx ? mx : t);
    return roundf(t);  // half away from zero
}

// FromPrimitive<S> for u8 of a subpixel value in its own scale.
__device__ __forceinline__ uint32_t sub_to_u8(uint32_t s, float v) {
    if (s == SAMPLE_U8) return uint32_t(v);
    if (s == SAMPLE_U16) return (uint32_t(v) + 128u) / 257u;
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    return uint32_t(roundf(__fmul_rn(v, 255.0f)));
}

// Writes one pixel (values v[0..d.c) of subpixel type d.s_in) to canvas position (cx, cy): as it is (EPI_PLAIN, the
// canvas has the same subpixel type), or viewed as Rgba<u8> -- to_rgba().into_color() -- and blended onto the fill
// colour / stored as RGBA8.
__device__ __forceinline__ void store_px_deep(const StageDesc &d, uint32_t cx, uint32_t cy, const float v[4]) {
    if (d.epi == EPI_PLAIN) {
        uint8_t *q = d.dst + size_t(cy) * d.dst_pitch + size_t(cx) * d.c_out * sample_bytes(d.s_out);
        if (d.s_out == SAMPLE_F32) { float *o = reinterpret_cast<float *>(q); for (uint32_t k = 0; k < d.c; k++) o[k] = v[k]; }
        else if (d.s_out == SAMPLE_U16) { uint16_t *o = reinterpret_cast<uint16_t *>(q); for (uint32_t k = 0; k < d.c; k++) o[k] = uint16_t(v[k]); }
        else { for (uint32_t k = 0; k < d.c; k++) q[k] = uint8_t(v[k]); }
        return;
    }
    uint32_t b[4];
#pragma unroll
    for (int k = 0; k < 4; k++) b[k] = sub_to_u8(d.s_in, v[k]);
    uint32_t px = to_rgba_packed(b, d.c);  // missing alpha: MAX -> 255
    if ((d.epi & EPI_MASK) == EPI_BLEND_FILL) px = blend_rgba(d.fill, px);
    if (d.epi & EPI_RGB8) {
        uint8_t *q = d.dst + size_t(cy) * d.dst_pitch + size_t(cx) * 3;
        q[0] = uint8_t(px); q[1] = uint8_t(px >> 8); q[2] = uint8_t(px >> 16);
        return;
    }
    store_rgba(d.dst + size_t(cy) * d.dst_pitch + size_t(cx) * 4, px);
}

// vertical_sample: one thread per (produced row, source column) of the f32 intermediate.
__global__ void __launch_bounds__(TX) vpass_deep_kernel(const StageDesc *__restrict__ descs, const TapEntry *__restrict__ tab,
                                                        const float *__restrict__ tw) {
    const StageDesc d = descs[blockIdx.y];
    const uint32_t xtiles = (d.n_sx + TX - 1) / TX;
    if (blockIdx.x >= d.n_rows * xtiles) return;
    const uint32_t r = blockIdx.x / xtiles;
    const uint32_t x = (blockIdx.x % xtiles) * TX + threadIdx.x;
    if (x >= d.n_sx) return;
    const TapEntry e = tab[d.v_tab + d.oy0 + r];
    const float *w = tw + e.woff;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (uint32_t i = 0; i < e.count; i++) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        load_px_deep(d, d.sx0 + x, e.left + i, v);
        const float wi = w[i];
#pragma unroll
        for (int k = 0; k < 4; k++) acc[k] = __fadd_rn(acc[k], __fmul_rn(v[k], wi));
    }
    float *o = d.tmp + size_t(r) * d.tmp_pitch + size_t(x) * d.c;
    for (uint32_t k = 0; k < d.c; k++) o[k] = acc[k];
}

// horizontal_sample + epilogue: one thread per canvas pixel (the fill colour outside the placed rectangle).
__global__ void __launch_bounds__(TX) hpass_deep_kernel(const StageDesc *__restrict__ descs, const TapEntry *__restrict__ tab,
                                                        const float *__restrict__ tw) {
    const StageDesc d = descs[blockIdx.y];
    const uint32_t xtiles = (d.canvas_w + TX - 1) / TX;
    if (blockIdx.x >= d.canvas_h * xtiles) return;
    const uint32_t cy = blockIdx.x / xtiles;
    const uint32_t cx = (blockIdx.x % xtiles) * TX + threadIdx.x;
    if (cx >= d.canvas_w) return;
    const uint32_t lx = cx - d.dst_x, ly = cy - d.dst_y;
    if (cx < d.dst_x || cy < d.dst_y || lx >= d.n_cols || ly >= d.n_rows) {
        store_fill(d, cx, cy);
        return;
    }
    const TapEntry e = tab[d.h_tab + d.ox0 + lx];
    const float *w = tw + e.woff;
    const float *t = d.tmp + size_t(ly) * d.tmp_pitch + size_t(e.left - d.sx0) * d.c;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (uint32_t i = 0; i < e.count; i++) {
        const float wi = w[i];
        for (uint32_t k = 0; k < d.c; k++) acc[k] = __fadd_rn(acc[k], __fmul_rn(t[size_t(i) * d.c + k], wi));
    }
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = finish_sample(d.s_in, acc[k]);
    store_px_deep(d, cx, cy, v);
}

// Stages without a resample: colour op, crop copy, letterbox, to_rgba8.
__global__ void __launch_bounds__(TX) compose_deep_kernel(const StageDesc *__restrict__ descs) {
    const StageDesc d = descs[blockIdx.y];
    const uint32_t xtiles = (d.canvas_w + TX - 1) / TX;
    if (blockIdx.x >= d.canvas_h * xtiles) return;
    const uint32_t cy = blockIdx.x / xtiles;
    const uint32_t cx = (blockIdx.x % xtiles) * TX + threadIdx.x;
    if (cx >= d.canvas_w) return;
    const uint32_t lx = cx - d.dst_x, ly = cy - d.dst_y;
    if (cx < d.dst_x || cy < d.dst_y || lx >= d.n_cols || ly >= d.n_rows) {
        store_fill(d, cx, cy);
        return;
    }
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    load_px_deep(d, d.ox0 + lx, d.oy0 + ly, v);
    store_px_deep(d, cx, cy, v);
}

// EXIF orientation (+ colour op) of the stored image into scratch: oriented rows [oy0, oy0 + n_rows) -> dst rows
// [0, n_rows); the index map of orient_pass_kernel (kernels_exact.cu).  One thread per oriented pixel.
__global__ void __launch_bounds__(TX) orient_deep_kernel(const StageDesc *__restrict__ descs) {
    const StageDesc d = descs[blockIdx.y];
    const uint32_t ow = d.canvas_w;
    const uint32_t xtiles = (ow + TX - 1) / TX;
    if (blockIdx.x >= d.n_rows * xtiles) return;
    const uint32_t row = blockIdx.x / xtiles, xo = (blockIdx.x % xtiles) * TX + threadIdx.x;
    if (xo >= ow) return;
    const uint32_t yo = d.oy0 + row, o = d.orient, W = d.src_w, H = d.src_h;
    uint32_t sx, sy;
    if (o >= 5) {
        sx = (o == 5 || o == 6) ? yo : W - 1 - yo;
        sy = (o == 5 || o == 8) ? xo : H - 1 - xo;
    } else {
        sx = (o == 2 || o == 3) ? W - 1 - xo : xo;
        sy = (o == 3 || o == 4) ? H - 1 - yo : yo;
    }
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    load_px_deep(d, sx, sy, v);
    store_px_deep(d, xo, row, v);  // EPI_PLAIN
}

// DynamicImage::to_rgb8 of a final image with 16-bit / f32 subpixels: tight src [h][w][c_mem] -> dst [h][w][3] u8.
__global__ void __launch_bounds__(256) to_rgb8_deep_kernel(const StageDesc *__restrict__ descs) {
    const StageDesc &d = descs[blockIdx.y];
    const uint32_t n = d.canvas_w * d.canvas_h, c = d.c_mem, s = d.s_in;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        float v[3];
        for (uint32_t k = 0; k < 3; k++) {
            const size_t idx = size_t(i) * c + (c <= 2 ? 0 : k);
            v[k] = s == SAMPLE_F32 ? reinterpret_cast<const float *>(d.src)[idx]
                 : s == SAMPLE_U16 ? float(reinterpret_cast<const uint16_t *>(d.src)[idx]) : float(d.src[idx]);
        }
        uint8_t *o = d.dst + size_t(i) * 3;
        if (d.epi) { rgb_to_ycbcr_u8(sub_to_u8(s, v[0]), sub_to_u8(s, v[1]), sub_to_u8(s, v[2]), d.dst + i, d.dst + n + i, d.dst + 2 * size_t(n) + i); continue; }
        o[0] = uint8_t(sub_to_u8(s, v[0])); o[1] = uint8_t(sub_to_u8(s, v[1])); o[2] = uint8_t(sub_to_u8(s, v[2]));
    }
}

}  // namespace

int launch_sep_deep(const StageDesc *d_descs, const TapEntry *d_tab, const float *d_w, const LaunchGeom &g, LaunchCtx &lc) {
    if (g.n_jobs == 0) return 0;
    const uint32_t vx = g.max_n_rows * ((g.max_n_sx + TX - 1) / TX);
    const uint32_t hx = g.max_canvas_h * ((g.max_canvas_w + TX - 1) / TX);
    int n = 0;
    if (vx) {
        lc.begin("vpass_deep_kernel");
        vpass_deep_kernel<<<dim3(vx, g.n_jobs), TX, 0, lc.st>>>(d_descs, d_tab, d_w);
        lc.end();
        n++;
    }
    if (hx) {
        lc.begin("hpass_deep_kernel");
        hpass_deep_kernel<<<dim3(hx, g.n_jobs), TX, 0, lc.st>>>(d_descs, d_tab, d_w);
        lc.end();
        n++;
    }
    return n;
}

int launch_compose_deep(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc) {
    if (g.n_jobs == 0) return 0;
    const uint32_t hx = g.max_canvas_h * ((g.max_canvas_w + TX - 1) / TX);
    if (!hx) return 0;
    lc.begin("compose_deep_kernel");
    compose_deep_kernel<<<dim3(hx, g.n_jobs), TX, 0, lc.st>>>(d_descs);
    lc.end();
    return 1;
}

int launch_orient_deep(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc) {
    if (g.n_jobs == 0 || !g.max_canvas_w || !g.max_canvas_h) return 0;
    lc.begin("orient_deep_kernel");
    orient_deep_kernel<<<dim3(g.max_canvas_h * ((g.max_canvas_w + TX - 1) / TX), g.n_jobs), TX, 0, lc.st>>>(d_descs);
    lc.end();
    return 1;
}

int launch_to_rgb8_deep(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc) {
    if (g.n_jobs == 0 || !g.max_canvas_w || !g.max_canvas_h) return 0;
    const uint32_t blocks = std::min<uint32_t>(1024, (g.max_canvas_w * g.max_canvas_h + 255) / 256);
    lc.begin("to_rgb8_deep_kernel");
    to_rgb8_deep_kernel<<<dim3(blocks, g.n_jobs), 256, 0, lc.st>>>(d_descs);
    lc.end();
    return 1;
}

}  // namespace fanlin
