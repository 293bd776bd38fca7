/*
 * fanlin_oracle.c -- CPU restatement of fanlin-rs's pixel-transform stage.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fanlin-rs_b200/ may include, link or
 * call this file; it exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs can check and time the CUDA
 * path against the reference's arithmetic.
 *
 * PARITY UNPINNED.  The arithmetic of this path is not in /root/reference: it
 * lives in the third-party crate `image` = 0.25.6 (reference Cargo.toml:14,
 * Cargo.lock:1948-1951, checksum db35664c...251a), which is not vendored, and
 * there is no Rust toolchain in this image, so the reference cannot be run to
 * produce golden pixels.  The reference's own tests assert no pixel value
 * (src/main.rs:457-468).  This file restates the crate's published algorithm
 * (imageops/sample.rs, imageops/mod.rs, imageops/colorops.rs, color.rs,
 * math/utils.rs, dynimage.rs of image 0.25.6) and the reference's sequencing
 * (src/handler.rs:224-255 for stills, :329-355 for GIF frames).  What pins it
 * instead: the crate's upstream known-answer tests for resize_dimensions,
 * an independent numpy restatement of the same spec (oracle/np_restatement.py)
 * that must agree bit for bit, a Pillow LANCZOS structural check, and analytic
 * invariants -- see tests/test_oracle.py.
 *
 * All float arithmetic is IEEE binary32 with separately rounded multiply and
 * add, in the crate's evaluation order (Rust never contracts a*b+c); compile
 * with -ffp-contract=off.  sinf/expf are glibc's, which is what Rust's
 * f32::sin / f32::exp call on the reference's linux-gnu target.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FO_OK 0
#define FO_EINVAL 1
#define FO_ENOMEM 2
#define FO_ECAP 3

enum { FO_NEAREST = 0, FO_LANCZOS3 = 1, FO_GAUSSIAN_BLUR = 100 };

enum {
    FO_GRAYSCALE = 1u << 0,
    FO_INVERSE = 1u << 1,
    FO_HAS_DIMS = 1u << 2,
    FO_CROP = 1u << 3,
    FO_GIF_FRAME = 1u << 4,
    FO_TO_RGB8 = 1u << 5 /* still handed to the JPEG encoder: DynamicImage::to_rgb8 (handler.rs:274-278 drops alpha) */
};

/* One request, the fields of query::Query the stage reads (src/query.rs:28-70)
 * plus the decoded image. */
typedef struct {
    const uint8_t *src;
    uint32_t src_w, src_h, src_c; /* 1=L8 2=La8 3=Rgb8 4=Rgba8 */
    uint32_t flags;
    uint32_t req_w, req_h;
    uint8_t fill[3];
    uint8_t orientation; /* EXIF orientation of the decoded still, 0 / 1 = none, 2..8 (handler.rs:206,221-223) */
    float blur_sigma; /* already through Query::blur(): 0 or clamp(.,10,20) */
    uint8_t *dst;
    uint64_t dst_cap;
    /* outputs */
    uint32_t out_w, out_h, out_c;
} fo_job;

/* ---- math/utils.rs: resize_dimensions ---------------------------------- */

static uint64_t round_to_u64(double v) {
    /* f64::round (half away from zero) then `as u64` (saturating, NaN -> 0) */
    double r = round(v);
    if (!(r > 0.0)) return 0;
    if (r >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)r;
}
static uint32_t round_to_u32(double v) {
    double r = round(v);
    if (!(r > 0.0)) return 0;
    if (r >= 4294967295.0) return UINT32_MAX;
    return (uint32_t)r;
}

void fo_resize_dimensions(uint32_t width, uint32_t height, uint32_t nwidth, uint32_t nheight,
                          int fill, uint32_t *ow, uint32_t *oh) {
    double wratio = (double)nwidth / (double)width;
    double hratio = (double)nheight / (double)height;
    double ratio = fill ? fmax(wratio, hratio) : fmin(wratio, hratio);
    uint64_t nw = round_to_u64((double)width * ratio);
    uint64_t nh = round_to_u64((double)height * ratio);
    if (nw < 1) nw = 1;
    if (nh < 1) nh = 1;
    if (nw > (uint64_t)UINT32_MAX) {
        double r2 = (double)UINT32_MAX / (double)width;
        uint32_t h2 = round_to_u32((double)height * r2);
        *ow = UINT32_MAX;
        *oh = h2 < 1 ? 1 : h2;
    } else if (nh > (uint64_t)UINT32_MAX) {
        double r2 = (double)UINT32_MAX / (double)height;
        uint32_t w2 = round_to_u32((double)width * r2);
        *ow = w2 < 1 ? 1 : w2;
        *oh = UINT32_MAX;
    } else {
        *ow = (uint32_t)nw;
        *oh = (uint32_t)nh;
    }
}

/* ---- imageops/sample.rs: kernels --------------------------------------- */

static const float PI_F = 3.14159265358979323846f;

static float sinc_f(float t) {
    float a = t * PI_F;
    if (t == 0.0f) return 1.0f;
    return sinf(a) / a;
}
static float lanczos3_kernel(float x) {
    if (fabsf(x) < 3.0f) return sinc_f(x) * sinc_f(x / 3.0f);
    return 0.0f;
}
static float gaussian_f(float x, float r) {
    float lhs = 1.0f / (sqrtf(2.0f * PI_F) * r);
    float rhs = expf(-(x * x) / (2.0f * (r * r)));
    return lhs * rhs;
}

typedef struct {
    int kind;
    float support;
    float sigma;
} fo_filter;

static float filter_eval(const fo_filter *f, float x) {
    switch (f->kind) {
    case FO_NEAREST: return 1.0f; /* box_kernel */
    case FO_LANCZOS3: return lanczos3_kernel(x);
    default: return gaussian_f(x, f->sigma);
    }
}
static int make_filter(int kind, float sigma, fo_filter *f) {
    f->kind = kind;
    f->sigma = sigma;
    if (kind == FO_NEAREST) f->support = 0.0f;
    else if (kind == FO_LANCZOS3) f->support = 3.0f;
    else if (kind == FO_GAUSSIAN_BLUR) f->support = 2.0f * sigma;
    else return FO_EINVAL;
    return FO_OK;
}

static int64_t f32_to_i64(float v) {
    if (v != v) return 0;
    if (v >= 9223372036854775807.0f) return INT64_MAX;
    if (v <= -9223372036854775808.0f) return INT64_MIN;
    return (int64_t)v;
}
static int64_t clamp_i64(int64_t a, int64_t lo, int64_t hi) {
    if (a < lo) return lo;
    if (a > hi) return hi;
    return a;
}

/* The per-output tap window and normalised weights, shared by both passes
 * (sample.rs horizontal_sample / vertical_sample preambles).  Returns the tap
 * count; ws must hold n_in floats at most. */
static uint32_t tap_window(const fo_filter *f, uint32_t n_in, uint32_t n_out, uint32_t o,
                           uint32_t *left_out, float *ws) {
    float ratio = (float)n_in / (float)n_out;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float src_support = f->support * sratio;
    float inputx = ((float)o + 0.5f) * ratio;
    int64_t left = f32_to_i64(floorf(inputx - src_support));
    left = clamp_i64(left, 0, (int64_t)n_in - 1);
    int64_t right = f32_to_i64(ceilf(inputx + src_support));
    right = clamp_i64(right, left + 1, (int64_t)n_in);
    inputx = inputx - 0.5f;
    float sum = 0.0f;
    uint32_t n = 0;
    for (int64_t i = left; i < right; i++) {
        float w = filter_eval(f, ((float)i - inputx) / sratio);
        ws[n++] = w;
        sum += w;
    }
    for (uint32_t k = 0; k < n; k++) ws[k] /= sum;
    *left_out = (uint32_t)left;
    return n;
}

/* Exposed for tests and for checking the product's weight tables.
 * lefts/counts: n_out entries; weights: n_out * max_taps floats, row o holds
 * counts[o] weights.  Returns max taps needed when weights == NULL. */
uint32_t fo_weight_table(int kind, float sigma, uint32_t n_in, uint32_t n_out, uint32_t *lefts,
                         uint32_t *counts, float *weights, uint32_t max_taps) {
    fo_filter f;
    if (make_filter(kind, sigma, &f) != FO_OK || n_in == 0 || n_out == 0) return 0;
    float *ws = (float *)malloc(sizeof(float) * (size_t)n_in);
    if (!ws) return 0;
    uint32_t mx = 0;
    for (uint32_t o = 0; o < n_out; o++) {
        uint32_t left;
        uint32_t n = tap_window(&f, n_in, n_out, o, &left, ws);
        if (n > mx) mx = n;
        if (lefts) lefts[o] = left;
        if (counts) counts[o] = n;
        if (weights && n <= max_taps) memcpy(weights + (size_t)o * max_taps, ws, sizeof(float) * n);
    }
    free(ws);
    return mx;
}

/* sample.rs vertical_sample: u8 (w x h x c) -> f32 (w x new_h x c), unclamped. */
static int vertical_sample(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, uint32_t new_h,
                           const fo_filter *f, float *out) {
    float *ws = (float *)malloc(sizeof(float) * (size_t)h);
    if (!ws) return FO_ENOMEM;
    size_t row = (size_t)w * c;
    for (uint32_t oy = 0; oy < new_h; oy++) {
        uint32_t left;
        uint32_t n = tap_window(f, h, new_h, oy, &left, ws);
        float *o = out + (size_t)oy * row;
        for (size_t x = 0; x < row; x++) {
            float t = 0.0f;
            const uint8_t *p = src + (size_t)left * row + x;
            for (uint32_t i = 0; i < n; i++) t += (float)p[(size_t)i * row] * ws[i];
            o[x] = t;
        }
    }
    free(ws);
    return FO_OK;
}

static uint8_t store_u8(float t) {
    /* clamp(t, 0, 255) then FloatNearest: f32::round, half away from zero */
    if (t < 0.0f) t = 0.0f;
    else if (t > 255.0f) t = 255.0f;
    return (uint8_t)roundf(t);
}

/* sample.rs horizontal_sample: f32 (w x h x c) -> u8 (new_w x h x c). */
static int horizontal_sample(const float *tmp, uint32_t w, uint32_t h, uint32_t c, uint32_t new_w,
                             const fo_filter *f, uint8_t *out) {
    float *ws = (float *)malloc(sizeof(float) * (size_t)w);
    if (!ws) return FO_ENOMEM;
    for (uint32_t ox = 0; ox < new_w; ox++) {
        uint32_t left;
        uint32_t n = tap_window(f, w, new_w, ox, &left, ws);
        for (uint32_t y = 0; y < h; y++) {
            const float *p = tmp + ((size_t)y * w + left) * c;
            uint8_t *o = out + ((size_t)y * new_w + ox) * c;
            for (uint32_t ch = 0; ch < c; ch++) {
                float t = 0.0f;
                for (uint32_t i = 0; i < n; i++) t += p[(size_t)i * c + ch] * ws[i];
                o[ch] = store_u8(t);
            }
        }
    }
    free(ws);
    return FO_OK;
}

/* imageops::resize (sample.rs): same size -> copy; else vertical then horizontal. */
int fo_resize(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, uint32_t nw, uint32_t nh,
              int kind, uint8_t *dst) {
    if (w == 0 || h == 0 || c < 1 || c > 4 || nw == 0 || nh == 0) return FO_EINVAL;
    if (nw == w && nh == h) {
        memcpy(dst, src, (size_t)w * h * c);
        return FO_OK;
    }
    fo_filter f;
    if (kind != FO_NEAREST && kind != FO_LANCZOS3) return FO_EINVAL;
    make_filter(kind, 0.0f, &f);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)w * nh * c);
    if (!tmp) return FO_ENOMEM;
    int rc = vertical_sample(src, w, h, c, nh, &f, tmp);
    if (rc == FO_OK) rc = horizontal_sample(tmp, w, nh, c, nw, &f, dst);
    free(tmp);
    return rc;
}

/* imageops::blur (sample.rs, 0.25.6 classic form): gaussian, support 2*sigma. */
int fo_blur(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, float sigma, uint8_t *dst) {
    if (w == 0 || h == 0 || c < 1 || c > 4) return FO_EINVAL;
    if (sigma <= 0.0f) sigma = 1.0f;
    fo_filter f;
    make_filter(FO_GAUSSIAN_BLUR, sigma, &f);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)w * h * c);
    if (!tmp) return FO_ENOMEM;
    int rc = vertical_sample(src, w, h, c, h, &f, tmp);
    if (rc == FO_OK) rc = horizontal_sample(tmp, w, h, c, w, &f, dst);
    free(tmp);
    return rc;
}

/* ---- color.rs / colorops.rs -------------------------------------------- */

static uint8_t rgb_to_luma(uint8_t r, uint8_t g, uint8_t b) {
    uint32_t l = 2126u * r + 7152u * g + 722u * b;
    return (uint8_t)(l / 10000u);
}

/* DynamicImage::grayscale: L8->L8, La8->La8, Rgb8->L8, Rgba8->La8.  Returns out channels. */
uint32_t fo_grayscale(const uint8_t *src, size_t npix, uint32_t c, uint8_t *dst) {
    if (c == 1 || c == 2) {
        memcpy(dst, src, npix * c);
        return c;
    }
    if (c == 3) {
        for (size_t i = 0; i < npix; i++) dst[i] = rgb_to_luma(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
        return 1;
    }
    for (size_t i = 0; i < npix; i++) {
        dst[2 * i] = rgb_to_luma(src[4 * i], src[4 * i + 1], src[4 * i + 2]);
        dst[2 * i + 1] = src[4 * i + 3];
    }
    return 2;
}

/* DynamicImage::invert: colour channels 255-v, alpha kept, in place. */
void fo_invert(uint8_t *px, size_t npix, uint32_t c) {
    uint32_t ncol = (c == 2 || c == 4) ? c - 1 : c;
    for (size_t i = 0; i < npix; i++)
        for (uint32_t k = 0; k < ncol; k++) px[i * c + k] = (uint8_t)(255 - px[i * c + k]);
}

static void to_rgba(const uint8_t *p, uint32_t c, uint8_t o[4]) {
    switch (c) {
    case 1: o[0] = o[1] = o[2] = p[0]; o[3] = 255; break;
    case 2: o[0] = o[1] = o[2] = p[0]; o[3] = p[1]; break;
    case 3: o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = 255; break;
    default: o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3]; break;
    }
}

/* DynamicImage::to_rgba8 */
/* DynamicImage::to_rgb8: L -> (l,l,l), La -> (l,l,l), Rgb -> itself, Rgba -> (r,g,b); alpha is dropped, not blended */
void fo_to_rgb8(const uint8_t *src, size_t npix, uint32_t c, uint8_t *dst) {
    for (size_t i = 0; i < npix; i++) {
        const uint8_t *p = src + i * c;
        uint8_t *o = dst + i * 3;
        if (c <= 2) { o[0] = o[1] = o[2] = p[0]; }
        else { o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; }
    }
}

void fo_to_rgba8(const uint8_t *src, size_t npix, uint32_t c, uint8_t *dst) {
    for (size_t i = 0; i < npix; i++) to_rgba(src + i * c, c, dst + 4 * i);
}

/* num-traits NumCast f32 -> u8 as used by Blend: truncation; values the cast
 * would reject (>= 256 or <= -1, NaN) would panic upstream -- saturate here and
 * let the tests assert they never occur. */
static uint8_t cast_u8_trunc(float v) {
    if (!(v > -1.0f)) return 0;
    if (v >= 256.0f) return 255;
    return (uint8_t)v;
}

/* color.rs: impl Blend for Rgba<u8> */
static void blend_rgba(uint8_t bg[4], const uint8_t fg[4]) {
    if (fg[3] == 0) return;
    if (fg[3] == 255) {
        memcpy(bg, fg, 4);
        return;
    }
    const float max_t = 255.0f;
    float bg_r = (float)bg[0] / max_t, bg_g = (float)bg[1] / max_t, bg_b = (float)bg[2] / max_t,
          bg_a = (float)bg[3] / max_t;
    float fg_r = (float)fg[0] / max_t, fg_g = (float)fg[1] / max_t, fg_b = (float)fg[2] / max_t,
          fg_a = (float)fg[3] / max_t;
    float alpha_final = bg_a + fg_a - bg_a * fg_a;
    if (alpha_final == 0.0f) return;
    float bg_r_a = bg_r * bg_a, bg_g_a = bg_g * bg_a, bg_b_a = bg_b * bg_a;
    float fg_r_a = fg_r * fg_a, fg_g_a = fg_g * fg_a, fg_b_a = fg_b * fg_a;
    float out_r_a = fg_r_a + bg_r_a * (1.0f - fg_a);
    float out_g_a = fg_g_a + bg_g_a * (1.0f - fg_a);
    float out_b_a = fg_b_a + bg_b_a * (1.0f - fg_a);
    float out_r = out_r_a / alpha_final, out_g = out_g_a / alpha_final, out_b = out_b_a / alpha_final;
    bg[0] = cast_u8_trunc(max_t * out_r);
    bg[1] = cast_u8_trunc(max_t * out_g);
    bg[2] = cast_u8_trunc(max_t * out_b);
    bg[3] = cast_u8_trunc(max_t * alpha_final);
}

/* imageops::overlay with overlay_bounds_ext clipping; bottom is RGBA8, top is
 * any u8 variant viewed through to_rgba (DynamicImage as GenericImageView). */
void fo_overlay(uint8_t *bottom, uint32_t bw, uint32_t bh, const uint8_t *top, uint32_t tw,
                uint32_t th, uint32_t tc, int64_t x, int64_t y) {
    if (x > (int64_t)bw || y > (int64_t)bh || x + (int64_t)tw <= 0 || y + (int64_t)th <= 0) return;
    int64_t max_x = x + (int64_t)tw < (int64_t)bw ? x + (int64_t)tw : (int64_t)bw;
    int64_t max_y = y + (int64_t)th < (int64_t)bh ? y + (int64_t)th : (int64_t)bh;
    int64_t ob_x = x > 0 ? x : 0, ob_y = y > 0 ? y : 0;
    int64_t ot_x = x < 0 ? -x : 0, ot_y = y < 0 ? -y : 0;
    int64_t rw = max_x - ob_x, rh = max_y - ob_y;
    for (int64_t j = 0; j < rh; j++)
        for (int64_t i = 0; i < rw; i++) {
            uint8_t p[4];
            to_rgba(top + ((size_t)(ot_y + j) * tw + (size_t)(ot_x + i)) * tc, tc, p);
            blend_rgba(bottom + ((size_t)(ob_y + j) * bw + (size_t)(ob_x + i)) * 4, p);
        }
}

/* imageops::crop + crop_dimms, copying the sub-rectangle (to_image). */
static void crop_copy(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, uint32_t x,
                      uint32_t y, uint32_t cw, uint32_t ch, uint8_t *dst, uint32_t *ow,
                      uint32_t *oh) {
    if (x > w) x = w;
    if (y > h) y = h;
    if (ch > h - y) ch = h - y;
    if (cw > w - x) cw = w - x;
    for (uint32_t j = 0; j < ch; j++)
        memcpy(dst + (size_t)j * cw * c, src + ((size_t)(y + j) * w + x) * c, (size_t)cw * c);
    *ow = cw;
    *oh = ch;
}

/* ---- the stage: src/handler.rs:224-255 (still) and :329-355 (GIF frame) -- */

/* ---- metadata.rs Orientation::from_exif + dynimage.rs apply_orientation ---------------
 * (handler.rs:221-223).  EXIF 1 NoTransforms, 2 FlipHorizontal, 3 Rotate180, 4 FlipVertical,
 * 5 Rotate90FlipH, 6 Rotate90, 7 Rotate270FlipH, 8 Rotate270.  The crate composes them from
 * imageops::{rotate90, rotate180, rotate270, flip_horizontal, flip_vertical}:
 *   rotate90 (clockwise):  out.put_pixel(h - 1 - y, x, p)   -> out is h x w
 *   rotate270:             out.put_pixel(y, w - 1 - x, p)   -> out is h x w
 *   rotate180:             out.put_pixel(w - 1 - x, h - 1 - y, p)
 *   flip_horizontal:       out.put_pixel(w - 1 - x, y, p);   flip_vertical: (x, h - 1 - y)
 * Rotate90FlipH = rotate90 then flip_horizontal; Rotate270FlipH = rotate270 then flip_horizontal. */
static void px_copy(uint8_t *d, const uint8_t *s, uint32_t c) { for (uint32_t k = 0; k < c; k++) d[k] = s[k]; }

static void op_rotate90(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, uint8_t *dst) { /* dst: h x w */
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) px_copy(dst + ((size_t)x * h + (h - 1 - y)) * c, src + ((size_t)y * w + x) * c, c);
}
static void op_rotate270(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, uint8_t *dst) { /* dst: h x w */
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) px_copy(dst + ((size_t)(w - 1 - x) * h + y) * c, src + ((size_t)y * w + x) * c, c);
}
static void op_rotate180(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, uint8_t *dst) {
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) px_copy(dst + ((size_t)(h - 1 - y) * w + (w - 1 - x)) * c, src + ((size_t)y * w + x) * c, c);
}
static void op_fliph(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, uint8_t *dst) {
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) px_copy(dst + ((size_t)y * w + (w - 1 - x)) * c, src + ((size_t)y * w + x) * c, c);
}
static void op_flipv(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, uint8_t *dst) {
    for (uint32_t y = 0; y < h; y++) memcpy(dst + (size_t)(h - 1 - y) * w * c, src + (size_t)y * w * c, (size_t)w * c);
}

/* dst holds w*h*c bytes; *ow, *oh receive the oriented size.  Returns FO_EINVAL for exif > 8. */
int fo_apply_orientation(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, uint32_t exif, uint8_t *dst, uint32_t *ow,
                         uint32_t *oh) {
    size_t n = (size_t)w * h * c;
    *ow = w; *oh = h;
    if (exif > 8) return FO_EINVAL;
    if (exif <= 1) { memcpy(dst, src, n); return FO_OK; }
    if (exif == 2) { op_fliph(src, w, h, c, dst); return FO_OK; }
    if (exif == 3) { op_rotate180(src, w, h, c, dst); return FO_OK; }
    if (exif == 4) { op_flipv(src, w, h, c, dst); return FO_OK; }
    *ow = h; *oh = w;
    if (exif == 6) { op_rotate90(src, w, h, c, dst); return FO_OK; }
    if (exif == 8) { op_rotate270(src, w, h, c, dst); return FO_OK; }
    uint8_t *t = (uint8_t *)malloc(n);
    if (!t) return FO_ENOMEM;
    if (exif == 5) op_rotate90(src, w, h, c, t); else op_rotate270(src, w, h, c, t); /* 5, 7: then flip_horizontal */
    op_fliph(t, h, w, c, dst);
    free(t);
    return FO_OK;
}

int fo_process(fo_job *job) {
    uint32_t w = job->src_w, h = job->src_h, c = job->src_c;
    if (!job->src || w == 0 || h == 0 || c < 1 || c > 4) return FO_EINVAL;
    int gif = (job->flags & FO_GIF_FRAME) != 0;
    if (gif && c != 4) return FO_EINVAL; /* frames are composited RGBA8 (handler.rs:328) */
    int filter = gif ? FO_NEAREST : FO_LANCZOS3; /* handler.rs:233,235 vs :338,:340 */
    uint8_t *img = (uint8_t *)malloc((size_t)w * h * c);
    if (!img) return FO_ENOMEM;
    int rc = FO_OK;
    if (!gif && job->orientation > 1) { /* handler.rs:221-223: stills only; process_gif never looks at EXIF */
        rc = fo_apply_orientation(job->src, w, h, c, job->orientation, img, &w, &h);
        if (rc != FO_OK) { free(img); return rc; }
    } else {
        memcpy(img, job->src, (size_t)w * h * c);
    }

    /* handler.rs:224-228 / :329-333 -- grayscale wins, inverse only otherwise */
    if (job->flags & FO_GRAYSCALE) {
        uint8_t *g = (uint8_t *)malloc((size_t)w * h * c);
        if (!g) { free(img); return FO_ENOMEM; }
        c = fo_grayscale(img, (size_t)w * h, c, g);
        free(img);
        img = g;
    } else if (job->flags & FO_INVERSE) {
        fo_invert(img, (size_t)w * h, c);
    }

    if (job->flags & FO_HAS_DIMS) {
        uint32_t width = job->req_w, height = job->req_h;
        if (width == 0 || height == 0) { free(img); return FO_EINVAL; }
        if (width != w || height != h) { /* handler.rs:231 / :336 */
            if (job->flags & FO_CROP) {
                /* DynamicImage::resize_to_fill */
                uint32_t w2, h2;
                fo_resize_dimensions(w, h, width, height, 1, &w2, &h2);
                uint8_t *mid = (uint8_t *)malloc((size_t)w2 * h2 * c);
                if (!mid) { free(img); return FO_ENOMEM; }
                rc = fo_resize(img, w, h, c, w2, h2, filter, mid);
                free(img);
                if (rc != FO_OK) { free(mid); return rc; }
                uint64_t ratio = (uint64_t)w2 * height;
                uint64_t nratio = (uint64_t)width * h2;
                uint8_t *out = (uint8_t *)malloc((size_t)width * height * c);
                if (!out) { free(mid); return FO_ENOMEM; }
                if (nratio > ratio) crop_copy(mid, w2, h2, c, 0, (h2 - height) / 2, width, height, out, &w, &h);
                else crop_copy(mid, w2, h2, c, (w2 - width) / 2, 0, width, height, out, &w, &h);
                free(mid);
                img = out;
            } else {
                /* DynamicImage::resize (aspect fit) */
                uint32_t w2, h2;
                fo_resize_dimensions(w, h, width, height, 0, &w2, &h2);
                uint8_t *out = (uint8_t *)malloc((size_t)w2 * h2 * c);
                if (!out) { free(img); return FO_ENOMEM; }
                rc = fo_resize(img, w, h, c, w2, h2, filter, out);
                free(img);
                if (rc != FO_OK) { free(out); return rc; }
                img = out;
                w = w2;
                h = h2;
            }
        }
        if (width > w || height > h) { /* handler.rs:238-248 / :343-353 */
            uint8_t *bg = (uint8_t *)malloc((size_t)width * height * 4);
            if (!bg) { free(img); return FO_ENOMEM; }
            for (size_t i = 0; i < (size_t)width * height; i++) {
                bg[4 * i] = job->fill[0];
                bg[4 * i + 1] = job->fill[1];
                bg[4 * i + 2] = job->fill[2];
                bg[4 * i + 3] = 255;
            }
            uint32_t dx = (width > w ? width - w : w - width) / 2;
            uint32_t dy = (height > h ? height - h : h - height) / 2;
            fo_overlay(bg, width, height, img, w, h, c, (int64_t)dx, (int64_t)dy);
            free(img);
            img = bg;
            w = width;
            h = height;
            c = 4;
        }
    }

    if (!gif && job->blur_sigma > 0.0f) { /* handler.rs:250-255; GIF path never blurs */
        uint8_t *b = (uint8_t *)malloc((size_t)w * h * c);
        if (!b) { free(img); return FO_ENOMEM; }
        rc = fo_blur(img, w, h, c, job->blur_sigma, b);
        free(img);
        if (rc != FO_OK) { free(b); return rc; }
        img = b;
    }

    if (gif && c != 4) { /* handler.rs:355 to_rgba8 */
        uint8_t *r = (uint8_t *)malloc((size_t)w * h * 4);
        if (!r) { free(img); return FO_ENOMEM; }
        fo_to_rgba8(img, (size_t)w * h, c, r);
        free(img);
        img = r;
        c = 4;
    }
    if (!gif && (job->flags & FO_TO_RGB8) && c != 3) { /* the JPEG branch: the encoder works on RGB8 */
        uint8_t *r = (uint8_t *)malloc((size_t)w * h * 3);
        if (!r) { free(img); return FO_ENOMEM; }
        fo_to_rgb8(img, (size_t)w * h, c, r);
        free(img);
        img = r;
        c = 3;
    }

    job->out_w = w;
    job->out_h = h;
    job->out_c = c;
    uint64_t need = (uint64_t)w * h * c;
    if (job->dst) {
        if (job->dst_cap < need) { free(img); return FO_ECAP; }
        memcpy(job->dst, img, need);
    }
    free(img);
    return FO_OK;
}

/* ---- CPU baseline driver: one image per thread over n_threads ----------- */

typedef struct {
    fo_job *jobs;
    uint32_t n;
    volatile uint32_t *next;
    int rc;
} fo_worker;

static void *worker_main(void *arg) {
    fo_worker *wk = (fo_worker *)arg;
    for (;;) {
        uint32_t i = __atomic_fetch_add(wk->next, 1, __ATOMIC_RELAXED);
        if (i >= wk->n) break;
        int rc = fo_process(&wk->jobs[i]);
        if (rc != FO_OK) wk->rc = rc;
    }
    return NULL;
}

int fo_process_batch(fo_job *jobs, uint32_t n, uint32_t n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    volatile uint32_t next = 0;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    fo_worker *wk = (fo_worker *)malloc(sizeof(fo_worker) * n_threads);
    if (!th || !wk) { free(th); free(wk); return FO_ENOMEM; }
    uint32_t started = 0;
    for (uint32_t t = 0; t < n_threads; t++) {
        wk[t].jobs = jobs;
        wk[t].n = n;
        wk[t].next = &next;
        wk[t].rc = FO_OK;
        if (pthread_create(&th[t], NULL, worker_main, &wk[t]) != 0) break;
        started++;
    }
    if (started == 0) { /* no threads: run inline */
        wk[0].jobs = jobs; wk[0].n = n; wk[0].next = &next; wk[0].rc = FO_OK;
        worker_main(&wk[0]);
        started = 0;
        int rc = wk[0].rc;
        free(th); free(wk);
        return rc;
    }
    int rc = FO_OK;
    for (uint32_t t = 0; t < started; t++) {
        pthread_join(th[t], NULL);
        if (wk[t].rc != FO_OK) rc = wk[t].rc;
    }
    free(th);
    free(wk);
    return rc;
}

/* ---- YCCK -> CMYK (decode side) -----------------------------------------
 * Restates /root/reference/src/handler.rs:420-439 -- IN TREE, so this function is pinned by the
 * reference's own source (unlike the image-crate restatement above): per 4-byte pixel
 *   r = clamp(y + 1.40200 cr - 179.456, 0, 255)              ((y + 1.402 cr) - 179.456, f32, no FMA)
 *   g = clamp(y - 0.34414 cb - 0.71414 cr + 135.45984, 0, 255)
 *   b = clamp(y + 1.77200 cb - 226.816, 0, 255)
 *   k = 255 - raw[3]
 * stored with Rust's `as u8` (truncation), in place.  A trailing partial pixel is left as it is
 * (the reference's loop would index out of bounds; zune-jpeg never yields one). */
void fo_ycck_to_cmyk(uint8_t *raw, size_t n_bytes) {
    for (size_t i = 0; i + 3 < n_bytes; i += 4) {
        const float y = (float)raw[i], cb = (float)raw[i + 1], cr = (float)raw[i + 2];
        float r = y + 1.40200f * cr;
        r = r - 179.456f;
        float g = y - 0.34414f * cb;
        g = g - 0.71414f * cr;
        g = g + 135.45984f;
        float b = y + 1.77200f * cb;
        b = b - 226.816f;
        r = r < 0.0f ? 0.0f : (r > 255.0f ? 255.0f : r);
        g = g < 0.0f ? 0.0f : (g > 255.0f ? 255.0f : g);
        b = b < 0.0f ? 0.0f : (b > 255.0f ? 255.0f : b);
        raw[i] = (uint8_t)r;
        raw[i + 1] = (uint8_t)g;
        raw[i + 2] = (uint8_t)b;
        raw[i + 3] = (uint8_t)(255u - raw[i + 3]);
    }
}

uint32_t fo_job_size(void) { return (uint32_t)sizeof(fo_job); }
