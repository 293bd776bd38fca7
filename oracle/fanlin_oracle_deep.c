/*
 * fanlin_oracle_deep.c -- CPU restatement of fanlin-rs's pixel-transform stage for the DynamicImage variants with
 * 16-bit and f32 samples (ImageLuma16 / LumaA16 / Rgb16 / Rgba16 from 16-bit PNG / TIFF / PNM, ImageRgb32F / Rgba32F from
 * HDR / EXR), SURVEY.md 8f rank 4: in the reference they flow through the same generic code as the u8 variants
 * (src/handler.rs:219 DynamicImage::from_decoder, then :224-255).
 *
 * TEST INFRASTRUCTURE ONLY, PARITY UNPINNED -- the header of fanlin_oracle.c applies word for word: the arithmetic is the
 * crate `image` = 0.25.6, not vendored, no Rust toolchain here.  What this file restates beyond the u8 oracle:
 *
 *   sample.rs vertical_sample / horizontal_sample are generic over the subpixel S: taps accumulate `sample as f32 * w` in
 *     f32 (u16 -> f32 is exact), the horizontal pass stores NumCast::from(FloatNearest(clamp(t, S::DEFAULT_MIN_VALUE,
 *     S::DEFAULT_MAX_VALUE))): [0, 65535] + f32::round for u16; [0.0, 1.0] and NO rounding for f32 (so resize and blur
 *     clamp HDR values to 1.0);
 *   color.rs rgb_to_luma::<T> works in T::Larger: u32 for u16 ((2126 r + 7152 g + 722 b) / 10000 < 2^32), f64 for f32
 *     (((2126.0 r + 7152.0 g) + 722.0 b) / 10000.0, then `as f32`);
 *   dynimage.rs grayscale(): Luma16 -> Luma16, LumaA16 -> LumaA16, Rgb16 -> Luma16, Rgba16 -> LumaA16, but
 *     Rgb32F -> Rgb32F and Rgba32F -> Rgba32F (grayscale_with_type: the luma replicated, alpha kept) -- the pixel type of
 *     the f32 variants does NOT change;                                                     (version-sensitive point)
 *   color.rs Invert: T::DEFAULT_MAX_VALUE - c: 65535 - c, 1.0 - c; alpha kept;
 *   GenericImageView for DynamicImage (what imageops::overlay reads the top image through):
 *     p.get_pixel(x, y).to_rgba().into_color() -> Rgba<u8>, subpixels converted by FromPrimitive:
 *       u16 -> u8: (c + 128) / 257;   f32 -> u8: round(clamp(c, 0, 1) * 255) (f32);   missing alpha = MAX -> 255;
 *     so a letterboxed 16-bit or f32 image becomes an ImageRgba8 canvas (handler.rs:240-247) and a blur behind it is a u8 blur;
 *   dynimage.rs to_rgba8 / to_rgb8 (the WebP and JPEG branches, handler.rs:287 / the FANLIN_TO_RGB8 flag): the same
 *     per-subpixel conversions, luma replicated, alpha dropped for to_rgb8.
 *
 * What pins it: sample kind 0 (u8) of this file must reproduce fanlin_oracle.c bit for bit on every request
 * (tests/test_deep.py), an independent numpy restatement (oracle/np_restatement_deep.py) must agree bit for bit on
 * u16 and f32, and the analytic invariants (a u16 image whose samples are 257 x a u8 image, constant images, Nearest = gather).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FO_OK 0
#define FO_EINVAL 1
#define FO_ENOMEM 2
#define FO_ECAP 3

enum { FO_NEAREST = 0, FO_LANCZOS3 = 1, FO_GAUSSIAN_BLUR = 100 };
enum { FOD_U8 = 0, FOD_U16 = 1, FOD_F32 = 2 };
enum {
    FO_GRAYSCALE = 1u << 0,
    FO_INVERSE = 1u << 1,
    FO_HAS_DIMS = 1u << 2,
    FO_CROP = 1u << 3,
    FO_GIF_FRAME = 1u << 4, /* Nearest, no blur, to_rgba8 (handler.rs:329-355; GIF frames are RGBA8, kept for the u8 cross-check) */
    FO_TO_RGB8 = 1u << 5,   /* DynamicImage::to_rgb8 of the result */
    FO_TO_RGBA8 = 1u << 6,  /* DynamicImage::into_rgba8 of the result (the WebP branch, handler.rs:287) */
    FO_TO_YCBCR = 1u << 7   /* the planes Y, Cb, Cr the JPEG encoder derives from the result (codecs/jpeg/encoder.rs rgb_to_ycbcr on
                               to_rgb8() of it): [3][h][w] u8 */
};

/* image-0.25.6 codecs/jpeg/encoder.rs:
 *   fn rgb_to_ycbcr<P: Pixel>(pixel: P) -> (u8, u8, u8) {
 *       let [r, g, b] = pixel.to_rgb().0;  let max: f32 = P::Subpixel::DEFAULT_MAX_VALUE.to_f32().unwrap();  (r, g, b as f32)
 *       // Coefficients from JPEG File Interchange Format (Version 1.02), multiplied for 255 maximum.
 *       let y  =   76.245  / max * r + 149.685  / max * g +  29.07   / max * b;
 *       let cb = - 43.0185 / max * r -  84.4815 / max * g + 127.5    / max * b + 128.;
 *       let cr =  127.5    / max * r - 106.7685 / max * g -  20.7315 / max * b + 128.;
 *       (y as u8, cb as u8, cr as u8) }                       -- restated from memory, like the rest (version-sensitive) */
static uint8_t sat_u8(float v) { return v != v ? 0 : (v <= 0.0f ? 0 : (v >= 255.0f ? 255 : (uint8_t)v)); }
static void rgb_to_ycbcr(const uint8_t p[3], uint8_t *y, uint8_t *cb, uint8_t *cr) {
    const float max = 255.0f, r = (float)p[0], g = (float)p[1], b = (float)p[2];
    float yy = 76.245f / max * r + 149.685f / max * g;
    yy = yy + 29.07f / max * b;
    float c1 = -43.0185f / max * r - 84.4815f / max * g;
    c1 = c1 + 127.5f / max * b;
    c1 = c1 + 128.0f;
    float c2 = 127.5f / max * r - 106.7685f / max * g;
    c2 = c2 - 20.7315f / max * b;
    c2 = c2 + 128.0f;
    *y = sat_u8(yy); *cb = sat_u8(c1); *cr = sat_u8(c2);
}

/* exported by fanlin_oracle.c */
uint32_t fo_weight_table(int kind, float sigma, uint32_t n_in, uint32_t n_out, uint32_t *lefts, uint32_t *counts, float *weights,
                         uint32_t max_taps);
void fo_resize_dimensions(uint32_t width, uint32_t height, uint32_t nwidth, uint32_t nheight, int fill, uint32_t *ow, uint32_t *oh);
int fo_apply_orientation(const uint8_t *src, uint32_t w, uint32_t h, uint32_t c, uint32_t exif, uint8_t *dst, uint32_t *ow, uint32_t *oh);

typedef struct {
    const void *src;
    uint32_t src_w, src_h, src_c; /* channels 1..4 (f32: 3 or 4) */
    uint32_t sample;              /* FOD_U8 / FOD_U16 / FOD_F32; native byte order, tight rows */
    uint32_t flags;
    uint32_t req_w, req_h;
    uint8_t fill[3];
    uint8_t orientation;
    float blur_sigma;
    void *dst;
    uint64_t dst_cap; /* bytes */
    /* outputs */
    uint32_t out_w, out_h, out_c, out_sample;
} fod_job;

typedef struct {
    uint32_t w, h, c, sk;
    void *px;
} img_t;

static size_t bps(uint32_t sk) { return sk == FOD_U8 ? 1 : sk == FOD_U16 ? 2 : 4; }
static size_t img_bytes(const img_t *m) { return (size_t)m->w * m->h * m->c * bps(m->sk); }

static int img_alloc(img_t *m, uint32_t w, uint32_t h, uint32_t c, uint32_t sk) {
    m->w = w; m->h = h; m->c = c; m->sk = sk;
    m->px = malloc(img_bytes(m) ? img_bytes(m) : 1);
    return m->px ? FO_OK : FO_ENOMEM;
}

static float get_f(const img_t *m, size_t i) { /* `sample as f32` */
    switch (m->sk) {
    case FOD_U8: return (float)((const uint8_t *)m->px)[i];
    case FOD_U16: return (float)((const uint16_t *)m->px)[i];
    default: return ((const float *)m->px)[i];
    }
}

/* NumCast::from(FloatNearest(clamp(t, MIN, MAX))) of horizontal_sample */
static void put_sampled(img_t *m, size_t i, float t) {
    switch (m->sk) {
    case FOD_U8:
        t = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
        ((uint8_t *)m->px)[i] = (uint8_t)roundf(t);
        break;
    case FOD_U16:
        t = t < 0.0f ? 0.0f : (t > 65535.0f ? 65535.0f : t);
        ((uint16_t *)m->px)[i] = (uint16_t)roundf(t);
        break;
    default:
        t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
        ((float *)m->px)[i] = t;
        break;
    }
}

/* sample.rs vertical_sample then horizontal_sample with one filter kind (resize: Nearest / Lanczos3; blur: gaussian) */
static int two_pass(const img_t *in, uint32_t nw, uint32_t nh, int kind, float sigma, img_t *out) {
    const uint32_t w = in->w, h = in->h, c = in->c;
    uint32_t vmax = fo_weight_table(kind, sigma, h, nh, NULL, NULL, NULL, 0);
    uint32_t hmax = fo_weight_table(kind, sigma, w, nw, NULL, NULL, NULL, 0);
    if (!vmax || !hmax) return FO_EINVAL;
    uint32_t *vl = malloc(sizeof(uint32_t) * nh), *vc = malloc(sizeof(uint32_t) * nh);
    uint32_t *hl = malloc(sizeof(uint32_t) * nw), *hc = malloc(sizeof(uint32_t) * nw);
    float *vw = malloc(sizeof(float) * (size_t)nh * vmax), *hw = malloc(sizeof(float) * (size_t)nw * hmax);
    const size_t row = (size_t)w * c;
    float *tmp = malloc(sizeof(float) * row * nh);
    int rc = FO_ENOMEM;
    if (vl && vc && hl && hc && vw && hw && tmp && img_alloc(out, nw, nh, c, in->sk) == FO_OK) {
        rc = FO_OK;
        fo_weight_table(kind, sigma, h, nh, vl, vc, vw, vmax);
        fo_weight_table(kind, sigma, w, nw, hl, hc, hw, hmax);
        for (uint32_t oy = 0; oy < nh; oy++) {
            const float *ws = vw + (size_t)oy * vmax;
            for (size_t x = 0; x < row; x++) {
                float t = 0.0f;
                for (uint32_t i = 0; i < vc[oy]; i++) t += get_f(in, (size_t)(vl[oy] + i) * row + x) * ws[i];
                tmp[(size_t)oy * row + x] = t; /* unclamped f32 */
            }
        }
        for (uint32_t ox = 0; ox < nw; ox++) {
            const float *ws = hw + (size_t)ox * hmax;
            for (uint32_t y = 0; y < nh; y++)
                for (uint32_t ch = 0; ch < c; ch++) {
                    float t = 0.0f;
                    for (uint32_t i = 0; i < hc[ox]; i++) t += tmp[((size_t)y * w + hl[ox] + i) * c + ch] * ws[i];
                    put_sampled(out, ((size_t)y * nw + ox) * c + ch, t);
                }
        }
    }
    free(vl); free(vc); free(hl); free(hc); free(vw); free(hw); free(tmp);
    return rc;
}

static int resize(const img_t *in, uint32_t nw, uint32_t nh, int kind, img_t *out) {
    if (nw == in->w && nh == in->h) { /* imageops::resize copies when the size is unchanged */
        if (img_alloc(out, nw, nh, in->c, in->sk) != FO_OK) return FO_ENOMEM;
        memcpy(out->px, in->px, img_bytes(in));
        return FO_OK;
    }
    return two_pass(in, nw, nh, kind, 0.0f, out);
}

/* ---- colour ops ----------------------------------------------------------------------------- */

static int grayscale(const img_t *in, img_t *out) {
    const size_t n = (size_t)in->w * in->h;
    const uint32_t c = in->c;
    if (c <= 2) {
        if (img_alloc(out, in->w, in->h, c, in->sk) != FO_OK) return FO_ENOMEM;
        memcpy(out->px, in->px, img_bytes(in));
        return FO_OK;
    }
    if (in->sk == FOD_F32) { /* Rgb32F -> Rgb32F, Rgba32F -> Rgba32F: luma in f64, replicated */
        if (img_alloc(out, in->w, in->h, c, in->sk) != FO_OK) return FO_ENOMEM;
        const float *s = in->px;
        float *d = out->px;
        for (size_t i = 0; i < n; i++) {
            double l = 2126.0 * (double)s[i * c] + 7152.0 * (double)s[i * c + 1];
            l = l + 722.0 * (double)s[i * c + 2];
            const float lf = (float)(l / 10000.0);
            d[i * c] = d[i * c + 1] = d[i * c + 2] = lf;
            if (c == 4) d[i * c + 3] = s[i * c + 3];
        }
        return FO_OK;
    }
    const uint32_t oc = c - 2;
    if (img_alloc(out, in->w, in->h, oc, in->sk) != FO_OK) return FO_ENOMEM;
    for (size_t i = 0; i < n; i++) {
        uint32_t r, g, b, a = 0;
        if (in->sk == FOD_U8) {
            const uint8_t *s = (const uint8_t *)in->px + i * c;
            r = s[0]; g = s[1]; b = s[2]; if (c == 4) a = s[3];
        } else {
            const uint16_t *s = (const uint16_t *)in->px + i * c;
            r = s[0]; g = s[1]; b = s[2]; if (c == 4) a = s[3];
        }
        const uint32_t l = (2126u * r + 7152u * g + 722u * b) / 10000u;
        if (in->sk == FOD_U8) {
            uint8_t *d = (uint8_t *)out->px + i * oc;
            d[0] = (uint8_t)l; if (oc == 2) d[1] = (uint8_t)a;
        } else {
            uint16_t *d = (uint16_t *)out->px + i * oc;
            d[0] = (uint16_t)l; if (oc == 2) d[1] = (uint16_t)a;
        }
    }
    return FO_OK;
}

static void invert(img_t *m) {
    const size_t n = (size_t)m->w * m->h;
    const uint32_t c = m->c, ncol = (c == 2 || c == 4) ? c - 1 : c;
    for (size_t i = 0; i < n; i++)
        for (uint32_t k = 0; k < ncol; k++) {
            const size_t j = i * c + k;
            if (m->sk == FOD_U8) ((uint8_t *)m->px)[j] = (uint8_t)(255 - ((uint8_t *)m->px)[j]);
            else if (m->sk == FOD_U16) ((uint16_t *)m->px)[j] = (uint16_t)(65535 - ((uint16_t *)m->px)[j]);
            else ((float *)m->px)[j] = 1.0f - ((float *)m->px)[j];
        }
}

/* FromPrimitive<S> for u8 */
static uint8_t sub_to_u8(const img_t *m, size_t i) {
    if (m->sk == FOD_U8) return ((const uint8_t *)m->px)[i];
    if (m->sk == FOD_U16) return (uint8_t)(((uint32_t)((const uint16_t *)m->px)[i] + 128u) / 257u);
    float f = ((const float *)m->px)[i];
    f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
    return (uint8_t)roundf(f * 255.0f);
}

/* pixel i viewed as Rgba<u8>: to_rgba() in the subpixel type (missing alpha = MAX), then into_color() */
static void px_rgba8(const img_t *m, size_t i, uint8_t o[4]) {
    const size_t b = i * m->c;
    switch (m->c) {
    case 1: o[0] = o[1] = o[2] = sub_to_u8(m, b); o[3] = 255; break;
    case 2: o[0] = o[1] = o[2] = sub_to_u8(m, b); o[3] = sub_to_u8(m, b + 1); break;
    case 3: o[0] = sub_to_u8(m, b); o[1] = sub_to_u8(m, b + 1); o[2] = sub_to_u8(m, b + 2); o[3] = 255; break;
    default: o[0] = sub_to_u8(m, b); o[1] = sub_to_u8(m, b + 1); o[2] = sub_to_u8(m, b + 2); o[3] = sub_to_u8(m, b + 3); break;
    }
}

static uint8_t cast_u8_trunc(float v) {
    if (!(v > -1.0f)) return 0;
    if (v >= 256.0f) return 255;
    return (uint8_t)v;
}

/* color.rs: impl Blend for Rgba<u8> (as in fanlin_oracle.c) */
static void blend_rgba(uint8_t bg[4], const uint8_t fg[4]) {
    if (fg[3] == 0) return;
    if (fg[3] == 255) { memcpy(bg, fg, 4); return; }
    const float max_t = 255.0f;
    float bg_r = (float)bg[0] / max_t, bg_g = (float)bg[1] / max_t, bg_b = (float)bg[2] / max_t, bg_a = (float)bg[3] / max_t;
    float fg_r = (float)fg[0] / max_t, fg_g = (float)fg[1] / max_t, fg_b = (float)fg[2] / max_t, fg_a = (float)fg[3] / max_t;
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

static void img_free(img_t *m) { free(m->px); m->px = NULL; }

int fod_process(fod_job *job) {
    uint32_t w = job->src_w, h = job->src_h;
    const uint32_t c0 = job->src_c, sk = job->sample;
    if (!job->src || w == 0 || h == 0 || c0 < 1 || c0 > 4 || sk > FOD_F32) return FO_EINVAL;
    if (sk == FOD_F32 && c0 < 3) return FO_EINVAL; /* DynamicImage has no Luma32F variants */
    const int gif = (job->flags & FO_GIF_FRAME) != 0;
    if (gif && (c0 != 4 || sk != FOD_U8)) return FO_EINVAL;
    const int filter = gif ? FO_NEAREST : FO_LANCZOS3;
    img_t img;
    if (img_alloc(&img, w, h, c0, sk) != FO_OK) return FO_ENOMEM;
    int rc = FO_OK;
    if (!gif && job->orientation > 1) { /* handler.rs:221-223; a pixel is c * bps bytes */
        rc = fo_apply_orientation(job->src, w, h, (uint32_t)(c0 * bps(sk)), job->orientation, img.px, &w, &h);
        if (rc != FO_OK) { img_free(&img); return rc; }
        img.w = w; img.h = h;
    } else {
        memcpy(img.px, job->src, img_bytes(&img));
    }
    if (job->flags & FO_GRAYSCALE) { /* handler.rs:224-228 */
        img_t g;
        rc = grayscale(&img, &g);
        img_free(&img);
        if (rc != FO_OK) return rc;
        img = g;
    } else if (job->flags & FO_INVERSE) {
        invert(&img);
    }
    if (job->flags & FO_HAS_DIMS) {
        const uint32_t width = job->req_w, height = job->req_h;
        if (width == 0 || height == 0) { img_free(&img); return FO_EINVAL; }
        if (width != img.w || height != img.h) { /* handler.rs:231 */
            img_t mid;
            uint32_t w2, h2;
            if (job->flags & FO_CROP) { /* resize_to_fill */
                fo_resize_dimensions(img.w, img.h, width, height, 1, &w2, &h2);
                rc = resize(&img, w2, h2, filter, &mid);
                img_free(&img);
                if (rc != FO_OK) return rc;
                const uint64_t ratio = (uint64_t)w2 * height, nratio = (uint64_t)width * h2;
                uint32_t x = 0, y = 0;
                if (nratio > ratio) y = (h2 - height) / 2; else x = (w2 - width) / 2;
                uint32_t cw = width, ch = height;
                if (x > w2) x = w2;
                if (y > h2) y = h2;
                if (cw > w2 - x) cw = w2 - x;
                if (ch > h2 - y) ch = h2 - y;
                img_t out;
                if (img_alloc(&out, cw, ch, mid.c, mid.sk) != FO_OK) { img_free(&mid); return FO_ENOMEM; }
                const size_t pb = mid.c * bps(mid.sk);
                for (uint32_t j = 0; j < ch; j++)
                    memcpy((uint8_t *)out.px + (size_t)j * cw * pb, (const uint8_t *)mid.px + ((size_t)(y + j) * w2 + x) * pb, (size_t)cw * pb);
                img_free(&mid);
                img = out;
            } else {
                fo_resize_dimensions(img.w, img.h, width, height, 0, &w2, &h2);
                rc = resize(&img, w2, h2, filter, &mid);
                img_free(&img);
                if (rc != FO_OK) return rc;
                img = mid;
            }
        }
        if (width > img.w || height > img.h) { /* handler.rs:238-248: Rgba<u8> canvas, overlay through GenericImageView for DynamicImage */
            img_t bg;
            if (img_alloc(&bg, width, height, 4, FOD_U8) != FO_OK) { img_free(&img); return FO_ENOMEM; }
            uint8_t *b = bg.px;
            for (size_t i = 0; i < (size_t)width * height; i++) { b[4 * i] = job->fill[0]; b[4 * i + 1] = job->fill[1]; b[4 * i + 2] = job->fill[2]; b[4 * i + 3] = 255; }
            const uint32_t dx = (width > img.w ? width - img.w : img.w - width) / 2;
            const uint32_t dy = (height > img.h ? height - img.h : img.h - height) / 2;
            const uint32_t rw = img.w < width - dx ? img.w : width - dx, rh = img.h < height - dy ? img.h : height - dy;
            for (uint32_t j = 0; j < rh; j++)
                for (uint32_t i = 0; i < rw; i++) {
                    uint8_t p[4];
                    px_rgba8(&img, (size_t)j * img.w + i, p);
                    blend_rgba(b + ((size_t)(dy + j) * width + dx + i) * 4, p);
                }
            img_free(&img);
            img = bg;
        }
    }
    if (!gif && job->blur_sigma > 0.0f) { /* handler.rs:250-255: in the image's own subpixel type */
        img_t bl;
        rc = two_pass(&img, img.w, img.h, FO_GAUSSIAN_BLUR, job->blur_sigma, &bl);
        img_free(&img);
        if (rc != FO_OK) return rc;
        img = bl;
    }
    const int ycc = !gif && (job->flags & FO_TO_YCBCR);
    const int to4 = gif || (job->flags & FO_TO_RGBA8), to3 = !to4 && ((job->flags & FO_TO_RGB8) || ycc);
    if ((to4 && !(img.c == 4 && img.sk == FOD_U8)) || (to3 && !(img.c == 3 && img.sk == FOD_U8))) {
        img_t r;
        const uint32_t oc = to4 ? 4 : 3;
        if (img_alloc(&r, img.w, img.h, oc, FOD_U8) != FO_OK) { img_free(&img); return FO_ENOMEM; }
        for (size_t i = 0; i < (size_t)img.w * img.h; i++) {
            uint8_t p[4];
            px_rgba8(&img, i, p);
            memcpy((uint8_t *)r.px + i * oc, p, oc);
        }
        img_free(&img);
        img = r;
    }
    if (ycc) { /* img is RGB8 here */
        img_t pl;
        const size_t n = (size_t)img.w * img.h;
        if (img_alloc(&pl, img.w, img.h, 3, FOD_U8) != FO_OK) { img_free(&img); return FO_ENOMEM; }
        uint8_t *o = pl.px;
        for (size_t i = 0; i < n; i++) rgb_to_ycbcr((const uint8_t *)img.px + 3 * i, o + i, o + n + i, o + 2 * n + i);
        img_free(&img);
        img = pl;
    }
    job->out_w = img.w; job->out_h = img.h; job->out_c = img.c; job->out_sample = img.sk;
    const uint64_t need = img_bytes(&img);
    if (job->dst) {
        if (job->dst_cap < need) { img_free(&img); return FO_ECAP; }
        memcpy(job->dst, img.px, need);
    }
    img_free(&img);
    return FO_OK;
}

uint32_t fod_job_size(void) { return (uint32_t)sizeof(fod_job); }
