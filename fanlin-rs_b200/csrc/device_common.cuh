// Device helpers shared by the kernels: colour op on load, overlay blend,
// rounding and the canvas epilogue.  Arithmetic follows image-0.25.6 color.rs /
// imageops op for op; *_rn intrinsics keep nvcc from
// contracting a*b+c where the CPU path rounds twice.
#pragma once
#include <cuda_runtime.h>

#include "common.h"

namespace fanlin {

// color.rs rgb_to_luma: (2126 R + 7152 G + 722 B) / 10000 in u32, floor.
__device__ __forceinline__ uint32_t luma_u8(uint32_t r, uint32_t g, uint32_t b) {
    return (2126u * r + 7152u * g + 722u * b) / 10000u;
}

// Loads pixel (x, y) of the stage input and applies the colour op; v[0..c) are the
// channel values the filter sees (c = d.c).
__device__ __forceinline__ void load_px(const StageDesc &d, uint32_t x, uint32_t y, uint32_t v[4]) {
    const uint8_t *p = d.src + size_t(y) * d.src_pitch + size_t(x) * d.c_mem;
    if (d.color_op == COLOR_GRAY) {  // c_mem is 3 or 4 here
        v[0] = luma_u8(p[0], p[1], p[2]);
        if (d.c_mem == 4) v[1] = p[3];
    } else if (d.color_op == COLOR_INVERT) {
        const uint32_t ncol = (d.c_mem == 2 || d.c_mem == 4) ? d.c_mem - 1 : d.c_mem;
        for (uint32_t k = 0; k < d.c_mem; k++) v[k] = k < ncol ? 255u - p[k] : p[k];
    } else {
        for (uint32_t k = 0; k < d.c_mem; k++) v[k] = p[k];
    }
}

// clamp(t, 0, 255) then f32::round (half away from zero), as FloatNearest.  For t >= 0,
// round(t) = floor(t + 0.5); adding with round-toward-zero keeps the sum from rounding up to
// the next integer (t = 0.5 - 2^-25), and the saturating conversion is the clamp (NaN -> 0).
__device__ __forceinline__ uint32_t round_u8(float t) {
    uint32_t r;
    asm("cvt.rzi.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(__fadd_rz(t, 0.5f)));
    return r;
}

// DynamicImage pixel viewed as Rgba<u8> (to_rgba): L->(l,l,l,255) La->(l,l,l,a) Rgb->(r,g,b,255).
__device__ __forceinline__ uint32_t to_rgba_packed(const uint32_t v[4], uint32_t c) {
    switch (c) {
    case 1: return v[0] | v[0] << 8 | v[0] << 16 | 0xff000000u;
    case 2: return v[0] | v[0] << 8 | v[0] << 16 | v[1] << 24;
    case 3: return v[0] | v[1] << 8 | v[2] << 16 | 0xff000000u;
    default: return v[0] | v[1] << 8 | v[2] << 16 | v[3] << 24;
    }
}

__device__ __forceinline__ uint32_t trunc_u8(float v) {
    // num-traits f32 -> u8 (truncate); out-of-range would panic upstream, saturate here
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    return uint32_t(v);
}

// color.rs `impl Blend for Rgba<u8>`: src-over of fg onto bg, both packed RGBA.
__device__ __forceinline__ uint32_t blend_rgba(uint32_t bg, uint32_t fg) {
    const uint32_t fa8 = fg >> 24;
    if (fa8 == 0) return bg;
    if (fa8 == 255) return fg;
    const float m = 255.0f;
    const float bg_a = __fdiv_rn(float(bg >> 24), m), fg_a = __fdiv_rn(float(fa8), m);
    const float a_final = __fsub_rn(__fadd_rn(bg_a, fg_a), __fmul_rn(bg_a, fg_a));
    if (a_final == 0.0f) return bg;
    const float one_m = __fsub_rn(1.0f, fg_a);
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float b = __fdiv_rn(float((bg >> (8 * k)) & 255u), m);
        const float f = __fdiv_rn(float((fg >> (8 * k)) & 255u), m);
        const float o_a = __fadd_rn(__fmul_rn(f, fg_a), __fmul_rn(__fmul_rn(b, bg_a), one_m));
        out |= trunc_u8(__fmul_rn(m, __fdiv_rn(o_a, a_final))) << (8 * k);
    }
    out |= trunc_u8(__fmul_rn(m, a_final)) << 24;
    return out;
}

__device__ __forceinline__ void store_rgba(uint8_t *q, uint32_t px) {
    if ((reinterpret_cast<size_t>(q) & 3) == 0) {
        *reinterpret_cast<uint32_t *>(q) = px;
    } else {  // caller-provided device buffers need not be 4-byte aligned
        q[0] = uint8_t(px); q[1] = uint8_t(px >> 8); q[2] = uint8_t(px >> 16); q[3] = uint8_t(px >> 24);
    }
}

// Writes one produced pixel (channel values v[0..d.c)) to canvas position (cx, cy).
__device__ __forceinline__ void store_px(const StageDesc &d, uint32_t cx, uint32_t cy, const uint32_t v[4]) {
    uint8_t *q = d.dst + size_t(cy) * d.dst_pitch + size_t(cx) * d.c_out;
    if (d.epi == EPI_PLAIN || (d.epi & EPI_GRAY)) {  // (gray canvas: the opaque one-channel pixel itself)
        for (uint32_t k = 0; k < d.c; k++) q[k] = uint8_t(v[k]);
    } else {
        uint32_t px = to_rgba_packed(v, d.c);
        if ((d.epi & EPI_MASK) == EPI_BLEND_FILL) px = blend_rgba(d.fill, px);
        if (d.epi & EPI_RGB8) { q[0] = uint8_t(px); q[1] = uint8_t(px >> 8); q[2] = uint8_t(px >> 16); }
        else store_rgba(q, px);
    }
}

// image-0.25.6 codecs/jpeg/encoder.rs rgb_to_ycbcr for u8 subpixels (max = 255): f32, the coefficients divided by max
// first, products summed left to right, truncating (saturating) casts -- every operation rounded separately.
__device__ __forceinline__ void rgb_to_ycbcr_u8(uint32_t r8, uint32_t g8, uint32_t b8, uint8_t *y, uint8_t *cb, uint8_t *cr) {
    const float r = float(r8), g = float(g8), b = float(b8), m = 255.0f;
    const float yy = __fadd_rn(__fadd_rn(__fmul_rn(__fdiv_rn(76.245f, m), r), __fmul_rn(__fdiv_rn(149.685f, m), g)), __fmul_rn(__fdiv_rn(29.07f, m), b));
    const float cbv = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(__fdiv_rn(-43.0185f, m), r), __fmul_rn(__fdiv_rn(84.4815f, m), g)), __fmul_rn(__fdiv_rn(127.5f, m), b)), 128.0f);
    const float crv = __fadd_rn(__fsub_rn(__fsub_rn(__fmul_rn(__fdiv_rn(127.5f, m), r), __fmul_rn(__fdiv_rn(106.7685f, m), g)), __fmul_rn(__fdiv_rn(20.7315f, m), b)), 128.0f);
    *y = uint8_t(trunc_u8(yy)); *cb = uint8_t(trunc_u8(cbv)); *cr = uint8_t(trunc_u8(crv));
}

// Letterbox bars: canvas pixels outside the placed rect get the fill colour.
__device__ __forceinline__ void store_fill(const StageDesc &d, uint32_t cx, uint32_t cy) {
    if (d.epi & EPI_RGB8) {
        uint8_t *q = d.dst + size_t(cy) * d.dst_pitch + size_t(cx) * 3;
        q[0] = uint8_t(d.fill); q[1] = uint8_t(d.fill >> 8); q[2] = uint8_t(d.fill >> 16);
        return;
    }
    if (d.epi & EPI_GRAY) { d.dst[size_t(cy) * d.dst_pitch + cx] = uint8_t(d.fill); return; }
    store_rgba(d.dst + size_t(cy) * d.dst_pitch + size_t(cx) * 4, d.fill);
}

}  // namespace fanlin
