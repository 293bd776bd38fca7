// Exact path of the separable stages (resample, blur) and the compose stage.
//
// "Exact" = the operation order of image-0.25.6 imageops/sample.rs: vertical pass
// first into an unclamped f32 intermediate, then horizontal pass, every tap a
// separately rounded multiply and add (vertical_sample / horizontal_sample,
// SURVEY.md A.3).  Output is bit-identical to the CPU path; the f32 intermediate
// lives in HBM, so this path is the parity anchor, not the fast one.
#include "device_common.cuh"
#include "kernels.h"

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

namespace fanlin {

namespace {

constexpr int TX = 128;

// One thread per (produced row, source column) of the intermediate.
__global__ void __launch_bounds__(TX) vpass_exact_kernel(const StageDesc *__restrict__ descs,
                                                         const TapEntry *__restrict__ tab,
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
        uint32_t v[4] = {0, 0, 0, 0};
        load_px(d, d.sx0 + x, e.left + i, v);
        const float wi = w[i];
#pragma unroll
        for (int k = 0; k < 4; k++) acc[k] = __fadd_rn(acc[k], __fmul_rn(float(v[k]), wi));
    }
    float *o = d.tmp + size_t(r) * d.tmp_pitch + size_t(x) * d.c;
    for (uint32_t k = 0; k < d.c; k++) o[k] = acc[k];
}

// One thread per canvas pixel: horizontal pass over the intermediate + epilogue,
// or the fill colour outside the placed rect.
__global__ void __launch_bounds__(TX) hpass_exact_kernel(const StageDesc *__restrict__ descs,
                                                         const TapEntry *__restrict__ tab,
                                                         const float *__restrict__ tw) {
    const StageDesc d = descs[blockIdx.y];
    const uint32_t xtiles = (d.canvas_w + TX - 1) / TX;
    if (blockIdx.x >= d.canvas_h * xtiles) return;
    const uint32_t cy = blockIdx.x / xtiles;
    const uint32_t cx = (blockIdx.x % xtiles) * TX + threadIdx.x;
    if (cx >= d.canvas_w) return;
    const uint32_t lx = cx - d.dst_x, ly = cy - d.dst_y;  // wraps when outside
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
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = round_u8(acc[k]);
    store_px(d, cx, cy, v);
}

// One stored pixel of CM channels with the colour op applied, packed one channel per byte.
template <int CM>
__device__ __forceinline__ uint32_t fetch_packed(const uint8_t *p, uint32_t color_op) {
    uint32_t b[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < CM; k++) b[k] = p[k];
    if (color_op == COLOR_GRAY && CM >= 3) return luma_u8(b[0], b[1], b[2]) | (CM == 4 ? b[3] << 8 : 0u);
    if (color_op == COLOR_INVERT) {
        constexpr int NCOL = (CM == 2 || CM == 4) ? CM - 1 : CM;  // alpha stays
#pragma unroll
        for (int k = 0; k < NCOL; k++) b[k] = 255u - b[k];
    }
    return b[0] | b[1] << 8 | b[2] << 16 | b[3] << 24;
}

// fetch_packed<4> of a pixel that is already in a register (r | g << 8 | b << 16 | a << 24)
__device__ __forceinline__ uint32_t op_packed4(uint32_t px, uint32_t color_op) {
    if (color_op == COLOR_GRAY) return luma_u8(px & 255u, (px >> 8) & 255u, (px >> 16) & 255u) | (px >> 24) << 8;
    if (color_op == COLOR_INVERT) return px ^ 0x00ffffffu;  // alpha stays
    return px;
}

// Compose-only: colour op on load, crop copy, letterbox, to_rgba8 -- and the Nearest resample of
// GIF frames (handler.rs:338,340), which is a gather: with tables (v_tab != NO_TABLE) the source
// pixel of an output is (h_tab[x].left, v_tab[y].left), its single tap.  Four consecutive canvas
// pixels of a row per thread; RGBA output leaves as one 16-byte store when the row allows.
constexpr uint32_t CMP_ROWS = 4;  // canvas rows per block: one row per block is 54 k blocks of 120 busy threads for C4 (A/B on one box: 0.129 -> 0.095 ms per 200 frames with four rows and the 16-byte loads below)
__global__ void __launch_bounds__(TX) compose_kernel(const StageDesc *__restrict__ descs, const TapEntry *__restrict__ tab) {
    const StageDesc &d = descs[blockIdx.y];
    const uint32_t cw = d.canvas_w, ch = d.canvas_h;
    const uint32_t xtiles = (cw + 4 * TX - 1) / (4 * TX);
    if (blockIdx.x >= ((ch + CMP_ROWS - 1) / CMP_ROWS) * xtiles) return;
    const uint32_t cy0 = (blockIdx.x / xtiles) * CMP_ROWS;  // CMP_ROWS canvas rows per block
    const uint32_t cx0 = ((blockIdx.x % xtiles) * TX + threadIdx.x) * 4;
    if (cx0 >= cw) return;
    const uint32_t c_mem = d.c_mem, C = d.c, c_out = d.c_out, epi = d.epi & EPI_MASK, fill = d.fill, color_op = d.color_op;  // (EPI_RGB8: c_out = 3 bytes of the packed pixel leave)
    const uint32_t dst_x = d.dst_x, dst_y = d.dst_y, n_cols = d.n_cols, n_rows = d.n_rows;
    const bool gather = d.v_tab != NO_TABLE;
    for (uint32_t cy = cy0; cy < min(cy0 + CMP_ROWS, ch); cy++) {
    const bool row_in = cy >= dst_y && cy - dst_y < n_rows;
    const uint32_t sy = !row_in ? 0u : gather ? tab[d.v_tab + d.oy0 + (cy - dst_y)].left : d.oy0 + (cy - dst_y);
    const uint8_t *srow = d.src + size_t(sy) * d.src_pitch;
    const uint32_t ox0 = d.ox0, h_tab = d.h_tab;
    // orient >= 2 (orientation applied after the resample, runtime.cpp): the source is a small image AS STORED and pixel
    // (ox0 + lx, oy0 + ly) of its oriented view is wanted -- the index map of orient_pass_kernel below
    const uint32_t orient = gather ? 0u : d.orient, SW = d.src_w, SH = d.src_h, yo = d.oy0 + (cy - dst_y);
    uint8_t *q = d.dst + size_t(cy) * d.dst_pitch + size_t(cx0) * c_out;
    uint32_t out[4];
    // four consecutive RGBA source pixels on a 16-byte boundary (GIF frames: 200 x 480x270 of them is C4) come as ONE load
    // instead of sixteen byte loads
    uint4 quad = make_uint4(0, 0, 0, 0);
    bool have_quad = false;
    if (c_mem == 4 && !gather && orient < 2 && row_in && cx0 >= dst_x && cx0 - dst_x + 4 <= n_cols && cx0 + 4 <= cw) {
        const uint8_t *p4 = srow + size_t(ox0 + cx0 - dst_x) * 4;
        if ((reinterpret_cast<uintptr_t>(p4) & 15) == 0) { quad = __ldg(reinterpret_cast<const uint4 *>(p4)); have_quad = true; }
    }
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
        const uint32_t cx = cx0 + k, lx = cx - dst_x;
        out[k] = fill;  // letterbox bar (only EPI_BLEND_FILL stages have pixels outside the placed rect)
        if (cx < cw && row_in && cx >= dst_x && lx < n_cols) {
            const uint32_t sx = gather ? tab[h_tab + ox0 + lx].left : ox0 + lx;
            const uint8_t *p = srow + size_t(sx) * c_mem;
            if (orient >= 2) {
                const uint32_t xo = ox0 + lx;
                uint32_t ux, uy;
                if (orient >= 5) {
                    ux = (orient == 5 || orient == 6) ? yo : SW - 1 - yo;
                    uy = (orient == 5 || orient == 8) ? xo : SH - 1 - xo;
                } else {
                    ux = (orient == 2 || orient == 3) ? SW - 1 - xo : xo;
                    uy = (orient == 3 || orient == 4) ? SH - 1 - yo : yo;
                }
                p = d.src + size_t(uy) * d.src_pitch + size_t(ux) * c_mem;
            }
            const uint32_t v = have_quad ? op_packed4(k == 0 ? quad.x : k == 1 ? quad.y : k == 2 ? quad.z : quad.w, color_op)
                             : c_mem == 4 ? fetch_packed<4>(p, color_op) : c_mem == 3 ? fetch_packed<3>(p, color_op)
                             : c_mem == 1 ? fetch_packed<1>(p, color_op) : fetch_packed<2>(p, color_op);
            if (epi == EPI_PLAIN) {
                out[k] = v;
            } else {  // the pixel viewed as Rgba<u8> (to_rgba): L -> (l,l,l,255), La -> (l,l,l,a), Rgb -> (r,g,b,255)
                const uint32_t l = v & 255u;
                uint32_t px = C == 1 ? (l * 0x010101u | 0xff000000u) : C == 2 ? (l * 0x010101u | ((v >> 8) & 255u) << 24)
                            : C == 3 ? (v | 0xff000000u) : v;
                if (epi == EPI_BLEND_FILL) px = blend_rgba(fill, px);
                out[k] = px;
            }
        }
    }
    const uint32_t n = min(4u, cw - cx0);
    if (c_out == 4 && n == 4 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
        *reinterpret_cast<uint4 *>(q) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
        for (uint32_t k = 0; k < n; k++)
            for (uint32_t b = 0; b < c_out; b++) q[k * c_out + b] = uint8_t(out[k] >> (8 * b));
    }
    }
}

// Colour op alone (grayscale / inverse), source rows [oy0, oy0 + n_rows) -> dst rows [0, n_rows):
// the pass in front of the tensor-core resample.  Four pixels of each of CP_ROWS rows per thread
// through 32-bit words when the rows are word-aligned, all loads issued before the first store
// (HBM-bound: c_mem + c bytes per pixel; one row per thread was bound by the load latency).
constexpr uint32_t CP_ROWS = 8;
__global__ void __launch_bounds__(256) color_pass_kernel(const StageDesc *__restrict__ descs) {
    const StageDesc &d = descs[blockIdx.z];
    const uint32_t row0 = blockIdx.y * CP_ROWS, x = (blockIdx.x * 256 + threadIdx.x) * 4;
    const uint32_t n_rows = d.n_rows, w = d.src_w, c_mem = d.c_mem, c = d.c, op = d.color_op;
    if (row0 >= n_rows || x >= w) return;
    const uint8_t *sp0 = d.src + size_t(d.oy0 + row0) * d.src_pitch + size_t(x) * c_mem;
    uint8_t *dp0 = d.dst + size_t(row0) * d.dst_pitch + size_t(x) * c;
    const uint32_t sp = d.src_pitch, dpp = d.dst_pitch;
    const bool words = x + 4 <= w && ((reinterpret_cast<uintptr_t>(sp0) | reinterpret_cast<uintptr_t>(dp0) | sp | dpp) & 3) == 0;
    if (!words) {  // row tail or unaligned rows
        const StageDesc dd = d;
        for (uint32_t r = 0; r < CP_ROWS && row0 + r < n_rows; r++)
            for (uint32_t q = 0; q < 4 && x + q < w; q++) {
                uint32_t v[4] = {0, 0, 0, 0};
                load_px(dd, x + q, dd.oy0 + row0 + r, v);
                for (uint32_t k = 0; k < c; k++) dp0[size_t(r) * dpp + q * c + k] = uint8_t(v[k]);
            }
        return;
    }
    uint32_t in[CP_ROWS][4];
#pragma unroll
    for (uint32_t r = 0; r < CP_ROWS; r++) {
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(sp0 + size_t(r) * sp);
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) in[r][k] = (row0 + r < n_rows && k < c_mem) ? __ldg(sw + k) : 0u;
    }
#pragma unroll
    for (uint32_t r = 0; r < CP_ROWS; r++) {
        if (row0 + r >= n_rows) break;
        uint32_t *dw = reinterpret_cast<uint32_t *>(dp0 + size_t(r) * dpp);
        const uint32_t *i4 = in[r];
        if (op == COLOR_INVERT) {  // c_mem == c words in, c words out; alpha (LA / RGBA) stays
            const uint32_t mask = c == 2 ? 0x00ff00ffu : c == 4 ? 0x00ffffffu : 0xffffffffu;
#pragma unroll
            for (uint32_t k = 0; k < 4; k++)
                if (k < c) dw[k] = i4[k] ^ mask;
        } else if (c_mem == 3) {  // RGB -> L: 12 bytes in, 4 out
            const uint32_t a = i4[0], b = i4[1], cc = i4[2];
            dw[0] = luma_u8(a & 255, (a >> 8) & 255, (a >> 16) & 255) | luma_u8(a >> 24, b & 255, (b >> 8) & 255) << 8 |
                    luma_u8((b >> 16) & 255, b >> 24, cc & 255) << 16 | luma_u8((cc >> 8) & 255, (cc >> 16) & 255, cc >> 24) << 24;
        } else {  // RGBA -> LA: 16 bytes in, 8 out
            uint32_t o[2] = {0, 0};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t p = i4[q];
                o[q / 2] |= (luma_u8(p & 255, (p >> 8) & 255, (p >> 16) & 255) | (p >> 24) << 8) << (16 * (q & 1));
            }
            dw[0] = o[0]; dw[1] = o[1];
        }
    }
}

// DynamicImage::to_rgb8 of the final image (handler.rs:274-278, the JPEG branch): src [h][w][c] -> dst
// [h][w][3], alpha dropped, luma replicated.  The output of the stage is small; one thread per pixel.
__global__ void __launch_bounds__(256) to_rgb8_kernel(const StageDesc *__restrict__ descs) {
    const StageDesc &d = descs[blockIdx.y];
    const uint32_t n = d.canvas_w * d.canvas_h, c = d.c_mem;
    const bool planes = d.epi != 0;  // FANLIN_TO_YCBCR: three planes of n bytes instead of interleaved RGB
    // four pixels per thread through whole words (one byte per access made this pass cost as much as the resample of a
    // C1 image: 0.35 us); the tail, unaligned buffers and planes of a size that is not a multiple of four go pixel by pixel
    const uint32_t c_out = d.c_out;  // 3, or 4 = (l, l, l, 255) of a one-channel image (the expansion behind a gray canvas, EPI_GRAY)
    const bool vec = (reinterpret_cast<uintptr_t>(d.src) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.dst) & (c_out == 4 ? 15 : 3)) == 0 && (!planes || (n & 3) == 0);
    const uint32_t n4 = vec ? n / 4 : 0;
    for (uint32_t t = blockIdx.x * 256 + threadIdx.x; t < n4; t += gridDim.x * 256) {
        uint32_t px[4];  // r | g << 8 | b << 16 of pixels 4 t .. 4 t + 3
        if (c == 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(d.src) + t);
            px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
        } else if (c == 3) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(d.src) + 3 * size_t(t);
            const uint32_t a0 = __ldg(w), a1 = __ldg(w + 1), a2 = __ldg(w + 2);
            px[0] = a0; px[1] = a0 >> 24 | a1 << 8; px[2] = a1 >> 16 | a2 << 16; px[3] = a2 >> 8;
        } else if (c == 2) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(d.src) + t);
            px[0] = (v.x & 0xffu) * 0x010101u; px[1] = ((v.x >> 16) & 0xffu) * 0x010101u;
            px[2] = (v.y & 0xffu) * 0x010101u; px[3] = ((v.y >> 16) & 0xffu) * 0x010101u;
        } else {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(d.src) + t);
#pragma unroll
            for (int k = 0; k < 4; k++) px[k] = ((v >> (8 * k)) & 0xffu) * 0x010101u;
        }
        if (c_out == 4) {  // behind a gray canvas: (l, l, l, 255)
            uint4 o4 = make_uint4(px[0] | 0xff000000u, px[1] | 0xff000000u, px[2] | 0xff000000u, px[3] | 0xff000000u);
            reinterpret_cast<uint4 *>(d.dst)[t] = o4;
        } else if (planes) {
            uint32_t wy = 0, wb = 0, wr = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint8_t y, cb, cr;
                rgb_to_ycbcr_u8(px[k] & 0xffu, (px[k] >> 8) & 0xffu, (px[k] >> 16) & 0xffu, &y, &cb, &cr);
                wy |= uint32_t(y) << (8 * k); wb |= uint32_t(cb) << (8 * k); wr |= uint32_t(cr) << (8 * k);
            }
            uint32_t *o = reinterpret_cast<uint32_t *>(d.dst);
            o[t] = wy; o[n / 4 + t] = wb; o[n / 2 + t] = wr;
        } else {
            uint32_t *o = reinterpret_cast<uint32_t *>(d.dst) + 3 * size_t(t);
            o[0] = (px[0] & 0xffffffu) | px[1] << 24;
            o[1] = ((px[1] >> 8) & 0xffffu) | px[2] << 16;
            o[2] = ((px[2] >> 16) & 0xffu) | px[3] << 8;
        }
    }
    for (uint32_t i = 4 * n4 + blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const uint8_t *p = d.src + size_t(i) * c;
        uint8_t *o = d.dst + size_t(i) * 3;
        const uint8_t r = p[0], g = c <= 2 ? r : p[1], b = c <= 2 ? r : p[2];
        if (c_out == 4) { uint8_t *o4 = d.dst + size_t(i) * 4; o4[0] = r; o4[1] = g; o4[2] = b; o4[3] = 255; continue; }
        if (planes) { rgb_to_ycbcr_u8(r, g, b, d.dst + i, d.dst + n + i, d.dst + 2 * size_t(n) + i); continue; }
        o[0] = r; o[1] = g; o[2] = b;
    }
}

// EXIF orientation (+ colour op) of the stored image: oriented rows [oy0, oy0 + n_rows) -> dst rows
// [0, n_rows).  Orientation::from_exif / apply_orientation of the image crate (handler.rs:221-223):
//   2 flip horizontal, 3 rotate 180, 4 flip vertical, 5 transpose (rotate90 + flip_horizontal),
//   6 rotate 90 clockwise, 7 transverse (rotate270 + flip_horizontal), 8 rotate 270.
// A block moves a 32 x 32 pixel tile; the transposing cases go through shared memory so that both
// the reads (along stored rows) and the writes (along oriented rows) are coalesced.
__global__ void __launch_bounds__(256) orient_pass_kernel(const StageDesc *__restrict__ descs) {
    constexpr uint32_t TK = 64;  // tile extent across the stored rows: 8 pixels per thread in flight
    __shared__ uint32_t tile[TK][33];  // [k][tx]: k counts stored rows (oriented rows, or oriented columns when the axes swap)
    const StageDesc &d = descs[blockIdx.z];
    const uint32_t ow = d.canvas_w;  // oriented width; the oriented height is the other stored dimension
    const uint32_t orient = d.orient, c_mem = d.c_mem, color_op = d.color_op, src_pitch = d.src_pitch, oy0 = d.oy0, n_rows = d.n_rows;
    const uint8_t *src = d.src;
    const bool swap = d.orient >= 5;
    const uint32_t bx = blockIdx.x * (swap ? TK : 32), by = blockIdx.y * (swap ? 32 : TK);
    if (bx >= ow || by >= d.n_rows) return;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t W = d.src_w, H = d.src_h;
    // fetch: tx runs along the stored rows (oriented y when the axes swap), four pixels per thread, all loads in flight
    uint32_t px[TK / 8];
#pragma unroll
    for (uint32_t q = 0; q < TK / 8; q++) {
        const uint32_t k = ty + 8 * q;
        const uint32_t xo = swap ? bx + k : bx + tx, row = swap ? by + tx : by + k, yo = oy0 + row;
        px[q] = 0;
        if (xo < ow && row < n_rows) {
            uint32_t sx, sy;
            if (swap) {
                sx = (orient == 5 || orient == 6) ? yo : W - 1 - yo;
                sy = (orient == 5 || orient == 8) ? xo : H - 1 - xo;
            } else {
                sx = (orient == 2 || orient == 3) ? W - 1 - xo : xo;
                sy = (orient == 3 || orient == 4) ? H - 1 - yo : yo;
            }
            const uint8_t *p = src + size_t(sy) * src_pitch + size_t(sx) * c_mem;
            px[q] = c_mem == 3 ? fetch_packed<3>(p, color_op) : c_mem == 4 ? fetch_packed<4>(p, color_op)
                  : c_mem == 1 ? fetch_packed<1>(p, color_op) : fetch_packed<2>(p, color_op);
        }
    }
#pragma unroll
    for (uint32_t q = 0; q < TK / 8; q++) tile[ty + 8 * q][tx] = px[q];
    __syncthreads();
    // store: oriented row segments of the tile as whole words, 8 threads per row
    const uint32_t C = d.c, sub = threadIdx.x & 7;
    const uint32_t tile_w = swap ? TK : 32, tile_h = swap ? 32 : TK;
    const uint32_t n_px = min(tile_w, ow - bx), n_bytes = n_px * C;
    for (uint32_t row = threadIdx.x >> 3; row < tile_h; row += 32) {
        if (by + row >= n_rows) break;
        uint8_t *q0 = d.dst + size_t(by + row) * d.dst_pitch + size_t(bx) * C;
        auto at = [&](uint32_t p) { return swap ? tile[p][row] : tile[row][p]; };  // pixel p of this oriented row
        if ((reinterpret_cast<uintptr_t>(q0) & 3) == 0) {
            for (uint32_t w = sub; 4 * w < n_bytes; w += 8) {
                uint32_t out = 0;
                uint32_t p = C == 4 ? w : C == 2 ? 2 * w : C == 1 ? 4 * w : (4 * w) / 3, ch = 4 * w - p * C;  // C is 1..4: constant divisors
#pragma unroll
                for (uint32_t bb = 0; bb < 4; bb++) {
                    out |= ((at(min(p, tile_w - 1)) >> (8 * ch)) & 0xffu) << (8 * bb);
                    if (++ch == C) { ch = 0; p++; }
                }
                if (4 * w + 4 <= n_bytes) *reinterpret_cast<uint32_t *>(q0 + 4 * w) = out;
                else for (uint32_t bb = 0; 4 * w + bb < n_bytes; bb++) q0[4 * w + bb] = uint8_t(out >> (8 * bb));
            }
        } else {
            for (uint32_t p = sub; p < n_px; p += 8)
                for (uint32_t ch = 0; ch < C; ch++) q0[p * C + ch] = uint8_t(at(p) >> (8 * ch));
        }
    }
}

}  // namespace

int launch_orient_pass(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc) {
    if (g.n_jobs == 0 || !g.max_canvas_w || !g.max_canvas_h) return 0;
    lc.begin("orient_pass_kernel");
    // tiles are 32 x 64 or 64 x 32 (axes swapped); the grid covers the larger count either way
    orient_pass_kernel<<<dim3((g.max_canvas_w + 31) / 32, (g.max_canvas_h + 31) / 32, g.n_jobs), 256, 0, lc.st>>>(d_descs);
    lc.end();
    return 1;
}

// YCCK -> CMYK of convert_jpeg_color_if_needed (reference src/handler.rs:420-439): per 4-byte
// pixel r/g/b = clamp(f32 expression of y, cb, cr) truncated, k = 255 - k.  The reference's f32 evaluation order
// with separately rounded multiplies and adds (__fmul_rn / __fadd_rn: no FMA contraction) makes this bit-exact.
// Four pixels (16 bytes) per thread and step, grid-stride; src == dst is allowed (the reference works in place).
__device__ __forceinline__ uint32_t ycck_px(uint32_t p) {
    const float y = float(p & 0xffu), cb = float((p >> 8) & 0xffu), cr = float((p >> 16) & 0xffu);
    float r = __fadd_rn(__fadd_rn(y, __fmul_rn(1.40200f, cr)), -179.456f);
    float g = __fadd_rn(__fadd_rn(__fadd_rn(y, -__fmul_rn(0.34414f, cb)), -__fmul_rn(0.71414f, cr)), 135.45984f);
    float b = __fadd_rn(__fadd_rn(y, __fmul_rn(1.77200f, cb)), -226.816f);
    r = fminf(fmaxf(r, 0.0f), 255.0f);
    g = fminf(fmaxf(g, 0.0f), 255.0f);
    b = fminf(fmaxf(b, 0.0f), 255.0f);
    return uint32_t(r) | uint32_t(g) << 8 | uint32_t(b) << 16 | (255u - (p >> 24)) << 24;  // `as u8`: truncation
}

__global__ void __launch_bounds__(256) ycck_to_cmyk_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t n_px) {
    const size_t n4 = n_px / 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    const size_t stride = size_t(gridDim.x) * 256;
    if (vec) {
        for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < n4; i += stride) {
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(src) + i);
            v.x = ycck_px(v.x); v.y = ycck_px(v.y); v.z = ycck_px(v.z); v.w = ycck_px(v.w);
            reinterpret_cast<uint4 *>(dst)[i] = v;
        }
    }
    // unaligned buffers, and the last n_px % 4 pixels: one pixel per thread, byte accesses
    const size_t first = vec ? n4 * 4 : 0;
    for (size_t i = first + size_t(blockIdx.x) * 256 + threadIdx.x; i < n_px; i += stride) {
        const uint8_t *p = src + 4 * i;
        const uint32_t o = ycck_px(uint32_t(p[0]) | uint32_t(p[1]) << 8 | uint32_t(p[2]) << 16 | uint32_t(p[3]) << 24);
        uint8_t *q = dst + 4 * i;
        q[0] = uint8_t(o); q[1] = uint8_t(o >> 8); q[2] = uint8_t(o >> 16); q[3] = uint8_t(o >> 24);
    }
}

int launch_ycck_to_cmyk(const uint8_t *d_src, uint8_t *d_dst, size_t n_px, LaunchCtx &lc) {
    if (n_px == 0) return 0;
    // 148 SMs x 8 resident blocks of 256 threads; fewer for small buffers
    const size_t want = (n_px / 4 + 255) / 256 + 1;
    const uint32_t blocks = uint32_t(std::min<size_t>(148 * 8, want));
    lc.begin("ycck_to_cmyk_kernel");
    ycck_to_cmyk_kernel<<<blocks, 256, 0, lc.st>>>(d_src, d_dst, n_px);
    lc.end();
    return 1;
}

int launch_to_rgb8(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc) {
    if (g.n_jobs == 0 || !g.max_canvas_w || !g.max_canvas_h) return 0;
    // four pixels per thread and about four steps per thread: a 300x200 output is 15 blocks (one pixel per thread and step
    // made the pass a quarter of a million tiny blocks per 1024 images)
    const uint32_t blocks = std::max<uint32_t>(1, std::min<uint32_t>(256, (g.max_canvas_w * g.max_canvas_h + 4095) / 4096));
    lc.begin("to_rgb8_kernel");
    to_rgb8_kernel<<<dim3(blocks, g.n_jobs), 256, 0, lc.st>>>(d_descs);
    lc.end();
    return 1;
}

int launch_color_pass(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc) {
    if (g.n_jobs == 0 || !g.max_canvas_w || !g.max_canvas_h) return 0;
    lc.begin("color_pass_kernel");
    color_pass_kernel<<<dim3((g.max_canvas_w + 1023) / 1024, (g.max_canvas_h + CP_ROWS - 1) / CP_ROWS, g.n_jobs), 256, 0, lc.st>>>(d_descs);
    lc.end();
    return 1;
}

int launch_sep_exact(const StageDesc *d_descs, const TapEntry *d_tab, const float *d_w, const LaunchGeom &g,
                     LaunchCtx &lc) {
    if (g.n_jobs == 0) return 0;
    const uint32_t vx = g.max_n_rows * ((g.max_n_sx + TX - 1) / TX);
    const uint32_t hx = g.max_canvas_h * ((g.max_canvas_w + TX - 1) / TX);
    int n = 0;
    if (vx) {
        lc.begin("vpass_exact_kernel");
        vpass_exact_kernel<<<dim3(vx, g.n_jobs), TX, 0, lc.st>>>(d_descs, d_tab, d_w);
        lc.end();
        n++;
    }
    if (hx) {
        lc.begin("hpass_exact_kernel");
        hpass_exact_kernel<<<dim3(hx, g.n_jobs), TX, 0, lc.st>>>(d_descs, d_tab, d_w);
        lc.end();
        n++;
    }
    return n;
}

int launch_compose(const StageDesc *d_descs, const TapEntry *d_tab, const LaunchGeom &g, LaunchCtx &lc) {
    if (g.n_jobs == 0) return 0;
    const uint32_t hx = ((g.max_canvas_h + CMP_ROWS - 1) / CMP_ROWS) * ((g.max_canvas_w + 4 * TX - 1) / (4 * TX));  // four pixels per thread, CMP_ROWS rows per block
    if (!hx) return 0;
    lc.begin("compose_kernel");
    compose_kernel<<<dim3(hx, g.n_jobs), TX, 0, lc.st>>>(d_descs, d_tab);
    lc.end();
    return 1;
}


bool ensure_dynamic_smem(const void *kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> granted;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lk(mu);
    size_t &have = granted[std::make_pair(dev, kernel)];
    if (bytes <= have) return true;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)) != cudaSuccess) return false;
    have = bytes;
    return true;
}

}  // namespace fanlin
