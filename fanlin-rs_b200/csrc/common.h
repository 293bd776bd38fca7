// Shared host/device definitions of the pixel-transform stage.
#pragma once
#include <stdint.h>

#include "../../include/fanlin_device.h"

namespace fanlin {

enum ColorOp : uint32_t { COLOR_NONE = 0, COLOR_GRAY = 1, COLOR_INVERT = 2 };
enum Epilogue : uint32_t { EPI_PLAIN = 0, EPI_BLEND_FILL = 1, EPI_TO_RGBA = 2 };
// Flag on EPI_BLEND_FILL / EPI_TO_RGBA: the Rgba<u8> pixel leaves without its alpha byte (c_out = 3) -- DynamicImage::to_rgb8
// of the result folded into the last kernel (FANLIN_TO_RGB8, the JPEG branch handler.rs:274-278).  Understood by the
// both-passes tensor-core kernels, the compose kernel and the generic (exact / deep) horizontal passes; the planner in
// runtime.cpp sets it only where one of them writes the final image.
constexpr uint32_t EPI_RGB8 = 4u;
constexpr uint32_t EPI_MASK = 3u;
// Flag on EPI_BLEND_FILL for a one-channel image on a GRAY fill colour with a blur behind it: the canvas holds one byte per
// pixel (c_out = 1) -- the luma where the image lies (an opaque pixel blended onto the fill is the pixel), the fill's gray value
// in the bars; the blur runs on that plane and a last pass expands it to (l, l, l, 255).  All four channels of the Rgba<u8>
// canvas the reference blurs are this plane or the constant 255, so the result is the same and the blur moves a quarter of the bytes.
constexpr uint32_t EPI_GRAY = 8u;
enum FilterKind : uint32_t { KIND_NEAREST = 0, KIND_LANCZOS3 = 1, KIND_GAUSSIAN = 100 };

// Subpixel types (enum fanlin_sample) and their size in bytes.
enum Sample : uint32_t { SAMPLE_U8 = 0, SAMPLE_U16 = 1, SAMPLE_F32 = 2 };
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t sample_bytes(uint32_t s) { return s == SAMPLE_U8 ? 1u : s == SAMPLE_U16 ? 2u : 4u; }

// One entry of an axis table: output index o reads source [left, left+count) with
// weights tab_w[woff .. woff+count).
struct TapEntry {
    uint32_t left, count, woff;
};

// Device descriptor of one separable stage (resample or blur) of one job, or of a
// compose-only stage (v_tab == NO_TABLE).  The ragged batch is an array of these.
constexpr uint32_t NO_TABLE = 0xffffffffu;

struct StageDesc {
    const uint8_t *src;
    uint8_t *dst;
    float *tmp;           // exact path: [n_rows][tmp_pitch] f32
    uint32_t src_pitch, src_w, src_h;
    uint32_t c_mem;       // channels in memory at src
    uint32_t c;           // channels after the colour op (what the filter sees)
    uint32_t color_op;
    uint32_t v_tab, h_tab;  // offsets (in TapEntry units) into the table arena
    uint32_t oy0, n_rows;   // rows of the filtered image that are produced
    uint32_t ox0, n_cols;   // columns of the filtered image that are produced
    uint32_t sx0, n_sx;     // source columns the produced columns depend on
    uint32_t sy0, n_sy;     // source rows the produced rows depend on
    uint32_t tmp_pitch;     // floats per tmp row
    uint32_t dst_pitch, c_out, canvas_w, canvas_h;
    uint32_t dst_x, dst_y;  // where the produced rect lands on the canvas
    uint32_t epi;
    uint32_t fill;          // r | g<<8 | b<<16 | 255<<24
    uint32_t v_max_taps, h_max_taps;
    uint32_t orient;        // orientation pass only: EXIF orientation (2..8) of the stored image
    uint32_t s_in, s_out;   // subpixel types at src and dst (Sample); everything but the kernels of kernels_deep.cu sees u8 only
};

}  // namespace fanlin
