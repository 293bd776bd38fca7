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


// ---- both blur passes on the tensor cores, no f32 intermediate in HBM (kernels_blur_tc.cu) ----------------
// One CTA per band of <= 128 rows, swept left to right in chunks of 128 bytes per row:
//   vertical    banded integer contraction of fused_tc.h (u8 rows x s8 weight digits -> s32 in TMEM, exact), built from the
//               crate's own tap tables, so the renormalised border rows are reproduced tap for tap; ONE source box per
//               chunk (the band's rows and their halo) shared by the four 32-row groups;
//   horizontal  the vertical results become f16 hi / lo operand tiles T[128 rows][128 bytes] in shared memory; per 32
//               bytes of T a tcgen05.mma kind::f16 series against ONE resident Toeplitz weight tile W[n_win][32]
//               accumulates into a ring of 256 TMEM columns (column = output byte mod 256); a sub-step finishes the 32
//               oldest columns of its window, which the consumers drain, scale by the border factor of their column,
//               round, store and zero.
// imageops::blur as called at reference src/handler.rs:250-255.
struct BlurTcItem {
    const uint8_t *src;
    uint8_t *dst;
    uint32_t src_pitch, dst_pitch, src_h;
    uint32_t n_e;                 // bytes per row (w * c)
    uint32_t n_chunks;            // 128-byte chunks including the flush: ceil((n_e + r_pad) / 128)
    uint32_t band_r0, band_rows;  // <= 128 rows
    uint32_t grp_off, n_groups;   // u32 offset of {row offset of the group's window in the box, kg, b_off, rows} x n_groups
    uint32_t box_row0, box_rows;  // source rows [box_row0, + box_rows) x 128 bytes per chunk: one TMA box
    uint32_t kg_max;              // largest K extent of a group (sizes the vertical weight slots)
    float scale;                  // 2^-s of the vertical weight digits
    uint32_t hw_off;              // byte offset of the horizontal weight tile in the tile arena: hi [n_win][32] | lo [n_win][32] f16
    uint32_t n_win, r_pad, slack; // window columns (32 + 2 r_pad), radius in bytes padded to 16, sub-steps the MMAs may run ahead of the drain
    uint32_t corr_off;            // float offset of the per-byte-column factors (border renormalisation / 16), 128 * n_chunks entries
    uint32_t c_lo, c_hi;          // output bytes [c_lo, c_hi) have their whole window inside the row: their factor is 1 / 16
};

struct FusedTables;
struct FusedTcTables;
struct BlurTcCache;
BlurTcCache *blur_tc_cache_new();
void blur_tc_cache_free(BlurTcCache *);
// True when the blur stage `s` (already blur_eligible) whose input rows lie `pitch` bytes apart at `src` fits the kernel.
bool blur_tc_eligible(const StagePlan &s, uint32_t pitch, const uint8_t *src);
// Appends one item per band; tables go to tabs->w (factors), tabs->info (group records), tct->b (weight tiles).
int blur_tc_build(const StagePlan &s, const uint8_t *src, uint32_t src_pitch, uint8_t *dst, uint32_t dst_pitch, BlurTcCache *cache,
                  BlurTables *bt, FusedTables *tabs, FusedTcTables *tct, std::vector<BlurTcItem> *items);
size_t blur_tc_smem_bytes(uint32_t box_rows, uint32_t kg_max, uint32_t n_win);

}  // namespace fanlin
