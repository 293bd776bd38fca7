// Fused separable resample: host-side work items and scatter tables.
//
// One CTA processes one (job, band of output rows) and sweeps the source window
// left to right in column chunks.  Per chunk: the vertical stage filters the
// chunk's columns into an f32 tile tmp[band rows][chunk elements] in shared
// memory; the horizontal stage then walks that tile along x with one thread per
// output row, scattering every intermediate value into the (at most 8) output
// pixels whose tap windows contain it.  Every source byte is read once from HBM,
// every output byte written once, and the f32 intermediate never leaves the SM.
//
// Both stages are driven by "scatter tables": for a source index s, the weights
// of s in each live output (slot = output index mod 8), a mask of live slots and
// a mask of slots whose output completes at s.  They are the transposed view of
// the per-output tap tables of plan.h (image-0.25.6 sample.rs).
#pragma once
#include <memory>
#include <vector>

#include "plan.h"

namespace fanlin {

constexpr uint32_t FUSED_SLOTS = 8;
constexpr uint32_t FUSED_WARPS = 8;
constexpr uint32_t FUSED_XEP = 260;      // floats per tmp row (256 elements + pad; == 4 mod 32, 16 B aligned)

// Weight scaling that lets u8 bits be used as (denormal) f32 operands without a
// conversion: float_from_bits(b) = b * 2^-149.  Vertical weights carry 2^100,
// horizontal weights 2^49; both are exact power-of-two scalings.
constexpr int FUSED_V_SCALE_LOG2 = 100;
constexpr int FUSED_H_SCALE_LOG2 = 49;

// Device descriptor of one fused work item.
struct FusedItem {
    const uint8_t *src;
    uint8_t *dst;
    uint32_t src_pitch;
    uint32_t c_mem, c, color_op;
    uint32_t px0;          // first source pixel of the window (multiple of 4)
    uint32_t n_px;         // window width in pixels
    uint32_t chunk_px, n_chunks;
    uint32_t band_r0, band_rows;  // rows of the produced rect this CTA makes
    uint32_t y0, n_y;      // source rows the band depends on
    uint32_t vw_off;       // float offset of the first warp's V weights (each warp: [rows streamed][8])
    uint32_t vinfo_off;    // u32 offset: FUSED_WARPS x {first row, n rows, first output, weights offset} + info words
    uint32_t vinfo_stride; // u32 words per warp block (header + rows)
    uint32_t hw_off;       // float offset of H weights [n_px][8]
    uint32_t hinfo_off;    // u32 offset of H info [n_px]
    uint32_t n_cols;       // produced columns
    uint32_t dst_pitch, c_out, canvas_w, canvas_h, dst_x, dst_y, epi, fill;
    uint32_t first_band, last_band;
};

// Shared tables of all fused items of a batch (deduplicated by geometry).
struct FusedTables {
    std::vector<float> w;
    std::vector<uint32_t> info;
};

// Scatter view of outputs [o0, o0+n) of `t` over sources [s0, s0+n_s): weights [n_s][8]
// (slot = (o - slot_base) % 8, times `scale`; w may be null) and live | flush << 8 masks.
// False when more than 8 outputs are live at one source index.
bool fused_scatter(const AxisTable &t, uint32_t o0, uint32_t n, uint32_t slot_base, uint32_t s0, uint32_t n_s, float scale,
                   float *w, uint32_t *info);

// True when stage `s` of a job can take the fused path (downscale-ish taps that fit
// 8 live outputs per source index, 4-byte aligned rows).
bool fused_eligible(const StagePlan &s, const fanlin_job &job);

// Appends the work items of stage `s` (one per band) and their tables.  `cache`
// deduplicates tables between jobs of identical geometry.
struct FusedCache;
FusedCache *fused_cache_new();
void fused_cache_free(FusedCache *);
// Builds (once per geometry) the tables of stage `s`; false when more than 8 outputs
// are live at one source index, in which case the stage takes the generic path.
bool fused_geometry_ok(const StagePlan &s, FusedCache *cache, FusedTables *tabs);
int fused_build(const StagePlan &s, const uint8_t *src, uint32_t src_pitch, uint8_t *dst, FusedCache *cache,
                FusedTables *tabs, std::vector<FusedItem> *items);
// Shared-memory geometry of the kernel (kernels_fused.cu): pixels per chunk, bytes for a band,
// and the most output rows one CTA can take.
uint32_t fused_chunk_px(uint32_t c);
size_t fused_smem_bytes(uint32_t c, uint32_t c_mem, uint32_t band_rows);
uint32_t fused_max_band(uint32_t c, uint32_t c_mem);
inline uint32_t fused_variant(const StagePlan &s) { return s.c | s.c_mem << 3 | s.color_op << 6; }

}  // namespace fanlin
