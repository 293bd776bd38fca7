// Fused separable resample, tensor-core vertical stage (sm_100a tcgen05 kind::i8).
//
// Same decomposition as fused.h (one CTA per image band, sweep of 128-byte column chunks,
// f32 tile in shared memory, scatter horizontal stage on the CUDA cores), but the
// vertical pass -- 6 FMA per source byte at Lanczos3, which no CUDA-core schedule
// gets under the HBM time on B200 -- runs as a banded integer
// contraction on the 5th-generation tensor cores:
//
//   D[128 columns x 96] (s32, TMEM) += A[128 columns x 32 rows] (u8) * B[32 rows x 96] (s8)
//
// A is the source tile exactly as it lies in the image (rows = K, bytes of a row =
// M, "MN-major"), fetched by TMA with the 128-byte swizzle; no byte is converted or
// shuffled by a thread.  B holds, for a group of 32 output rows, the
// filter weights as three signed base-128 digits of round(w * 2^s) (columns 0-31 hi,
// 32-63 mid, 64-95 lo); the epilogue recombines the three s32 sums exactly.  Integer
// arithmetic is exact; the only deviation from the f32 recipe is the 2^-s weight
// quantisation (s >= 20), below the rounding noise of the f32 accumulation it replaces.
//
// Where the outputs a 128-byte chunk touches fit a ring of accumulator columns (fused_tc.cpp
// build_hmma), the horizontal pass runs on the tensor cores too (fused_resample_tc2_kernel):
// the vertical results become f16 hi / lo operand tiles and are contracted with per-chunk f16
// weight tiles (kind::f16, f32 accumulators); the CUDA cores only drain, convert and store.
#pragma once
#include <vector>

#include "fused.h"

namespace fanlin {

constexpr uint32_t TC_M = 128;          // source bytes (elements) per chunk row = UMMA M
constexpr uint32_t TC_GROUP_ROWS = 32;  // most output rows per MMA group (N = 3 * 32); a band may use fewer (grp_rows)
constexpr uint32_t TC_N = 3 * TC_GROUP_ROWS;
constexpr uint32_t TC_KG_MAX = 256;     // source rows per group, multiple of 32
constexpr uint32_t TC_H_WARPS = 8;      // consumer warps; the horizontal stage gives each (row pair, channel) a thread

struct FusedTcItem {
    const uint8_t *src;
    uint8_t *dst;
    uint32_t src_pitch, src_h;
    uint32_t c;
    uint32_t b0, n_chunks, chunk_off, max_pairs;  // first source byte (16-byte aligned); 128-byte chunks; u32 offset of their records
    uint32_t band_r0, band_rows, r_pad;   // r_pad: floats per tile column (see r_pad_for)
    uint32_t grp_rows;                    // output rows per group (<= 32), chosen to minimise MMAs per row
    uint32_t grp_off, n_groups, kg_max;   // grp_off: u32 offset of {k0, kg, b_off, rows} x n_groups
    float scale;                          // 2^-s
    uint32_t hw_off, hinfo_off;           // horizontal pair table (unscaled weights), outputs finished per pair
    uint32_t cpre_off;                    // u32 offset of the per-chunk prefix of finished outputs (n_chunks + 1)
    uint32_t out_stride;                  // words per row of the shared-memory output staging buffer
    uint32_t n_a;                         // shared-memory slots for a group's source rows (2..4): as many as fit
    uint32_t n_cols;
    uint32_t dst_pitch, c_out, canvas_w, canvas_h, dst_x, dst_y, epi, fill;
    uint32_t first_band, last_band;
    // tensor-core horizontal stage (fused_resample_tc2_kernel): per-chunk records {byte offset of the chunk's f16
    // weight tiles, first output finished by the chunk (relative to the first produced column), outputs finished}
    uint32_t hmma, hrec_off, n_wh;
    // hmma == 2: the horizontal sums stay in TMEM across chunks (fused_resample_tc3_kernel): a ring of ring_cols accumulator
    // columns per row tile (output pixel o at column (o mod ring pixels) * c), n_vr vertical regions behind the rings,
    // stage_stride words per row of a consumer warp's output staging tile; per-chunk records of 8 words
    // {tile offset, first output finished, outputs finished, window column, window columns, ring slot of the first finished, 0, 0}
    uint32_t ring_cols, n_vr, stage_stride, wh_bytes;
    uint32_t inv_off;  // inverse on load: u32 offset of the band's per-row constants (f32 bits), 0xffffffff = no colour op
};

// Tensor-core horizontal stage: the chunk's 128 tile columns (K) are contracted with an f16 weight
// tile [N2][128] into N2 = ring positions x channels accumulator columns (ring of N2 / c outputs).
inline uint32_t fused_tc2_n(uint32_t c) { return c == 3 ? 48u : 64u; }
constexpr float TC2_WSCALE = 16.0f;  // horizontal weights are stored x16: their low halves stay f16 normals
size_t fused_tc2_smem_bytes(uint32_t c, uint32_t n_groups, uint32_t kg_max, uint32_t n_a, uint32_t n_wh, uint32_t band_rows, uint32_t out_stride);
size_t fused_tc3_smem_bytes(uint32_t n_groups, uint32_t kg_max, uint32_t n_a, uint32_t n_wh, uint32_t wh_bytes, uint32_t stage_stride);
size_t fused_tc_item_smem(const FusedTcItem &it);

// Vertical Gaussian pass of the blur on the tensor cores (kernels_fused_tc.cu blur_v_tc_kernel):
// the same banded integer contraction, output rows = input rows, written as the f32 intermediate
// [h][w * c] the horizontal blur kernel reads (imageops::blur, handler.rs:250-255).
struct BlurVTcItem {
    const uint8_t *src;
    float *dst;
    uint32_t src_pitch, src_h;
    uint32_t n_e, n_chunks;               // elements per row (w * c); 128-element column chunks
    uint32_t band_r0, band_rows;          // <= 24 groups of 32 rows
    uint32_t grp_off, n_groups, kg_max;   // u32 offset of {k0, kg, b_off, rows} x n_groups
    uint32_t n_a;                         // source slots
    float scale;                          // 2^-s
};

struct FusedTcTables {
    std::vector<uint8_t> b;  // weight digit tiles, core-matrix layout, 128-byte aligned per tile
};

bool fused_tc_eligible(const StagePlan &s, const fanlin_job &job);
struct FusedTcCache;
FusedTcCache *fused_tc_cache_new(bool allow_hmma = true);
void fused_tc_cache_free(FusedTcCache *);
bool fused_tc_geometry_ok(const StagePlan &s, FusedTcCache *cache, FusedTables *tabs, FusedTcTables *tctabs);
bool fused_tc_uses_hmma(const StagePlan &s, FusedTcCache *cache, FusedTables *tabs, FusedTcTables *tctabs);
bool fused_tc_uses_ring(const StagePlan &s, FusedTcCache *cache, FusedTables *tabs, FusedTcTables *tctabs);
int fused_tc_build(const StagePlan &s, const fanlin_job &job, const uint8_t *src, uint32_t src_pitch, uint8_t *dst,
                   FusedTcCache *cache, FusedTables *tabs, FusedTcTables *tctabs, std::vector<FusedTcItem> *items);

// Vertical blur: eligibility of a Gaussian stage whose input rows are `pitch` bytes apart, and its items (one per band).
bool blur_v_tc_eligible(const StagePlan &s, uint32_t pitch, const uint8_t *src);
int blur_v_tc_build(const StagePlan &s, const uint8_t *src, uint32_t src_pitch, float *dst, FusedTcCache *cache, FusedTables *tabs,
                    FusedTcTables *tctabs, std::vector<BlurVTcItem> *items);
size_t blur_v_tc_smem_bytes(uint32_t kg_max, uint32_t n_a);
uint32_t fused_tc_max_pairs(uint32_t c);
size_t fused_tc_smem_limit();  // kernels_fused_tc.cu: opt-in shared memory per block minus the kernel's static part
size_t fused_tc_smem_bytes(uint32_t c, uint32_t band_rows, uint32_t kg_max, uint32_t out_stride, uint32_t n_a);
uint32_t fused_tc_source_slots(uint32_t c, uint32_t band_rows, uint32_t kg_max, uint32_t out_stride);
uint32_t fused_tc_max_band(uint32_t c, uint32_t out_stride);

}  // namespace fanlin
