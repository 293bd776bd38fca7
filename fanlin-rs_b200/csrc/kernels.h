// Launchers of the sm_100a kernels of the pixel-transform stage.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "common.h"

namespace fanlin {

// Stream plus optional event bracketing of every kernel (fanlin_batch_set_timing).
struct LaunchCtx {
    cudaStream_t st = nullptr;
    std::vector<cudaEvent_t> *events = nullptr;  // pairs (begin, end) per kernel, grown on demand
    std::vector<const char *> *names = nullptr;
    size_t used = 0;
    void begin(const char *name) {
        if (!events) return;
        while (events->size() < used + 2) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            events->push_back(e);
        }
        names->push_back(name);
        cudaEventRecord((*events)[used], st);
    }
    void end() {
        if (!events) return;
        cudaEventRecord((*events)[used + 1], st);
        used += 2;
    }
};

// Dynamic shared memory opt-in of a kernel: only ever RAISED, under a lock, per (device, kernel) -- the context is
// Send + Sync, and a per-launch cudaFuncSetAttribute to that launch's own size could lower the limit between
// another thread's set and its launch.  Returns false (and leaves the error for cudaGetLastError) on failure.
bool ensure_dynamic_smem(const void *kernel, size_t bytes);

struct LaunchGeom {
    uint32_t n_jobs;
    uint32_t max_n_rows, max_n_sx, max_n_cols;  // over the descriptors of this launch
    uint32_t max_canvas_w, max_canvas_h;
};

// Exact path (crate operation order, no FMA contraction): vertical pass to an f32
// intermediate in HBM, then horizontal pass + epilogue.  Returns kernels launched.
int launch_sep_exact(const StageDesc *d_descs, const TapEntry *d_tab, const float *d_w, const LaunchGeom &g,
                     LaunchCtx &lc);
// The same stages for 16-bit / f32 subpixels (kernels_deep.cu): separable stage in the crate's operation order, compose-only
// stage, orientation pass, to_rgb8 -- StageDesc::s_in / s_out name the subpixel types.
int launch_sep_deep(const StageDesc *d_descs, const TapEntry *d_tab, const float *d_w, const LaunchGeom &g, LaunchCtx &lc);
int launch_compose_deep(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc);
int launch_orient_deep(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc);
int launch_to_rgb8_deep(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc);
// Fused separable resample (kernels_fused.cu); items of one (c, c_mem, colour op) variant.
struct FusedItem;
int launch_fused(const FusedItem *d_items, uint32_t n_items, uint32_t variant, uint32_t max_band_rows,
                 const float *d_w, const uint32_t *d_info, LaunchCtx &lc);
// Fused resample with the vertical pass on the tensor cores (kernels_fused_tc.cu).
struct FusedTcItem;
int launch_fused_tc(const FusedTcItem *d_items, const void *d_tmaps, uint32_t n_items, uint32_t c, size_t smem, const uint8_t *d_b,
                    const float *d_w, const uint32_t *d_info, LaunchCtx &lc);
int launch_fused_tc3(const FusedTcItem *d_items, const void *d_tmaps, uint32_t n_items, uint32_t c, size_t smem, const uint8_t *d_b,
                     const uint32_t *d_info, LaunchCtx &lc);
// Host: encodes the TMA tensor map (128 bytes at `out`) through which that kernel fetches rows of
// one image: dims {row bytes, rows}, box {128 B, box_rows}, 128-byte swizzle.  Returns false on failure.
// width_bytes: valid bytes per row (0 = pitch); columns beyond it are filled with zeros by the copy.
bool encode_row_tile_map(void *out, const void *base, uint32_t pitch, uint32_t rows, uint32_t box_rows, uint32_t width_bytes = 0);
// Fast Gaussian blur (kernels_blur.cu): items share channel count, radius and padded tap count.
struct BlurItem;
int launch_blur(const BlurItem *d_items, uint32_t n_items, uint32_t max_w, uint32_t max_h, uint32_t c, uint32_t radius,
                uint32_t taps_pad, const float *d_w, bool skip_v, LaunchCtx &lc);
// Compose-only stages: colour op / crop copy / letterbox / to_rgba8, and Nearest resamples (a gather through the tap tables).
int launch_compose(const StageDesc *d_descs, const TapEntry *d_tab, const LaunchGeom &g, LaunchCtx &lc);
// Vertical blur pass on the tensor cores (kernels_fused_tc.cu); writes the f32 intermediate the horizontal blur kernel reads.
struct BlurVTcItem;
int launch_blur_v_tc(const BlurVTcItem *d_items, const void *d_tmaps, uint32_t n_items, size_t smem, const uint8_t *d_b,
                     const uint32_t *d_info, LaunchCtx &lc);
// Blur with both passes on the tensor cores, u8 in, u8 out, no intermediate in HBM (kernels_blur_tc.cu).
struct BlurTcItem;
int launch_blur_tc(const BlurTcItem *d_items, const void *d_tmaps, uint32_t n_items, size_t smem, const uint8_t *d_b, const uint32_t *d_info,
                   const float *d_w, LaunchCtx &lc);
// Colour op alone over the needed source rows (in front of the tensor-core resample).
int launch_color_pass(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc);
// DynamicImage::to_rgb8 of the final image (FANLIN_TO_RGB8), scratch -> dst.
int launch_to_rgb8(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc);
// YCCK -> CMYK of the decode side (reference src/handler.rs:420-439); src == dst allowed.
int launch_ycck_to_cmyk(const uint8_t *d_src, uint8_t *d_dst, size_t n_px, LaunchCtx &lc);
// EXIF orientation (+ colour op) of the stored image into scratch, in front of every other stage.
int launch_orient_pass(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc);

}  // namespace fanlin
