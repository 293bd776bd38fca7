// Launchers of the sm_100a kernels of the pixel-transform stage.
#pragma once
#include <cuda_runtime.h>

#include "common.h"

namespace fanlin {

struct LaunchGeom {
    uint32_t n_jobs;
    uint32_t max_n_rows, max_n_sx, max_n_cols;  // over the descriptors of this launch
    uint32_t max_canvas_w, max_canvas_h;
};

// Exact path (crate operation order, no FMA contraction): vertical pass to an f32
// intermediate in HBM, then horizontal pass + epilogue.  Returns kernels launched.
int launch_sep_exact(const StageDesc *d_descs, const TapEntry *d_tab, const float *d_w, const LaunchGeom &g,
                     cudaStream_t st);
// Compose-only stages: colour op / crop copy / letterbox / to_rgba8.
int launch_compose(const StageDesc *d_descs, const LaunchGeom &g, cudaStream_t st);

}  // namespace fanlin
