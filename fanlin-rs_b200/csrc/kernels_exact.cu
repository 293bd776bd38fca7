// Exact path of the separable stages (resample, blur) and the compose stage.
//
// "Exact" = the operation order of image-0.25.6 imageops/sample.rs: vertical pass
// first into an unclamped f32 intermediate, then horizontal pass, every tap a
// separately rounded multiply and add (vertical_sample / horizontal_sample,
// SURVEY.md A.3).  Output is bit-identical to the CPU path; the f32 intermediate
// lives in HBM, so this path is the parity anchor, not the fast one.
#include "device_common.cuh"
#include "kernels.h"

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

// Compose-only: colour op on load, crop copy, letterbox, to_rgba8.
__global__ void __launch_bounds__(TX) compose_kernel(const StageDesc *__restrict__ descs) {
    const StageDesc d = descs[blockIdx.y];
    const uint32_t xtiles = (d.canvas_w + TX - 1) / TX;
    if (blockIdx.x >= d.canvas_h * xtiles) return;
    const uint32_t cy = blockIdx.x / xtiles;
    const uint32_t cx = (blockIdx.x % xtiles) * TX + threadIdx.x;
    if (cx >= d.canvas_w) return;
    const uint32_t lx = cx - d.dst_x, ly = cy - d.dst_y;
    if (cx < d.dst_x || cy < d.dst_y || lx >= d.n_cols || ly >= d.n_rows) {
        store_fill(d, cx, cy);
        return;
    }
    uint32_t v[4] = {0, 0, 0, 0};
    load_px(d, d.ox0 + lx, d.oy0 + ly, v);
    store_px(d, cx, cy, v);
}

}  // namespace

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

int launch_compose(const StageDesc *d_descs, const LaunchGeom &g, LaunchCtx &lc) {
    if (g.n_jobs == 0) return 0;
    const uint32_t hx = g.max_canvas_h * ((g.max_canvas_w + TX - 1) / TX);
    if (!hx) return 0;
    lc.begin("compose_kernel");
    compose_kernel<<<dim3(hx, g.n_jobs), TX, 0, lc.st>>>(d_descs);
    lc.end();
    return 1;
}

}  // namespace fanlin
