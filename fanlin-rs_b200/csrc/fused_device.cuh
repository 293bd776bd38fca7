// Device helpers shared by the fused resample kernels (CUDA-core and tensor-core
// vertical stage): asynchronous copies and the per-pixel epilogue.
#pragma once
#include "device_common.cuh"

namespace fanlin {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

// cp.async (LDGSTS): global -> shared, completion tracked per commit group (FIFO).
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void *g, bool on) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 4;\n\t}\n" ::"r"(saddr),
        "l"(g), "r"(int(on))
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16_if(uint32_t saddr, const void *g, bool on) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}\n" ::"r"(saddr),
        "l"(g), "r"(int(on))
        : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Epilogue for one produced pixel (already rounded channel values) at canvas address q.
template <int C>
__device__ __forceinline__ void emit_at(uint8_t *q, uint32_t epi, uint32_t fill, const uint32_t v[4]) {
    if (epi == EPI_PLAIN) {
        if constexpr (C == 4) {
            store_rgba(q, v[0] | v[1] << 8 | v[2] << 16 | v[3] << 24);
        } else {
#pragma unroll
            for (int k = 0; k < C; k++) q[k] = uint8_t(v[k]);
        }
    } else {
        uint32_t px = to_rgba_packed(v, C);
        if (epi == EPI_BLEND_FILL) px = blend_rgba(fill, px);
        store_rgba(q, px);
    }
}

template <int C, typename ITEM>
__device__ __forceinline__ void emit_px(const ITEM &it, uint32_t cx, uint32_t cy, const uint32_t v[4]) {
    uint8_t *q = it.dst + size_t(cy) * it.dst_pitch + size_t(cx) * it.c_out;
    if (it.epi == EPI_PLAIN) {
        if constexpr (C == 4) {
            store_rgba(q, v[0] | v[1] << 8 | v[2] << 16 | v[3] << 24);
        } else {
#pragma unroll
            for (int k = 0; k < C; k++) q[k] = uint8_t(v[k]);
        }
    } else {
        uint32_t px = to_rgba_packed(v, C);
        if (it.epi == EPI_BLEND_FILL) px = blend_rgba(it.fill, px);
        store_rgba(q, px);
    }
}


// Letterbox bars of a band: canvas pixels outside the placed rect (and the rows above /
// below the image for the first / last band) get the fill colour.
template <typename ITEM>
__device__ __forceinline__ void fill_bars(const ITEM &it, uint32_t warp, uint32_t lane, uint32_t n_warps) {
    if (it.epi != EPI_BLEND_FILL) return;
    const uint32_t ya = it.first_band ? 0u : it.dst_y + it.band_r0;
    const uint32_t yb = it.last_band ? it.canvas_h : it.dst_y + it.band_r0 + it.band_rows;
    const uint32_t iy0 = it.dst_y + it.band_r0, iy1 = iy0 + it.band_rows;
    for (uint32_t y = ya + warp; y < yb; y += n_warps) {
        const bool inside_rows = y >= iy0 && y < iy1;
        for (uint32_t x = lane; x < it.canvas_w; x += 32)
            if (!inside_rows || x < it.dst_x || x >= it.dst_x + it.n_cols)
                store_rgba(it.dst + size_t(y) * it.dst_pitch + size_t(x) * 4, it.fill);
    }
}

}  // namespace fanlin
