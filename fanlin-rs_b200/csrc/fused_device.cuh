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
// below the image for the first / last band) get the fill colour.  Only the bar pixels are
// visited, and every field of the item is read once: the item lives in shared memory and the
// stores go through generic pointers, so the compiler would re-read it around every store
// (measured: 58 k clk per 300 x 200 canvas for the pixel-by-pixel test of the whole canvas,
// 12 % of a C2 image's time on the SM).
template <typename ITEM>
__device__ __forceinline__ void fill_bars(const ITEM &it, uint32_t warp, uint32_t lane, uint32_t n_warps) {
    if ((it.epi & EPI_MASK) != EPI_BLEND_FILL) return;
    if (it.epi & EPI_GRAY) {  // one-byte canvas: the fill's gray value
        uint8_t *const dst = it.dst;
        const uint32_t pitch = it.dst_pitch, cw = it.canvas_w, v = it.fill & 0xffu;
        const uint32_t iy0 = it.dst_y + it.band_r0, iy1 = iy0 + it.band_rows;
        const uint32_t ya = it.first_band ? 0u : iy0, yb = it.last_band ? it.canvas_h : iy1;
        const uint32_t x0 = it.dst_x, x1 = it.dst_x + it.n_cols;
        for (uint32_t y = ya + warp; y < yb; y += n_warps) {
            uint8_t *row = dst + size_t(y) * pitch;
            const bool img_row = y >= iy0 && y < iy1;
            for (uint32_t x = lane; x < cw; x += 32)
                if (!(img_row && x >= x0 && x < x1)) row[x] = uint8_t(v);
        }
        return;
    }
    if (it.epi & EPI_RGB8) {  // RGB8 canvas: the fill colour without its alpha byte
        uint8_t *const dst = it.dst;
        const uint32_t pitch = it.dst_pitch, cw = it.canvas_w, fill = it.fill;
        const uint32_t iy0 = it.dst_y + it.band_r0, iy1 = iy0 + it.band_rows;
        const uint32_t ya = it.first_band ? 0u : iy0, yb = it.last_band ? it.canvas_h : iy1;
        const uint32_t x0 = it.dst_x, x1 = it.dst_x + it.n_cols;
        for (uint32_t y = ya + warp; y < yb; y += n_warps) {
            uint8_t *row = dst + size_t(y) * pitch;
            const bool img_row = y >= iy0 && y < iy1;
            for (uint32_t b = lane; b < cw * 3; b += 32) {
                const uint32_t x = b / 3;
                if (img_row && x >= x0 && x < x1) continue;
                row[b] = uint8_t(fill >> (8 * (b - 3 * x)));
            }
        }
        return;
    }
    uint8_t *const dst = it.dst;
    const uint32_t pitch = it.dst_pitch, cw = it.canvas_w, fill = it.fill;
    const uint32_t iy0 = it.dst_y + it.band_r0, iy1 = iy0 + it.band_rows;
    const uint32_t ya = it.first_band ? 0u : iy0, yb = it.last_band ? it.canvas_h : iy1;
    const uint32_t x0 = it.dst_x, x1 = it.dst_x + it.n_cols;
    const uint32_t n_top = iy0 - ya, n_full = n_top + (yb - iy1);  // whole rows above and below the image rows
    for (uint32_t k = warp; k < n_full; k += n_warps) {
        uint8_t *row = dst + size_t(k < n_top ? ya + k : iy1 + (k - n_top)) * pitch;
        for (uint32_t x = lane; x < cw; x += 32) store_rgba(row + size_t(x) * 4, fill);
    }
    const uint32_t n_side = x0 + (cw - x1);  // bar pixels left and right of the image in its rows
    if (n_side) {
        for (uint32_t y = iy0 + warp; y < iy1; y += n_warps) {
            uint8_t *row = dst + size_t(y) * pitch;
            for (uint32_t k = lane; k < n_side; k += 32) store_rgba(row + size_t(k < x0 ? k : x1 + (k - x0)) * 4, fill);
        }
    }
}

}  // namespace fanlin
