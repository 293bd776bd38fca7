// Fused separable resample (Lanczos3 / Nearest) for sm_100a: vertical stage,
// horizontal stage and the crop / letterbox / to_rgba8 epilogue in one kernel, the
// f32 intermediate held in shared memory.  See fused.h for the decomposition.
//
// Arithmetic: the same taps and normalised f32 weights as image-0.25.6
// imageops/sample.rs, accumulated with FMA in f32, vertical pass
// first.  Differences from the crate's separately-rounded multiply and add are
// below 1e-4 of a u8 step; the parity tests bound the result to 1 LSB.
//
// Source bytes are never converted to float: the u32 holding one byte IS the f32
// operand (a denormal, value b * 2^-149), and the vertical weights carry 2^100,
// the horizontal ones 2^49, so the products land back on the true scale.  Power
// of two scalings are exact, so this changes no rounding.
#include "fused.h"
#include "fused_device.cuh"
#include "kernels.h"

namespace fanlin {

namespace {

constexpr int NT = FUSED_WARPS * 32;
constexpr int S = FUSED_SLOTS;
constexpr int XEP = FUSED_XEP;
constexpr int P = 8;  // source rows in flight per warp in the vertical stage

// Raw 32-bit words behind 4 consecutive elements (post colour-op channel values)
// of a source row: 1 word when the colour op keeps the channel count, 3 for
// Rgb8 -> L8 (4 pixels), 2 for Rgba8 -> La8 (2 pixels).
template <int C, int CMEM>
struct Raw {
    static constexpr int NW = (CMEM == C) ? 1 : (CMEM == 3 ? 3 : 2);
    uint32_t w[NW];
};

// Issues the copies of the words behind elements [e, e+4) of `row` (e multiple of 4).
template <int C, int CMEM>
__device__ __forceinline__ void fetch4_async(uint32_t saddr, const uint8_t *__restrict__ row, uint32_t e, bool on) {
    constexpr int NW = Raw<C, CMEM>::NW;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(row + size_t(e) * (NW == 1 ? 1 : NW == 3 ? 3 : 2));
#pragma unroll
    for (int k = 0; k < NW; k++) cp_async4(saddr + 4 * k, p + k, on);
}
template <int C, int CMEM>
__device__ __forceinline__ void read4(const uint32_t *sp, Raw<C, CMEM> &r) {
#pragma unroll
    for (int k = 0; k < Raw<C, CMEM>::NW; k++) r.w[k] = sp[k];
}

// Raw words -> the 4 integer channel values, kept as raw bits: the u32 holding a
// byte is used directly as an f32 operand (denormal b * 2^-149).
template <int C, int CMEM, int OP>
__device__ __forceinline__ void decode4(const Raw<C, CMEM> &r, float f[4]) {
    if constexpr (CMEM == C) {
        uint32_t w = r.w[0];
        if constexpr (OP == COLOR_INVERT) {
            constexpr uint32_t m = (C == 2) ? 0x00ff00ffu : (C == 4) ? 0x00ffffffu : 0xffffffffu;  // alpha untouched
            w ^= m;
        }
        f[0] = __uint_as_float(w & 0xffu);
        f[1] = __uint_as_float(__byte_perm(w, 0, 0x4441));
        f[2] = __uint_as_float(__byte_perm(w, 0, 0x4442));
        f[3] = __uint_as_float(w >> 24);
    } else if constexpr (CMEM == 3) {  // Rgb8 -> L8: 4 pixels = 12 bytes
        const uint32_t w0 = r.w[0], w1 = r.w[1], w2 = r.w[2];
        f[0] = __uint_as_float(luma_u8(w0 & 0xff, (w0 >> 8) & 0xff, (w0 >> 16) & 0xff));
        f[1] = __uint_as_float(luma_u8(w0 >> 24, w1 & 0xff, (w1 >> 8) & 0xff));
        f[2] = __uint_as_float(luma_u8((w1 >> 16) & 0xff, w1 >> 24, w2 & 0xff));
        f[3] = __uint_as_float(luma_u8((w2 >> 8) & 0xff, (w2 >> 16) & 0xff, w2 >> 24));
    } else {  // Rgba8 -> La8: 2 pixels = 8 bytes -> (l0, a0, l1, a1)
        const uint32_t w0 = r.w[0], w1 = r.w[1];
        f[0] = __uint_as_float(luma_u8(w0 & 0xff, (w0 >> 8) & 0xff, (w0 >> 16) & 0xff));
        f[1] = __uint_as_float(w0 >> 24);
        f[2] = __uint_as_float(luma_u8(w1 & 0xff, (w1 >> 8) & 0xff, (w1 >> 16) & 0xff));
        f[3] = __uint_as_float(w1 >> 24);
    }
}

__host__ __device__ __forceinline__ uint32_t it_rows_pad(uint32_t rows) { return (rows + 3) & ~3u; }
// Shared-memory layout after the tmp tile: per-warp staging rings (each row slot: the lanes'
// A and B words, then that row's 8 weights and its info word), then the chunk's slice of the
// horizontal table.  Tables are staged because the tmp tile leaves almost no L1 behind.
constexpr uint32_t ring_row_words(int nw) { return 64u * nw + 16u; }
constexpr size_t ring_bytes(int nw) { return size_t(FUSED_WARPS) * P * ring_row_words(nw) * 4; }
constexpr size_t htab_bytes(uint32_t chunk_px) { return size_t(chunk_px) * 36; }

template <int C, int CMEM, int OP>
__global__ void __launch_bounds__(NT, 1) fused_resample_kernel(const FusedItem *__restrict__ items,
                                                               const float *__restrict__ tw,
                                                               const uint32_t *__restrict__ tinfo) {
    extern __shared__ __align__(16) float tmp[];  // [band_rows][XEP], then the per-warp staging rings
    uint32_t *ring = reinterpret_cast<uint32_t *>(tmp + size_t(it_rows_pad(items[blockIdx.x].band_rows)) * XEP);
    float *hw_s = reinterpret_cast<float *>(ring) + ring_bytes(Raw<C, CMEM>::NW) / 4;  // [chunk_px][8]
    __shared__ FusedItem it_s;
    if (threadIdx.x == 0) it_s = items[blockIdx.x];
    __syncthreads();
    const FusedItem &it = it_s;
    // warp index through a shuffle: tells the compiler it is warp-uniform, so table reads and the
    // flush branches below stay on the uniform path
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

    // ---- letterbox bars of this band (and above / below the image for the first / last band)
    if (it.epi == EPI_BLEND_FILL) {
        const uint32_t ya = it.first_band ? 0u : it.dst_y + it.band_r0;
        const uint32_t yb = it.last_band ? it.canvas_h : it.dst_y + it.band_r0 + it.band_rows;
        const uint32_t iy0 = it.dst_y + it.band_r0, iy1 = iy0 + it.band_rows;
        for (uint32_t y = ya + warp; y < yb; y += FUSED_WARPS) {
            const bool inside_rows = y >= iy0 && y < iy1;
            for (uint32_t x = lane; x < it.canvas_w; x += 32)
                if (!inside_rows || x < it.dst_x || x >= it.dst_x + it.n_cols)
                    store_rgba(it.dst + size_t(y) * it.dst_pitch + size_t(x) * 4, it.fill);
        }
    }

    // ---- item fields used inside the loops, in registers
    const uint32_t pitch = it.src_pitch, chunk_px = it.chunk_px, n_px = it.n_px, n_chunks = it.n_chunks;

    // ---- vertical-stage role of this warp: a sub-band of output rows
    const uint32_t *vi = tinfo + it.vinfo_off + warp * it.vinfo_stride;
    const uint32_t v_ny = vi[1], v_r0 = vi[2];
    const float *vw = tw + vi[3];  // this warp's weights [v_ny][8]
    const uint8_t *vsrc = it.src + size_t(it.y0 + vi[0]) * pitch;
    constexpr int NW = Raw<C, CMEM>::NW;
    constexpr uint32_t RW = ring_row_words(NW);            // words per ring row
    constexpr uint32_t EB = (NW == 1 ? 1 : NW == 3 ? 3 : 2);  // source bytes per element
    // ring row: [A words of the 32 lanes][B words of the 32 lanes] (4-byte lane stride: the
    // cp.async writes stay bank-conflict free), then 8 weights + info
    uint32_t *ring_w = ring + size_t(warp) * P * RW;
    const uint32_t *ring_l = ring_w + lane * NW;
    const uint32_t ring_sa = uint32_t(__cvta_generic_to_shared(ring_l));
    const uint32_t tab_lane = lane < 9 ? lane : 8;  // lanes 0-7 copy the weights, lane 8 (and up, redundantly) the info word
    const uint32_t tab_sa = uint32_t(__cvta_generic_to_shared(ring_w + 64 * NW + tab_lane));
    const uint32_t *tab_g0 = tab_lane < 8 ? reinterpret_cast<const uint32_t *>(vw) + tab_lane : vi + 4;
    const uint32_t tab_step = tab_lane < 8 ? S : 1;

    // ---- horizontal-stage role of this thread: one output row of the band
    const bool h_active = threadIdx.x < it.band_rows;
    float hacc[S][C];
#pragma unroll
    for (int j = 0; j < S; j++)
#pragma unroll
        for (int k = 0; k < C; k++) hacc[j][k] = 0.f;
    uint32_t h_next = 0;
    const float *hw = tw + it.hw_off;
    const uint32_t *hinfo = tinfo + it.hinfo_off;
    const float *trow = tmp + size_t(threadIdx.x) * XEP;
    const uint32_t *hinfo_s = reinterpret_cast<const uint32_t *>(hw_s + size_t(chunk_px) * S);
    const uint32_t h_cx0 = it.dst_x, h_cy = it.dst_y + it.band_r0 + threadIdx.x;

    for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
        const uint32_t cpx0 = chunk * chunk_px;  // window-relative first pixel
        const uint32_t npx = min(chunk_px, n_px - cpx0);
        const uint32_t xe = (npx * C + 3) & ~3u;  // elements, whole words
        // the chunk's slice of the horizontal table -> shared memory (lands during the vertical stage)
        {
            const uint32_t sa_w = uint32_t(__cvta_generic_to_shared(hw_s)), sa_i = uint32_t(__cvta_generic_to_shared(hinfo_s));
            const uint32_t *gw = reinterpret_cast<const uint32_t *>(hw + size_t(cpx0) * S);
            for (uint32_t k = threadIdx.x; k < npx * S; k += NT) cp_async4(sa_w + 4 * k, gw + k, true);
            for (uint32_t k = threadIdx.x; k < npx; k += NT) cp_async4(sa_i + 4 * k, hinfo + cpx0 + k, true);
            cp_async_commit();
        }
        // ================= vertical stage =================
        // Each warp streams the source rows of its sub-band top to bottom; lane l owns elements
        // [4l, 4l+4) and [128+4l, 128+4l+4) of the chunk.  Rows are copied P ahead of the FMAs
        // into the warp's ring with cp.async, together with that row's weights and info word.
        {
            const uint32_t e_row = (it.px0 + cpx0) * C;  // element index in the source row
            const uint32_t ea = 4 * lane, eb = 128 + 4 * lane;
            const bool a_on = ea < xe, b_on = eb < xe;
            // lanes past the chunk copy (and ignore) a valid word instead of being predicated
            const uint8_t *pa = vsrc + size_t(e_row + (a_on ? ea : 0)) * EB;
            const uint8_t *pb = vsrc + size_t(e_row + (b_on ? eb : 0)) * EB;
            const uint32_t *pt = tab_g0;
            float acc[S][8];
#pragma unroll
            for (int j = 0; j < S; j++)
#pragma unroll
                for (int k = 0; k < 8; k++) acc[j][k] = 0.f;
            uint32_t next_r = v_r0;
#pragma unroll
            for (int p = 0; p < P; p++) {
                if (uint32_t(p) < v_ny) {
#pragma unroll
                    for (int k = 0; k < NW; k++) {
                        cp_async4(ring_sa + (p * RW + k) * 4, pa + 4 * k, true);
                        cp_async4(ring_sa + (p * RW + 32 * NW + k) * 4, pb + 4 * k, true);
                    }
                    cp_async4(tab_sa + p * RW * 4, pt, true);
                    pa += pitch; pb += pitch; pt += tab_step;
                }
                cp_async_commit();
            }
            uint32_t slot = 0;  // word offset of row i in the ring
#pragma unroll 2
            for (uint32_t i = 0; i < v_ny; i++) {
                cp_async_wait<P - 1>();  // the group of row i has landed
                Raw<C, CMEM> qa, qb;
                read4<C, CMEM>(ring_l + slot, qa);
                read4<C, CMEM>(ring_l + slot + 32 * NW, qb);
                const uint32_t *trow_s = ring_w + slot + 64 * NW;
                const float4 w0 = *reinterpret_cast<const float4 *>(trow_s);
                const float4 w1 = *reinterpret_cast<const float4 *>(trow_s + 4);
                const uint32_t info = trow_s[8];
                float f[8];
                decode4<C, CMEM, OP>(qa, f);
                decode4<C, CMEM, OP>(qb, f + 4);
                if (i + P < v_ny) {  // refill this slot with row i + P (the reads above are consumed)
#pragma unroll
                    for (int k = 0; k < NW; k++) {
                        cp_async4(ring_sa + (slot + k) * 4, pa + 4 * k, true);
                        cp_async4(ring_sa + (slot + 32 * NW + k) * 4, pb + 4 * k, true);
                    }
                    cp_async4(tab_sa + slot * 4, pt, true);
                    pa += pitch; pb += pitch; pt += tab_step;
                }
                cp_async_commit();
                slot = slot + RW == P * RW ? 0 : slot + RW;
                const float w[S] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int j = 0; j < S; j++)  // dead slots carry weight 0 and an accumulator at 0
#pragma unroll
                    for (int k = 0; k < 8; k++) acc[j][k] = fmaf(f[k], w[j], acc[j][k]);
                const uint32_t fl = (info >> 8) & 0xffu;
                if (fl) {
#pragma unroll
                    for (int j = 0; j < S; j++) {
                        if (fl & (1u << j)) {
                            const uint32_t r = next_r + ((uint32_t(j) - next_r) & (S - 1));
                            float *t = tmp + size_t(r) * XEP;
                            if (a_on) *reinterpret_cast<float4 *>(t + ea) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                            if (b_on) *reinterpret_cast<float4 *>(t + eb) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
#pragma unroll
                            for (int k = 0; k < 8; k++) acc[j][k] = 0.f;
                        }
                    }
                    next_r += __popc(fl);
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        // ================= horizontal stage =================
        if (h_active) {
            constexpr int G = (C == 4) ? 1 : (C == 2) ? 2 : 4;  // pixels per 16-byte aligned group of the tile
            for (uint32_t xl = 0; xl < npx; xl += G) {
                float v[G * C];
#pragma unroll
                for (int q = 0; q < G * C / 4; q++) {
                    const float4 t4 = *reinterpret_cast<const float4 *>(trow + xl * C + 4 * q);
                    v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
                }
#pragma unroll
                for (int g = 0; g < G; g++) {
                    if (xl + g >= npx) break;
                    const uint32_t info = hinfo_s[xl + g];
                    const float4 w0 = *reinterpret_cast<const float4 *>(hw_s + size_t(xl + g) * S);
                    const float4 w1 = *reinterpret_cast<const float4 *>(hw_s + size_t(xl + g) * S + 4);
                    const float w[S] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int j = 0; j < S; j++)
#pragma unroll
                        for (int k = 0; k < C; k++) hacc[j][k] = fmaf(v[g * C + k], w[j], hacc[j][k]);
                    const uint32_t fl = (info >> 8) & 0xffu;
                    if (fl) {
#pragma unroll
                        for (int j = 0; j < S; j++) {
                            if (fl & (1u << j)) {
                                const uint32_t o = h_next + ((uint32_t(j) - h_next) & (S - 1));
                                uint32_t u[4] = {0, 0, 0, 0};
#pragma unroll
                                for (int k = 0; k < C; k++) { u[k] = round_u8(hacc[j][k]); hacc[j][k] = 0.f; }
                                emit_px<C, FusedItem>(it, h_cx0 + o, h_cy, u);
                            }
                        }
                        h_next += __popc(fl);
                    }
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

uint32_t fused_chunk_px(uint32_t c) { return c == 1 ? 256u : c == 2 ? 128u : c == 3 ? 84u : 64u; }
size_t fused_smem_bytes(uint32_t c, uint32_t c_mem, uint32_t band_rows) {
    const int nw = c_mem == c ? 1 : (c_mem == 3 ? 3 : 2);
    return size_t(it_rows_pad(band_rows)) * XEP * sizeof(float) + ring_bytes(nw) + ((htab_bytes(fused_chunk_px(c)) + 15) & ~size_t(15));
}
uint32_t fused_max_band(uint32_t c, uint32_t c_mem) {
    const size_t limit = 232448 - 512;  // 227 KB per CTA minus static shared memory
    const size_t fixed = fused_smem_bytes(c, c_mem, 0);
    return uint32_t((limit - fixed) / (XEP * sizeof(float))) & ~3u;
}

namespace {

template <int C, int CMEM, int OP>
void launch_variant(const FusedItem *d_items, uint32_t n_items, uint32_t max_band_rows, const float *d_w,
                    const uint32_t *d_info, LaunchCtx &lc) {
    const size_t smem = fused_smem_bytes(C, CMEM, max_band_rows);
    auto kern = fused_resample_kernel<C, CMEM, OP>;
    ensure_dynamic_smem(reinterpret_cast<const void *>(kern), fused_smem_bytes(C, CMEM, fused_max_band(C, CMEM)));
    lc.begin("fused_resample_kernel");
    kern<<<n_items, NT, smem, lc.st>>>(d_items, d_w, d_info);
    lc.end();
}

}  // namespace

int launch_fused(const FusedItem *d_items, uint32_t n_items, uint32_t variant, uint32_t max_band_rows,
                 const float *d_w, const uint32_t *d_info, LaunchCtx &lc) {
    if (n_items == 0) return 0;
    const uint32_t c = variant & 7, cmem = (variant >> 3) & 7, op = variant >> 6;
#define FANLIN_V(C_, M_, O_)                                                                \
    if (c == C_ && cmem == M_ && op == O_) {                                                \
        launch_variant<C_, M_, O_>(d_items, n_items, max_band_rows, d_w, d_info, lc);       \
        return 1;                                                                           \
    }
    FANLIN_V(1, 1, COLOR_NONE) FANLIN_V(2, 2, COLOR_NONE) FANLIN_V(3, 3, COLOR_NONE) FANLIN_V(4, 4, COLOR_NONE)
    FANLIN_V(1, 1, COLOR_INVERT) FANLIN_V(2, 2, COLOR_INVERT) FANLIN_V(3, 3, COLOR_INVERT) FANLIN_V(4, 4, COLOR_INVERT)
    FANLIN_V(1, 3, COLOR_GRAY) FANLIN_V(2, 4, COLOR_GRAY)
#undef FANLIN_V
    return -1;
}

}  // namespace fanlin
