// Separable Gaussian blur (image-0.25.6 imageops::blur as called by reference
// src/handler.rs:250-255): vertical pass u8 -> f32, horizontal pass f32 -> u8,
// 2 R + 1 taps, R = ceil(2 sigma - 0.5) (41..81 for the integer sigmas of Query::blur()), windows truncated at the borders and renormalised.
//
// Both passes stage a tile with its halo in shared memory as f32 and slide a register window
// along the filtered axis: per 8 taps a thread loads 8 new values and issues 128 FMAs for its 16
// outputs (weights broadcast from shared memory), so the kernels run near the FP32 issue rate
// instead of one shared-memory load per FMA.  The crate's border handling -- weights divided by
// the sum of the taps that fall inside the image -- is applied as one per-row / per-column
// factor after accumulating with the interior weights and zeros outside the image; that is the
// same value up to f32 rounding (the parity tests bound the result to 1 LSB).
#include "blur.h"
#include "fused_device.cuh"
#include "kernels.h"

namespace fanlin {

namespace {

constexpr int VT = 128;    // vertical kernel: threads = element columns per block
constexpr int V_ROWS = 64; // output rows per block
constexpr int HT = 256;    // horizontal kernel threads
constexpr int H_PX = 32;   // output pixels per block row
constexpr int J = 16;      // outputs per thread and window pass: J + 8 live values, 8 J FMAs per 8 loads

// Register window along the filtered axis.  x[] is a ring of J + 8 values: phase ROT sees its
// window start at ring index 8 * ROT, takes 8 new values into the 8 slots the previous phase
// left behind and issues 8 J FMAs.  Three phases bring the ring back to where it started, so a
// loop unrolled by three needs no register moves at all (shifting the window cost 16 moves per
// 128 FMAs); taps_pad % 24 leaves one or two phases for the tail, starting again at rotation 0.
constexpr int XW = J + 8;
template <int ROT, typename LoadNew>
__device__ __forceinline__ void window_phase(float (&acc)[J], float (&x)[XW], const float *w8, LoadNew load_new) {
    const float4 w0 = *reinterpret_cast<const float4 *>(w8), w1 = *reinterpret_cast<const float4 *>(w8 + 4);
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 8; i++) x[(8 * ROT + J + i) % XW] = load_new(i);
#pragma unroll
    for (int kk = 0; kk < 8; kk++)
#pragma unroll
        for (int jj = 0; jj < J; jj++) acc[jj] = fmaf(w[kk], x[(8 * ROT + jj + kk) % XW], acc[jj]);
}

// acc[j] = sum over taps_pad taps of u[k] * value(first + j + k); value(i) for i < J is preloaded here
template <typename Load>
__device__ __forceinline__ void window_pass(float (&acc)[J], const float *u_s, uint32_t taps_pad, Load value) {
    static_assert(XW == 24, "the ring returns to rotation 0 after three phases of 8");
    float x[XW];
#pragma unroll
    for (int i = 0; i < J; i++) { acc[i] = 0.f; x[i] = value(uint32_t(i)); }
    uint32_t k0 = 0;
    for (; k0 + 24 <= taps_pad; k0 += 24) {
        window_phase<0>(acc, x, u_s + k0, [&](int i) { return value(k0 + J + i); });
        window_phase<1>(acc, x, u_s + k0 + 8, [&](int i) { return value(k0 + 8 + J + i); });
        window_phase<2>(acc, x, u_s + k0 + 16, [&](int i) { return value(k0 + 16 + J + i); });
    }
    if (k0 < taps_pad) {
        window_phase<0>(acc, x, u_s + k0, [&](int i) { return value(k0 + J + i); });
        k0 += 8;
        if (k0 < taps_pad) window_phase<1>(acc, x, u_s + k0, [&](int i) { return value(k0 + J + i); });
    }
}

// ---- vertical pass: src u8 [h][pitch] -> tmp f32 [h][w*c] ------------------------------------
__global__ void __launch_bounds__(VT) blur_v_kernel(const BlurItem *__restrict__ items, const float *__restrict__ tw) {
    extern __shared__ __align__(16) float sm[];  // u[taps_pad], then tile [(V_ROWS + 2R)][VT]
    const BlurItem it = items[blockIdx.z];
    const uint32_t n_e = it.w * it.c;
    const uint32_t e0 = blockIdx.x * VT, y0 = blockIdx.y * V_ROWS;
    if (e0 >= n_e || y0 >= it.h) return;
    const uint32_t R = it.radius, taps_pad = it.taps_pad;
    float *u_s = sm;
    float *tile = sm + taps_pad;
    const uint32_t t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (uint32_t k = t; k < taps_pad; k += VT) u_s[k] = tw[it.u_off + k];
    __shared__ float corr_s[V_ROWS];  // border correction per output row of the block
    if (t < V_ROWS) corr_s[t] = y0 + t < it.h ? tw[it.corrv_off + y0 + t] : 0.f;
    // stage rows [y0 - R, y0 + V_ROWS + R + 7): one warp per row, 4 bytes per lane; the loads of
    // SB rows are issued together (one row at a time is bound by the global-load latency: measured,
    // 48 % of the kernel's samples were long-scoreboard stalls here)
    const uint32_t n_rows_tile = V_ROWS + 2 * R + 8;  // + 8: the window reads ahead inside the last (zero-weight) taps
    constexpr uint32_t SB = 8;
    const uint32_t e4 = e0 + 4 * lane;
    const bool vec = it.aligned4 && e4 + 3 < n_e;
    for (uint32_t rb0 = warp; rb0 < n_rows_tile; rb0 += SB * (VT / 32)) {
        uint32_t wv[SB];
#pragma unroll
        for (uint32_t q = 0; q < SB; q++) {
            const uint32_t r = rb0 + q * (VT / 32);
            const int y = int(y0 + r) - int(R);
            wv[q] = 0;
            if (r < n_rows_tile && y >= 0 && y < int(it.h)) {
                const uint8_t *row = it.src + size_t(y) * it.src_pitch;
                if (vec) {
                    wv[q] = __ldg(reinterpret_cast<const uint32_t *>(row + e4));
                } else {
#pragma unroll
                    for (int b = 0; b < 4; b++)
                        if (e4 + b < n_e) wv[q] |= uint32_t(row[e4 + b]) << (8 * b);
                }
            }
        }
#pragma unroll
        for (uint32_t q = 0; q < SB; q++) {
            const uint32_t r = rb0 + q * (VT / 32);
            if (r < n_rows_tile)
                *reinterpret_cast<float4 *>(tile + size_t(r) * VT + 4 * lane) =
                    make_float4(float(wv[q] & 0xff), float((wv[q] >> 8) & 0xff), float((wv[q] >> 16) & 0xff), float(wv[q] >> 24));
        }
    }
    __syncthreads();
    const uint32_t e = e0 + t;
    const float *col = tile + t;
    for (uint32_t p = 0; p < V_ROWS / J; p++) {
        const uint32_t j0 = y0 + J * p;
        if (j0 >= it.h) break;
        float acc[J];
        window_pass(acc, u_s, taps_pad, [&](uint32_t i) { return col[size_t(J * p + i) * VT]; });
        if (e < n_e) {
#pragma unroll
            for (int jj = 0; jj < J; jj++)
                if (j0 + jj < it.h) it.tmp[size_t(j0 + jj) * n_e + e] = acc[jj] * corr_s[J * p + jj];
        }
    }
}

// ---- horizontal pass: tmp f32 [h][w*c] -> dst u8 [h][w*c] --------------------------------------
template <uint32_t C>  // channels: the window's loads are rowp[i * C], constant offsets once C is known
__global__ void __launch_bounds__(HT) blur_h_kernel(const BlurItem *__restrict__ items, const float *__restrict__ tw) {
    extern __shared__ __align__(16) float sm[];  // u[taps_pad], tile [rb][pitch], then the u8 output tile
    const BlurItem it = items[blockIdx.z];
    constexpr uint32_t rb = HT / C;  // rows per block
    const uint32_t x0 = blockIdx.x * H_PX, y0 = blockIdx.y * rb;
    if (x0 >= it.w || y0 >= it.h) return;
    const uint32_t R = it.radius, taps_pad = it.taps_pad;
    const uint32_t n_px_tile = H_PX + 2 * R + 8;
    uint32_t pitch = n_px_tile * C;             // floats per tile row, == C (mod 32): conflict-free for lanes = (row, channel)
    pitch += (C + 32 - (pitch & 31)) & 31;
    float *u_s = sm;
    float *tile = sm + taps_pad;
    uint8_t *out_s = reinterpret_cast<uint8_t *>(tile + size_t(rb) * pitch);  // [rb][H_PX * C + 4]: rows on different banks
    const uint32_t out_pitch = H_PX * C + 4;
    const uint32_t t = threadIdx.x;
    for (uint32_t k = t; k < taps_pad; k += HT) u_s[k] = tw[it.u_off + k];
    __shared__ float corr_s[H_PX];  // border correction per output pixel of the block
    if (t < H_PX) corr_s[t] = x0 + t < it.w ? tw[it.corrh_off + x0 + t] : 0.f;
    const uint32_t n_e = it.w * C;
    // stage: tile row r <- tmp[y0 + r][(x0 - R) * C ...), zeros outside the image
    const uint32_t row_elems = n_px_tile * C;
    // one warp per row: coalesced, no index division.  cp.async (global -> shared, no register hop)
    // keeps every load of the block in flight at once; a row-at-a-time register copy was bound by
    // the load latency (54 % long-scoreboard samples).
    for (uint32_t r = t >> 5; r < rb; r += HT / 32) {
        const uint32_t y = y0 + r;
        const float *src = it.tmp + size_t(y) * n_e;
        const int ge0 = int(x0 * C) - int(R * C);
        float *trow = tile + size_t(r) * pitch;
        // four floats per copy when both sides are 16-byte aligned (always for RGBA rows of even width)
        const bool quad = (row_elems & 3) == 0 && (ge0 & 3) == 0 && (n_e & 3) == 0 &&
                          ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(trow)) & 15) == 0;
        if (quad) {
            for (uint32_t i = 4 * (t & 31); i < row_elems; i += 128) {
                const int ge = ge0 + int(i);  // ge, n_e multiples of 4: a group is inside or outside as a whole
                if (y < it.h && ge >= 0 && ge < int(n_e)) cp_async16(smem_u32(trow + i), src + ge);
                else *reinterpret_cast<float4 *>(trow + i) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (uint32_t i = t & 31; i < row_elems; i += 32) {
                const int ge = ge0 + int(i);
                if (y < it.h && ge >= 0 && ge < int(n_e)) cp_async4(smem_u32(trow + i), src + ge, true);
                else trow[i] = 0.f;
            }
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const uint32_t rows_here = min(rb, it.h - y0);
    if (t < rb * C) {
        const uint32_t r = t / C, ch = t - r * C;
        const float *rowp = tile + size_t(r) * pitch + ch;
        for (uint32_t p = 0; p < H_PX / J; p++) {
            float acc[J];
            window_pass(acc, u_s, taps_pad, [&](uint32_t i) { return rowp[size_t(J * p + i) * C]; });
#pragma unroll
            for (int jj = 0; jj < J; jj++) {
                const float cf = corr_s[J * p + jj];
                out_s[size_t(r) * out_pitch + (J * p + jj) * C + ch] = uint8_t(round_u8(acc[jj] * cf));
            }
        }
    }
    __syncthreads();
    // coalesced store of the u8 tile
    const uint32_t px_here = min(uint32_t(H_PX), it.w - x0), bytes_row = px_here * C;
    for (uint32_t r = t >> 5; r < rows_here; r += HT / 32) {  // one warp per row
        uint8_t *d = it.dst + size_t(y0 + r) * n_e + size_t(x0) * C;
        for (uint32_t i = t & 31; i < bytes_row; i += 32) d[i] = out_s[size_t(r) * out_pitch + i];
    }
}

}  // namespace

size_t blur_v_smem(uint32_t radius, uint32_t taps_pad) { return (size_t(taps_pad) + size_t(V_ROWS + 2 * radius + 8) * VT) * 4; }
size_t blur_h_smem(uint32_t radius, uint32_t taps_pad, uint32_t c) {
    const uint32_t rb = HT / c;
    uint32_t pitch = (H_PX + 2 * radius + 8) * c;
    pitch += (c + 32 - (pitch & 31)) & 31;
    return (size_t(taps_pad) + size_t(rb) * pitch) * 4 + size_t(rb) * (H_PX * c + 4) + 16;
}

int launch_blur(const BlurItem *d_items, uint32_t n_items, uint32_t max_w, uint32_t max_h, uint32_t c, uint32_t radius,
                uint32_t taps_pad, const float *d_w, bool skip_v, LaunchCtx &lc) {
    if (n_items == 0) return 0;
    const size_t sv = blur_v_smem(radius, taps_pad), sh = blur_h_smem(radius, taps_pad, c);
    ensure_dynamic_smem(reinterpret_cast<const void *>(blur_v_kernel), sv);
    auto hk = c == 1 ? blur_h_kernel<1> : c == 2 ? blur_h_kernel<2> : c == 3 ? blur_h_kernel<3> : blur_h_kernel<4>;
    ensure_dynamic_smem(reinterpret_cast<const void *>(hk), sh);
    const uint32_t rb = HT / c;
    if (!skip_v) {  // else the f32 intermediate was written by blur_v_tc_kernel
        lc.begin("blur_v_kernel");
        blur_v_kernel<<<dim3((max_w * c + VT - 1) / VT, (max_h + V_ROWS - 1) / V_ROWS, n_items), VT, sv, lc.st>>>(d_items, d_w);
        lc.end();
    }
    lc.begin("blur_h_kernel");
    hk<<<dim3((max_w + H_PX - 1) / H_PX, (max_h + rb - 1) / rb, n_items), HT, sh, lc.st>>>(d_items, d_w);
    lc.end();
    return skip_v ? 1 : 2;
}

}  // namespace fanlin
