// Host side of the tensor-core fused resample: bands, 32-row output groups, the
// s8 digit tiles of the vertical weights and the horizontal scatter table.
#include "fused_tc.h"

#include <cstdio>
#include <cstdlib>

#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>

namespace fanlin {

namespace {

struct TcBand {
    uint32_t r0, rows, grp_off, n_groups, kg_max, grp_rows;
    uint32_t inv_off;  // u32 offset of the band's per-row constants 255 * (sum of the row's quantised weights) * 2^-s as f32 bits
};
struct TcGeom {
    bool ok = false;
    bool hm = false;        // horizontal stage on the tensor cores
    bool ring = false;      // ... with its sums accumulated in TMEM across chunks (fused_resample_tc3_kernel)
    uint32_t ring_cols = 0, n_vr = 0, stage_stride = 0, wh_bytes = 0;
    uint32_t hrec_off = 0;  // its per-chunk records
    uint32_t b0 = 0, n_chunks = 0, chunk_off = 0, hw_off = 0, hinfo_off = 0, cpre_off = 0, out_stride = 1;
    float scale = 1.f;
    std::vector<TcBand> bands;
};
using TcKey = std::tuple<const AxisTable *, const AxisTable *, uint32_t, uint32_t, uint32_t>;

// Floats per tile column.  Odd, so the drain's lanes (= consecutive tile columns) hit distinct
// banks; and congruent to ~32/c (mod 32), so the horizontal stage's lanes (= (row, channel)
// pairs, channel fastest) do too: bank = channel * r_pad + row.
uint32_t r_pad_for(uint32_t rows, uint32_t c) {
    const uint32_t want = c == 1 ? 1u : c == 2 ? 17u : c == 3 ? 11u : 9u;
    if (c == 1) return rows | 1u;
    return rows + ((want + 32u - (rows & 31u)) & 31u);
}

// Horizontal scatter table of the tensor-core kernel, one record per PAIR of source pixels
// (x0, x0 + 1; the second is absent in a chunk's odd last pair): w[p][0..8) / w[p][8..16) are
// the weights of the two pixels for the outputs base(p) + j, j < 8, where base(p) is the
// first output whose window has not ended before x0; cnt[p] = outputs whose window ends inside
// the pair.  The kernel keeps the 8 accumulators as a shift register: after the pair it writes
// out and shifts cnt[p] times, so slot j always means "j-th unfinished output" and no slot
// index is ever computed.
struct PxPair {
    uint32_t x0;
    bool two;
};
bool tc_pair_table(const AxisTable &t, uint32_t o0, uint32_t n, const std::vector<PxPair> &pairs, float *w, uint32_t *cnt) {
    const uint32_t o_end = o0 + n;
    uint32_t base = o0;
    for (size_t p = 0; p < pairs.size(); p++) {
        const uint32_t x0 = pairs[p].x0, x1 = x0 + 1, x_last = pairs[p].two ? x1 : x0;
        for (uint32_t o = base; o < o_end && t.entries[o].left <= x_last; o++) {
            const TapEntry &e = t.entries[o];
            const uint32_t j = o - base;
            if (j >= FUSED_SLOTS) return false;  // more than 8 unfinished outputs
            if (x0 >= e.left && x0 < e.left + e.count) w[p * 16 + j] = t.weights[e.woff + (x0 - e.left)];
            if (pairs[p].two && x1 >= e.left && x1 < e.left + e.count) w[p * 16 + 8 + j] = t.weights[e.woff + (x1 - e.left)];
        }
        uint32_t c = 0;
        while (base < o_end && t.entries[base].left + t.entries[base].count <= x_last + 1) { base++; c++; }
        cnt[p] = c;
    }
    return base == o_end;  // every output was finished, in order
}

}  // namespace

// A chunk is 128 source bytes per row; with the <= c - 1 bytes carried over from the chunk before
// it holds at most (128 + c - 1) / c whole pixels.
uint32_t fused_tc_max_pairs(uint32_t c) { return ((TC_M + c - 1) / c + 1) / 2; }

// dynamic shared memory the kernel may ask for: the opt-in maximum minus its static shared memory
#define TC_SMEM_LIMIT fused_tc_smem_limit()

size_t fused_tc_smem_bytes(uint32_t c, uint32_t band_rows, uint32_t kg_max, uint32_t out_stride, uint32_t n_a) {
    const size_t tmp = ((size_t(TC_M + c - 1) * r_pad_for(band_rows, c) + 3) & ~size_t(3)) * 4;  // + the columns carried over from the previous chunk
    const size_t a = size_t(n_a) * kg_max * TC_M, b = 2 * size_t(TC_N) * kg_max;       // source slots, two weight-tile slots
    const size_t ht = 2 * ((size_t(fused_tc_max_pairs(c)) * 68 + 15) & ~size_t(15));  // two copies of the chunk's horizontal table slice
    const size_t out = size_t(band_rows) * out_stride * 4;                            // output pixels finished in the chunk
    return tmp + a + b + ht + out + 1024;  // + slack to align the slots to 1024 bytes
}

// Source slots take the shared memory the band leaves: the more groups are in flight, the
// better the loads cover the L2 / HBM latency (2 = ping-pong, measured latency-bound).
uint32_t fused_tc_source_slots(uint32_t c, uint32_t band_rows, uint32_t kg_max, uint32_t out_stride) {
    uint32_t n = 4;
    while (n > 0 && fused_tc_smem_bytes(c, band_rows, kg_max, out_stride, n) > TC_SMEM_LIMIT) n--;
    return n;
}

uint32_t fused_tc_max_band(uint32_t c, uint32_t out_stride) {
    // the horizontal stage maps (row pair, channel) to the threads of TC_H_WARPS warps
    uint32_t rows = std::min(192u, 2 * (TC_H_WARPS * 32 / c));
    while (rows > 8 && fused_tc_source_slots(c, rows, TC_KG_MAX, out_stride) < 2) rows--;
    return rows;
}

struct FusedTcCache {
    std::map<TcKey, TcGeom> geoms;
    bool allow_hmma = true;
};
FusedTcCache *fused_tc_cache_new(bool allow_hmma) {
    FusedTcCache *c = new FusedTcCache();
    c->allow_hmma = allow_hmma;
    return c;
}

// ---- tensor-core horizontal stage ---------------------------------------------------------------
size_t fused_tc2_smem_bytes(uint32_t c, uint32_t n_groups, uint32_t kg_max, uint32_t n_a, uint32_t n_wh, uint32_t band_rows, uint32_t out_stride) {
    const size_t out = size_t(band_rows) * out_stride * 4;  // output pixels finished in a chunk
    const size_t a = size_t(n_a) * kg_max * TC_M, b = 2 * size_t(TC_N) * kg_max;  // source slots, two vertical weight-tile slots
    const size_t wh = size_t(n_wh) * fused_tc2_n(c) * 512;                        // hi + lo tiles [N2][128] f16
    const size_t t = 2 * size_t(n_groups) * 32 * 256;                             // T hi + lo: [32 rows per group][128] f16
    return a + b + wh + t + out;  // the kernel's static shared memory is a multiple of 1024 bytes: the dynamic part starts aligned
}

size_t fused_tc3_smem_bytes(uint32_t n_groups, uint32_t kg_max, uint32_t n_a, uint32_t n_wh, uint32_t wh_bytes, uint32_t stage_stride) {
    const size_t a = size_t(n_a) * kg_max * TC_M, b = 2 * size_t(TC_N) * kg_max;  // source slots, two vertical weight-tile slots
    const size_t t = 2 * size_t(n_groups) * 32 * 256;                             // T hi + lo
    const size_t stage = 8 * ((16 + size_t(32) * stage_stride * 4 + 15) & ~size_t(15));  // per (row tile, lane quarter): guard + [32 rows][stage_stride words]
    return a + b + t + size_t(n_wh) * wh_bytes + stage;
}

size_t fused_tc_item_smem(const FusedTcItem &it) {
    if (it.hmma == 2) return fused_tc3_smem_bytes(it.n_groups, it.kg_max, it.n_a, it.n_wh, it.wh_bytes, it.stage_stride);
    return it.hmma ? fused_tc2_smem_bytes(it.c, it.n_groups, it.kg_max, it.n_a, it.n_wh, it.band_rows, it.out_stride)
                   : fused_tc_smem_bytes(it.c, it.band_rows, it.kg_max, it.out_stride, it.n_a);
}
void fused_tc_cache_free(FusedTcCache *c) { delete c; }

bool fused_tc_eligible(const StagePlan &s, const fanlin_job &job) {
    if (!s.present || !s.separable || !s.vtab || !s.htab) return false;
    if (s.n_rows == 0 || s.n_cols == 0) return false;
    if (s.v_kind != KIND_LANCZOS3) return false;  // Nearest is a gather (bit-exact on the CUDA-core path); blur has its own kernel
    // Inverse rides on the vertical pass: sum q (255 - x) = 255 sum q - sum q x, exact in the integer contraction, so the
    // consumers subtract the vertical result from a per-row constant (colour channels only) instead of a pass over the
    // source inverting its bytes.  Grayscale changes the channel count and keeps its pass.
    if ((s.color_op != COLOR_NONE && s.color_op != COLOR_INVERT) || s.c_mem != s.c) return false;
    const uint32_t pitch = s.src_is_input ? (job.src_pitch ? job.src_pitch : job.src_w * job.src_channels) : (s.in_pitch ? s.in_pitch : s.in_w * s.c_mem);
    if (pitch % 16 != 0) return false;  // TMA: the row stride is a multiple of 16 bytes
    if (s.src_is_input && (reinterpret_cast<uintptr_t>(job.src) & 15)) return false;
    // a 32-row output group must fit 256 source rows, and the horizontal pass 8 live outputs
    const double ratio = double(s.in_h) / double(std::max(1u, s.v_out));
    if (7 * ratio + s.vtab->max_taps > TC_KG_MAX) return false;
    if (s.htab->max_taps > 8 * std::max(1u, s.in_w / std::max(1u, s.h_out)) + 8) return false;
    return true;
}

// One band of output rows [oy0 + r0, + rows) of a vertical filter: its groups of <= 32 rows, their
// source-row windows (in 32-row K steps) and their weight-digit tiles.  min_grp: smallest group size
// tried -- every MMA costs the same ~112 clk floor whatever N is, so the group size that minimises
// the number of K steps over the band wins (C2: 29 rows -> 6 x 7 instead of 6 x 8 at 32).
static bool build_band(const AxisTable &vt, uint32_t oy0, uint32_t r0, uint32_t rows, int sh, uint32_t min_grp, FusedTables *tabs,
                       FusedTcTables *tct, TcBand *out) {
    TcBand bt{};
    bt.r0 = r0;
    bt.rows = rows;
    // source rows spanned by output rows [ra, rb) of the band
    auto span = [&](uint32_t ra, uint32_t rb, uint32_t *k0) {
        *k0 = vt.entries[oy0 + r0 + ra].left;
        uint32_t k1 = 0;
        for (uint32_t r = ra; r < rb; r++) {
            const TapEntry &e = vt.entries[oy0 + r0 + r];
            k1 = std::max(k1, e.left + e.count);
        }
        return k1 - *k0;
    };
    uint32_t best_r = 0, best_cost = ~0u;
    for (uint32_t gr = TC_GROUP_ROWS; gr >= min_grp; gr--) {
        uint32_t cost = 0;
        bool fits = true;
        for (uint32_t ra = 0; ra < bt.rows && fits; ra += gr) {
            uint32_t k0;
            const uint32_t kspan = span(ra, std::min(bt.rows, ra + gr), &k0);
            fits = kspan <= TC_KG_MAX;
            cost += (kspan + 31) / 32;
        }
        if (fits && cost < best_cost) { best_cost = cost; best_r = gr; }
    }
    if (!best_r || (bt.rows + best_r - 1) / best_r > 24) return false;  // the kernels keep <= 24 group records in shared memory
    bt.grp_rows = best_r;
    bt.n_groups = (bt.rows + best_r - 1) / best_r;
    bt.grp_off = uint32_t(tabs->info.size());
    tabs->info.resize(tabs->info.size() + size_t(bt.n_groups) * 4, 0u);
    for (uint32_t gi = 0; gi < bt.n_groups; gi++) {
        const uint32_t ra = gi * best_r, rb = std::min(bt.rows, ra + best_r);
        uint32_t k0;
        const uint32_t kg = (span(ra, rb, &k0) + 31) / 32 * 32;
        bt.kg_max = std::max(bt.kg_max, kg);
        const size_t b_off = (tct->b.size() + 127) & ~size_t(127);
        tct->b.resize(b_off + size_t(TC_N) * kg, 0);
        int8_t *tile = reinterpret_cast<int8_t *>(&tct->b[b_off]);
        for (uint32_t r = ra; r < rb; r++) {
            const TapEntry &e = vt.entries[oy0 + bt.r0 + r];
            for (uint32_t t = 0; t < e.count; t++) {
                const long q = std::lround(std::ldexp(double(vt.weights[e.woff + t]), sh));
                const long lo = ((q + 64) & 127) - 64;
                const long q1 = (q - lo) / 128;
                const long mid = ((q1 + 64) & 127) - 64;
                const long hi = (q1 - mid) / 128;
                if (hi < -128 || hi > 127) return false;
                const uint32_t k = e.left + t - k0, j = r - ra;
                const long dig[3] = {hi, mid, lo};
                for (uint32_t d = 0; d < 3; d++) {
                    const uint32_t n = d * TC_GROUP_ROWS + j;  // B row: digit-major
                    tile[(size_t(n / 8) * (kg / 16) + k / 16) * 128 + (n % 8) * 16 + k % 16] = int8_t(dig[d]);
                }
            }
        }
        uint32_t *gw = &tabs->info[bt.grp_off + size_t(gi) * 4];
        gw[0] = k0; gw[1] = kg; gw[2] = uint32_t(b_off); gw[3] = rb - ra;
    }
    // inverse on load: per band row, 255 * sum of its quantised weights, at the scale of the recombined vertical result
    bt.inv_off = uint32_t(tabs->info.size());
    for (uint32_t r = 0; r < bt.rows; r++) {
        const TapEntry &e = vt.entries[oy0 + bt.r0 + r];
        double sq = 0.0;
        for (uint32_t t = 0; t < e.count; t++) sq += double(std::lround(std::ldexp(double(vt.weights[e.woff + t]), sh)));
        const float k = float(std::ldexp(255.0 * sq, -sh));
        uint32_t bits;
        std::memcpy(&bits, &k, 4);
        tabs->info.push_back(bits);
    }
    *out = bt;
    return true;
}

// q = round(w * 2^sh) must fit three signed base-128 digits
static int weight_shift(const AxisTable &vt) {
    float maxw = 0.f;
    for (float w : vt.weights) maxw = std::max(maxw, std::fabs(w));
    int sh = 30;
    while (sh > 0 && std::ldexp(double(maxw), sh) > 2080000.0) sh--;
    return sh;
}

// Group size and K extent build_band would choose for a band, without building anything.
static bool plan_band(const AxisTable &vt, uint32_t oy0, uint32_t r0, uint32_t rows, uint32_t min_grp, uint32_t *n_groups, uint32_t *kg_max) {
    uint32_t best_r = 0, best_cost = ~0u, best_kg = 0;
    for (uint32_t gr = TC_GROUP_ROWS; gr >= min_grp; gr--) {
        uint32_t cost = 0, kgm = 0;
        bool fits = true;
        for (uint32_t ra = 0; ra < rows && fits; ra += gr) {
            const uint32_t rb = std::min(rows, ra + gr), k0 = vt.entries[oy0 + r0 + ra].left;
            uint32_t k1 = 0;
            for (uint32_t r = ra; r < rb; r++) k1 = std::max(k1, vt.entries[oy0 + r0 + r].left + vt.entries[oy0 + r0 + r].count);
            fits = k1 - k0 <= TC_KG_MAX;
            cost += (k1 - k0 + 31) / 32;
            kgm = std::max(kgm, (k1 - k0 + 31) / 32 * 32);
        }
        if (fits && cost < best_cost) { best_cost = cost; best_r = gr; best_kg = kgm; }
    }
    if (!best_r) return false;
    *n_groups = (rows + best_r - 1) / best_r;
    *kg_max = best_kg;
    return true;
}

// Horizontal stage on the tensor cores: per chunk one f16 weight tile pair (hi, lo; K-major core-matrix
// layout, [N2][128 tile columns]) that maps the chunk's 128 tile columns (byte b = pixel b / c, channel
// b % c) to the accumulator columns (ring position of the output) * c + channel, ring = N2 / c outputs:
// the outputs a chunk touches and those still unfinished from earlier chunks must fit the ring.
// Records per chunk: {tile byte offset, first output it finishes (relative to ox0), outputs finished}.
static bool build_hmma(const StagePlan &s, TcGeom &g, FusedTables *tabs, FusedTcTables *tct) {
    const AxisTable &t = *s.htab;
    const uint32_t C = s.c, N2 = fused_tc2_n(C), RP = N2 / C;
    const uint32_t o0 = s.ox0, o_end = s.ox0 + s.n_cols;
    {  // dry run: the ring condition
        uint32_t fin = o0, touch = o0;
        for (uint32_t ch = 0; ch < g.n_chunks; ch++) {
            const uint32_t b1 = g.b0 + TC_M * (ch + 1), x_hi = (b1 - 1) / C;
            while (touch < o_end && t.entries[touch].left <= x_hi) touch++;
            if (touch - fin > RP) return false;
            while (fin < touch && (t.entries[fin].left + t.entries[fin].count) * C <= b1) fin++;
        }
        if (fin != o_end) return false;
    }
    std::vector<uint32_t> rec(size_t(3) * g.n_chunks);
    // f16 halves of every tap once (a tap is written c times per chunk it falls into)
    std::vector<uint16_t> w_hi(t.weights.size()), w_lo(t.weights.size());
    for (uint32_t o = o0; o < o_end; o++) {
        const TapEntry &e = t.entries[o];
        for (uint32_t tt = 0; tt < e.count; tt++) {
            const float w = t.weights[e.woff + tt] * TC2_WSCALE;
            const __half wh = __float2half_rn(w);
            w_hi[e.woff + tt] = __half_as_ushort(wh);
            w_lo[e.woff + tt] = __half_as_ushort(__float2half_rn(w - __half2float(wh)));
        }
    }
    uint32_t fin = o0, touch = o0;
    for (uint32_t ch = 0; ch < g.n_chunks; ch++) {
        const uint32_t b0 = g.b0 + TC_M * ch, b1 = b0 + TC_M, x_hi = (b1 - 1) / C;
        while (touch < o_end && t.entries[touch].left <= x_hi) touch++;
        const size_t off = (tct->b.size() + 127) & ~size_t(127);
        tct->b.resize(off + size_t(N2) * 512, 0);
        uint16_t *hi = reinterpret_cast<uint16_t *>(&tct->b[off]), *lo = hi + size_t(N2) * 128;
        for (uint32_t o = fin; o < touch; o++) {
            const TapEntry &e = t.entries[o];
            // taps whose pixel has a byte in [b0, b1)
            const uint32_t ta = e.left * C + C > b0 ? 0u : (b0 - (e.left * C + C - 1) + C - 1) / C;
            for (uint32_t tt = ta; tt < e.count && (e.left + tt) * C < b1; tt++) {
                const uint16_t wh = w_hi[e.woff + tt], wl = w_lo[e.woff + tt];
                for (uint32_t chn = 0; chn < C; chn++) {
                    const uint32_t byte = (e.left + tt) * C + chn;
                    if (byte < b0 || byte >= b1) continue;
                    const uint32_t k = byte - b0, n = ((o - o0) % RP) * C + chn;
                    const size_t at = (size_t(n / 8) * 16 + k / 8) * 64 + (n % 8) * 8 + k % 8;  // in f16 elements
                    hi[at] = wh;
                    lo[at] = wl;
                }
            }
        }
        const uint32_t first = fin;
        while (fin < touch && (t.entries[fin].left + t.entries[fin].count) * C <= b1) fin++;
        rec[3 * ch] = uint32_t(off); rec[3 * ch + 1] = first - o0; rec[3 * ch + 2] = fin - first;
    }
    g.hrec_off = uint32_t(tabs->info.size());
    tabs->info.insert(tabs->info.end(), rec.begin(), rec.end());
    return true;
}

// Horizontal stage with the sums accumulated in TMEM (fused_resample_tc3_kernel): output pixel o lives in ring slot
// (o - ox0) mod RP of its row tile, at accumulator columns slot * c + channel; a chunk's MMAs add into the window of slots
// its 128 tile columns touch (split in two where the window wraps), the consumers drain, round and zero the slots whose
// taps ended in the chunk.  The ring only has to hold the outputs a chunk touches, so ratios near 2 (C1, C3) fit.
// Per chunk: one f16 tile pair hi | lo [n_total][128] (K-major core matrices; row n = window column n) and a record of
// 8 words {tile offset, first output finished (relative to ox0), outputs finished, window column, window columns,
// ring slot of the first finished output, K-step ranges of the two window pieces (4 bytes: lo, hi, lo, hi), 0}.
static uint32_t ring_slot_align(uint32_t c) { return c == 4 ? 4u : c == 2 ? 8u : 16u; }  // RP * c is a multiple of 16

static bool build_ring(const StagePlan &s, TcGeom &g, FusedTables *tabs, FusedTcTables *tct) {
    const AxisTable &t = *s.htab;
    const uint32_t C = s.c, o0 = s.ox0, o_end = s.ox0 + s.n_cols;
    uint32_t live_max = 0, fin_max = 0;
    {
        uint32_t fin = o0, touch = o0;
        for (uint32_t ch = 0; ch < g.n_chunks; ch++) {
            const uint32_t b1 = g.b0 + TC_M * (ch + 1), x_hi = (b1 - 1) / C, first = fin;
            while (touch < o_end && t.entries[touch].left <= x_hi) touch++;
            live_max = std::max(live_max, touch - fin);
            while (fin < touch && (t.entries[fin].left + t.entries[fin].count) * C <= b1) fin++;
            fin_max = std::max(fin_max, fin - first);
        }
        if (fin != o_end) return false;
    }
    const uint32_t al = ring_slot_align(C);
    uint32_t RP = (std::max(live_max, 1u) + al - 1) / al * al, RINGC = RP * C;
    // the window of a chunk starts on a multiple of 16 columns: with that padding it must still fit the ring
    for (;; RP += al, RINGC = RP * C) {
        if (RINGC > 256) return false;
        uint32_t fin = o0, touch = o0, widest = 0;
        for (uint32_t ch = 0; ch < g.n_chunks; ch++) {
            const uint32_t b1 = g.b0 + TC_M * (ch + 1), x_hi = (b1 - 1) / C;
            while (touch < o_end && t.entries[touch].left <= x_hi) touch++;
            const uint32_t u0 = ((fin - o0) % RP) * C, len = (touch - fin) * C;
            widest = std::max(widest, (u0 + len - (u0 & ~15u) + 15) & ~15u);
            while (fin < touch && (t.entries[fin].left + t.entries[fin].count) * C <= b1) fin++;
        }
        if (widest <= RINGC) break;
    }
    g.ring_cols = RINGC;
    // words per staged row, == 4 (mod 8): 16-byte aligned rows whose four-word stores (lanes = rows) fall on distinct banks
    g.stage_stride = (fin_max * s.c_out + 3) / 4 + 1;
    // (RGBA in, RGBA out stages four pixels per 16-byte store; everything else stages words or bytes, for which an ODD stride
    // is conflict-free and up to seven words per row shorter -- the 7 KB that let a whole C1 image be ONE band of two row tiles)
    if (C == 4 && s.c_out == 4) g.stage_stride += (4 + 8 - (g.stage_stride & 7)) & 7;
    else g.stage_stride |= 1u;
    std::vector<uint16_t> w_hi(t.weights.size()), w_lo(t.weights.size());
    for (uint32_t o = o0; o < o_end; o++) {
        const TapEntry &e = t.entries[o];
        for (uint32_t tt = 0; tt < e.count; tt++) {
            const float w = t.weights[e.woff + tt] * TC2_WSCALE;
            const __half wh = __float2half_rn(w);
            w_hi[e.woff + tt] = __half_as_ushort(wh);
            w_lo[e.woff + tt] = __half_as_ushort(__float2half_rn(w - __half2float(wh)));
        }
    }
    std::vector<uint32_t> rec(size_t(8) * g.n_chunks, 0u);
    uint32_t fin = o0, touch = o0, n_max = 0;
    for (uint32_t ch = 0; ch < g.n_chunks; ch++) {
        const uint32_t b0 = g.b0 + TC_M * ch, b1 = b0 + TC_M, x_hi = (b1 - 1) / C;
        while (touch < o_end && t.entries[touch].left <= x_hi) touch++;
        const uint32_t u0 = ((fin - o0) % RP) * C, len = (touch - fin) * C;  // window in ring columns, before wrapping
        const uint32_t w0 = u0 & ~15u;
        uint32_t n_total = (u0 + len - w0 + 15) & ~15u;
        if (n_total == 0) n_total = 16;  // a chunk no output touches still issues its (all-zero) MMAs
        if (n_total > RINGC) return false;
        n_max = std::max(n_max, n_total);
        const size_t off = (tct->b.size() + 127) & ~size_t(127);
        tct->b.resize(off + size_t(n_total) * 512, 0);
        uint16_t *hi = reinterpret_cast<uint16_t *>(&tct->b[off]), *lo = hi + size_t(n_total) * 128;
        for (uint32_t o = fin; o < touch; o++) {
            const TapEntry &e = t.entries[o];
            const uint32_t ta = e.left * C + C > b0 ? 0u : (b0 - (e.left * C + C - 1) + C - 1) / C;
            for (uint32_t tt = ta; tt < e.count && (e.left + tt) * C < b1; tt++) {
                for (uint32_t chn = 0; chn < C; chn++) {
                    const uint32_t byte = (e.left + tt) * C + chn;
                    if (byte < b0 || byte >= b1) continue;
                    const uint32_t k = byte - b0, n = u0 - w0 + (o - fin) * C + chn;
                    const size_t at = (size_t(n / 8) * 16 + k / 8) * 64 + (n % 8) * 8 + k % 8;  // in f16 elements
                    hi[at] = w_hi[e.woff + tt];
                    lo[at] = w_lo[e.woff + tt];
                }
            }
        }
        // K steps (16 tile columns each) in which the two pieces of the window -- rows [0, n1) in front of the ring's end, rows
        // [n1, n_total) wrapped to its start -- have any weight: outputs and tile columns both run left to right, so the first
        // piece needs the leading steps only and the wrapped piece the trailing ones.  The kernel skips the rest (they would
        // add zeros, at the price of re-reading the T tile).
        const uint32_t n1 = std::min(n_total, RINGC - w0);
        uint32_t klo[2] = {8, 8}, khi[2] = {0, 0};
        for (uint32_t n = 0; n < n_total; n++) {
            const uint32_t pc = n < n1 ? 0u : 1u;
            for (uint32_t k = 0; k < 128; k++) {
                const size_t at = (size_t(n / 8) * 16 + k / 8) * 64 + (n % 8) * 8 + k % 8;
                if (hi[at] | lo[at]) { klo[pc] = std::min(klo[pc], k / 16); khi[pc] = std::max(khi[pc], k / 16 + 1); }
            }
        }
        for (int pc = 0; pc < 2; pc++) if (khi[pc] == 0) klo[pc] = 0;  // no weight at all: an empty range
        const uint32_t first = fin;
        while (fin < touch && (t.entries[fin].left + t.entries[fin].count) * C <= b1) fin++;
        uint32_t *r = &rec[size_t(8) * ch];
        r[0] = uint32_t(off); r[1] = first - o0; r[2] = fin - first; r[3] = w0; r[4] = n_total; r[5] = (first - o0) % RP;
        r[6] = klo[0] | khi[0] << 8 | klo[1] << 16 | khi[1] << 24;
    }
    g.wh_bytes = n_max * 512;
    g.hrec_off = uint32_t(tabs->info.size());
    tabs->info.insert(tabs->info.end(), rec.begin(), rec.end());
    return true;
}

// Bands, TMEM and shared-memory plan of the ring variant; fills g on success.
static bool try_ring(const StagePlan &s, TcGeom &g, FusedTables *tabs, FusedTcTables *tct) {
    const size_t info_mark = tabs->info.size(), b_mark = tct->b.size();
    if (!build_ring(s, g, tabs, tct)) { tabs->info.resize(info_mark); tct->b.resize(b_mark); return false; }
    const int sh = weight_shift(*s.vtab);
    bool fits = false;
    uint32_t band_rows = s.n_rows;
    for (uint32_t n_bands = std::max(1u, s.min_bands); n_bands <= 64 && !fits; n_bands++) {
        band_rows = (s.n_rows + n_bands - 1) / n_bands;
        fits = true;
        for (uint32_t r0 = 0; r0 < s.n_rows && fits; r0 += band_rows) {
            uint32_t ng = 0, kgm = 0;
            fits = plan_band(*s.vtab, s.oy0, r0, std::min(band_rows, s.n_rows - r0), 8, &ng, &kgm) && ng <= 8;
            if (!fits) break;
            const uint32_t n_mt = (ng + 3) / 4;
            fits = n_mt * g.ring_cols + 2 * TC_N <= 512 &&
                   fused_tc3_smem_bytes(ng, kgm, 2, 1, g.wh_bytes, g.stage_stride) <= TC_SMEM_LIMIT;
            if (std::getenv("FANLIN_TC_DEBUG"))
                std::fprintf(stderr, "try_ring: %u bands of %u rows: %u groups, kg_max %u, ring %u, wh %u B, stage stride %u -> %zu B of %zu: %s\n", n_bands, band_rows, ng, kgm,
                             g.ring_cols, g.wh_bytes, g.stage_stride, fused_tc3_smem_bytes(ng, kgm, 2, 1, g.wh_bytes, g.stage_stride), size_t(TC_SMEM_LIMIT), fits ? "fits" : "no");
        }
    }
    if (!fits) { tabs->info.resize(info_mark); tct->b.resize(b_mark); return false; }
    g.scale = std::ldexp(1.0f, -sh);
    g.out_stride = g.stage_stride;
    for (uint32_t r0 = 0; r0 < s.n_rows; r0 += band_rows) {
        TcBand bt{};
        if (!build_band(*s.vtab, s.oy0, r0, std::min(band_rows, s.n_rows - r0), sh, 8, tabs, tct, &bt)) {
            g.bands.clear();
            return false;  // (tables appended so far stay: unreferenced bytes)
        }
        g.bands.push_back(bt);
    }
    g.hm = true; g.ring = true; g.ok = true;
    return true;
}

static const TcGeom &geom_of(const StagePlan &s, FusedTcCache *cache, FusedTables *tabs, FusedTcTables *tct) {
    const TcKey key(s.vtab.get(), s.htab.get(), s.oy0 | (s.n_rows << 16), s.ox0 | (s.n_cols << 16), s.c | (s.c_out << 8) | (s.min_bands << 16));
    auto it = cache->geoms.find(key);
    if (it != cache->geoms.end()) return it->second;
    TcGeom g;
    // horizontal: the row is cut into chunks of 128 bytes starting at the 16-byte aligned b0 (the TMA
    // coordinate).  A pixel that straddles a chunk boundary waits: its leading bytes are carried to
    // the front of the next chunk's tile.  Per chunk {first pair, pairs | odd << 16, carried columns,
    // tile column of its first pixel}; per pair the weights (true scale: the tile holds true-scale f32).
    bool ok = true;
    const uint32_t C = s.c, xa = s.sx0, xe = s.sx0 + s.n_sx;
    g.b0 = (xa * C) & ~15u;
    g.n_chunks = (xe * C - g.b0 + TC_M - 1) / TC_M;
    // Horizontal stage on the tensor cores where the output ring and the shared-memory budget allow it: the band
    // keeps its rows as f16 hi / lo tiles (32 rows per group, <= 8 groups = two M = 128 tiles) next to >= 2 source slots.
    // FANLIN_PREFER_RING=1 (experiments): the ring variant also where the register-ring kernel fits
    static const bool prefer_ring = [] { const char *e = std::getenv("FANLIN_PREFER_RING"); return e && e[0] == '1'; }();
    if (cache->allow_hmma && prefer_ring && try_ring(s, g, tabs, tct)) return cache->geoms.emplace(key, std::move(g)).first->second;
    if (cache->allow_hmma) {
        const int sh = weight_shift(*s.vtab);
        uint32_t hm_out_stride = 1;
        {  // most outputs a chunk finishes -> words per row of the output staging buffer (odd: rows on distinct banks)
            uint32_t fin = s.ox0, widest = 0;
            for (uint32_t ch = 0; ch < g.n_chunks; ch++) {
                const uint32_t b1 = g.b0 + TC_M * (ch + 1), first = fin;
                while (fin < s.ox0 + s.n_cols && (s.htab->entries[fin].left + s.htab->entries[fin].count) * C <= b1) fin++;
                widest = std::max(widest, fin - first);
            }
            widest = std::max(widest, s.ox0 + s.n_cols - fin + 1);
            hm_out_stride = ((widest * s.c_out + 6) / 4) | 1u;
        }
        bool fits = false;
        uint32_t n_bands = std::max(1u, s.min_bands), band_rows = s.n_rows;
        for (; n_bands <= std::max(8u, s.min_bands) && !fits; n_bands++) {
            band_rows = (s.n_rows + n_bands - 1) / n_bands;
            fits = true;
            for (uint32_t r0 = 0; r0 < s.n_rows && fits; r0 += band_rows) {
                uint32_t ng = 0, kgm = 0;
                fits = plan_band(*s.vtab, s.oy0, r0, std::min(band_rows, s.n_rows - r0), 8, &ng, &kgm) && ng <= 8 &&
                       fused_tc2_smem_bytes(C, ng, kgm, 2, 1, std::min(band_rows, s.n_rows - r0), hm_out_stride) <= TC_SMEM_LIMIT;
            }
            if (fits) break;
        }
        if (fits && build_hmma(s, g, tabs, tct)) {
            g.scale = std::ldexp(1.0f, -sh);
            g.out_stride = hm_out_stride;
            bool okb = true;
            for (uint32_t r0 = 0; r0 < s.n_rows && okb; r0 += band_rows) {
                TcBand bt{};
                okb = build_band(*s.vtab, s.oy0, r0, std::min(band_rows, s.n_rows - r0), sh, 8, tabs, tct, &bt);
                g.bands.push_back(bt);
            }
            if (okb) {
                g.hm = true;
                g.ok = true;
                return cache->geoms.emplace(key, std::move(g)).first->second;
            }
            g.bands.clear();
        }
        if (try_ring(s, g, tabs, tct)) return cache->geoms.emplace(key, std::move(g)).first->second;
    }
    std::vector<PxPair> pairs;
    std::vector<uint32_t> recs;
    {
        uint32_t x = xa, carry = 0;
        for (uint32_t ch = 0; ch < g.n_chunks; ch++) {
            const uint32_t tile_byte0 = g.b0 + TC_M * ch - carry, end_byte = g.b0 + TC_M * (ch + 1);
            uint32_t npx = 0;
            while (x + npx < xe && (x + npx + 1) * C <= end_byte) npx++;
            recs.push_back(uint32_t(pairs.size()));
            recs.push_back(((npx + 1) / 2) | (npx & 1) << 16);
            recs.push_back(carry);
            recs.push_back(x * C - tile_byte0);
            for (uint32_t q = 0; q < npx; q += 2) pairs.push_back(PxPair{x + q, q + 1 < npx});
            x += npx;
            carry = x < xe ? end_byte - x * C : 0;
            if (carry >= C) ok = false;  // cannot happen: the pixel would have been whole
        }
        if (x != xe) ok = false;
    }
    tabs->w.resize((tabs->w.size() + 3) & ~size_t(3), 0.0f);  // the kernel stages the table with 16-byte copies
    g.hw_off = uint32_t(tabs->w.size());
    g.hinfo_off = uint32_t(tabs->info.size());
    const uint32_t n_pairs = uint32_t(pairs.size());
    tabs->w.resize(tabs->w.size() + size_t(n_pairs) * 16, 0.0f);
    tabs->info.resize(tabs->info.size() + n_pairs, 0u);
    ok = ok && tc_pair_table(*s.htab, s.ox0, s.n_cols, pairs, &tabs->w[g.hw_off], &tabs->info[g.hinfo_off]);
    g.chunk_off = uint32_t(tabs->info.size());
    tabs->info.insert(tabs->info.end(), recs.begin(), recs.end());
    // outputs finished before each chunk; the widest chunk sizes the output staging buffer
    {
        g.cpre_off = uint32_t(tabs->info.size());
        tabs->info.resize(tabs->info.size() + g.n_chunks + 1, 0u);
        uint32_t acc = 0, widest = 0;
        for (uint32_t ch = 0; ch < g.n_chunks; ch++) {
            tabs->info[g.cpre_off + ch] = acc;
            uint32_t here = 0;
            const uint32_t p0 = recs[4 * ch], np = recs[4 * ch + 1] & 0xffffu;
            for (uint32_t p = p0; p < p0 + np; p++) here += tabs->info[g.hinfo_off + p];
            acc += here;
            widest = std::max(widest, here);
        }
        tabs->info[g.cpre_off + g.n_chunks] = acc;
        g.out_stride = (widest * s.c_out + 6) / 4;  // ceil((3 bytes of alignment phase + pixels) / 4) words
    }
    // vertical: q = round(w * 2^sh) split into three signed base-128 digits
    const int sh = weight_shift(*s.vtab);
    g.scale = std::ldexp(1.0f, -sh);
    const uint32_t max_band = fused_tc_max_band(s.c, g.out_stride);
    if (max_band < 32) ok = false;  // strong upscales finish too many pixels per chunk to stage: CUDA-core path
    const uint32_t n_bands = (s.n_rows + max_band - 1) / max_band;
    const uint32_t band_rows = (s.n_rows + n_bands - 1) / n_bands;
    for (uint32_t b0 = 0; b0 < s.n_rows && ok; b0 += band_rows) {
        TcBand bt{};
        ok = build_band(*s.vtab, s.oy0, b0, std::min(band_rows, s.n_rows - b0), sh, 8, tabs, tct, &bt);
        g.bands.push_back(bt);
    }
    // An odd staging stride keeps the rows of a warp on distinct banks (a multiple of 16 words is a 4-way conflict
    // on every staged pixel); taken unless the extra word per row costs a band a source slot (C2 is that tight).
    if (ok && !(g.out_stride & 1)) {
        bool same = true;
        for (const TcBand &bt : g.bands)
            same = same && fused_tc_source_slots(s.c, bt.rows, bt.kg_max, g.out_stride + 1) == fused_tc_source_slots(s.c, bt.rows, bt.kg_max, g.out_stride);
        if (same) g.out_stride++;
    }
    g.ok = ok;
    return cache->geoms.emplace(key, std::move(g)).first->second;
}

bool fused_tc_geometry_ok(const StagePlan &s, FusedTcCache *cache, FusedTables *tabs, FusedTcTables *tct) {
    return geom_of(s, cache, tabs, tct).ok;
}

bool fused_tc_uses_hmma(const StagePlan &s, FusedTcCache *cache, FusedTables *tabs, FusedTcTables *tct) {
    const TcGeom &g = geom_of(s, cache, tabs, tct);
    return g.ok && g.hm && !g.ring;
}
bool fused_tc_uses_ring(const StagePlan &s, FusedTcCache *cache, FusedTables *tabs, FusedTcTables *tct) {
    const TcGeom &g = geom_of(s, cache, tabs, tct);
    return g.ok && g.ring;
}

int fused_tc_build(const StagePlan &s, const fanlin_job &job, const uint8_t *src, uint32_t src_pitch, uint8_t *dst,
                   FusedTcCache *cache, FusedTables *tabs, FusedTcTables *tct, std::vector<FusedTcItem> *items) {
    (void)job;
    const TcGeom &g = geom_of(s, cache, tabs, tct);
    if (!g.ok) return FANLIN_EINVAL;
    for (size_t b = 0; b < g.bands.size(); b++) {
        const TcBand &bt = g.bands[b];
        FusedTcItem f{};
        f.src = src; f.dst = dst; f.src_pitch = src_pitch; f.src_h = s.in_h;
        f.c = s.c;
        f.b0 = g.b0; f.n_chunks = g.n_chunks; f.chunk_off = g.chunk_off; f.max_pairs = fused_tc_max_pairs(s.c);
        f.band_r0 = bt.r0; f.band_rows = bt.rows; f.r_pad = r_pad_for(bt.rows, s.c);
        f.grp_off = bt.grp_off; f.n_groups = bt.n_groups; f.kg_max = bt.kg_max; f.grp_rows = bt.grp_rows;
        f.scale = g.scale;
        f.hw_off = g.hw_off; f.hinfo_off = g.hinfo_off; f.cpre_off = g.cpre_off; f.out_stride = g.out_stride;
        f.n_a = fused_tc_source_slots(s.c, bt.rows, bt.kg_max, g.out_stride);
        f.n_cols = s.n_cols;
        f.dst_pitch = s.canvas_pitch ? s.canvas_pitch : s.canvas_w * s.c_out; f.c_out = s.c_out; f.canvas_w = s.canvas_w; f.canvas_h = s.canvas_h;
        f.dst_x = s.dst_x; f.dst_y = s.dst_y; f.epi = s.epi; f.fill = s.fill;
        f.first_band = b == 0; f.last_band = b + 1 == g.bands.size();
        f.inv_off = s.color_op == COLOR_INVERT ? bt.inv_off : 0xffffffffu;
        if (g.ring) {
            f.hmma = 2; f.hrec_off = g.hrec_off;
            f.ring_cols = g.ring_cols; f.stage_stride = g.stage_stride; f.wh_bytes = g.wh_bytes;
            const uint32_t n_mt = (bt.n_groups + 3) / 4;
            f.n_vr = std::min(4u, (512u - n_mt * g.ring_cols) / TC_N);
            f.n_a = std::min(4u, f.n_vr); f.n_wh = 1;  // n_a <= n_vr: a slot's release is read off the barrier of the region its group used, one phase back at most
            while (f.n_a > 2 && fused_tc3_smem_bytes(bt.n_groups, bt.kg_max, f.n_a, 1, g.wh_bytes, g.stage_stride) > TC_SMEM_LIMIT) f.n_a--;
            if (fused_tc3_smem_bytes(bt.n_groups, bt.kg_max, f.n_a, 2, g.wh_bytes, g.stage_stride) <= TC_SMEM_LIMIT) f.n_wh = 2;
        } else if (g.hm) {
            f.hmma = 1; f.hrec_off = g.hrec_off;
            f.n_a = 4; f.n_wh = 1;
            while (f.n_a > 2 && fused_tc2_smem_bytes(s.c, bt.n_groups, bt.kg_max, f.n_a, 1, bt.rows, g.out_stride) > TC_SMEM_LIMIT) f.n_a--;
            if (fused_tc2_smem_bytes(s.c, bt.n_groups, bt.kg_max, f.n_a, 2, bt.rows, g.out_stride) <= TC_SMEM_LIMIT) f.n_wh = 2;
        }
        items->push_back(f);
    }
    return FANLIN_OK;
}

// ---- vertical blur on the tensor cores ------------------------------------------------------

size_t blur_v_tc_smem_bytes(uint32_t kg_max, uint32_t n_a) {
    return size_t(n_a) * kg_max * TC_M + 2 * size_t(TC_N) * kg_max + 1024;
}

bool blur_v_tc_eligible(const StagePlan &s, uint32_t pitch, const uint8_t *src) {
    if (!s.present || !s.separable || s.v_kind != KIND_GAUSSIAN || !s.vtab) return false;
    if (s.color_op != COLOR_NONE || s.c_mem != s.c) return false;
    if (s.n_rows != s.in_h || s.n_cols != s.in_w || s.oy0 != 0) return false;
    if (pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(src) & 15)) return false;  // TMA
    return 32 + s.vtab->max_taps <= TC_KG_MAX;
}

namespace {
struct BlurVGeom {
    bool ok = false;
    float scale = 1.f;
    std::vector<TcBand> bands;
};
}  // namespace

int blur_v_tc_build(const StagePlan &s, const uint8_t *src, uint32_t src_pitch, float *dst, FusedTcCache *cache, FusedTables *tabs,
                    FusedTcTables *tct, std::vector<BlurVTcItem> *items) {
    // geometry per (table, rows): cached next to the resample geometries under a key no resample uses (htab = null)
    const TcKey key(s.vtab.get(), nullptr, s.n_rows, 0u, 0u);
    auto it = cache->geoms.find(key);
    if (it == cache->geoms.end()) {
        TcGeom g;
        const int sh = weight_shift(*s.vtab);
        g.scale = std::ldexp(1.0f, -sh);
        bool ok = true;
        // bands of <= 8 groups: a band re-reads 2 * radius source rows (u8, against 4 bytes written per element),
        // and small batches of large images still fill the SMs
        const uint32_t max_band = 8 * TC_GROUP_ROWS;
        const uint32_t n_bands = (s.n_rows + max_band - 1) / max_band;
        const uint32_t band_rows = ((s.n_rows + n_bands - 1) / n_bands + TC_GROUP_ROWS - 1) / TC_GROUP_ROWS * TC_GROUP_ROWS;
        for (uint32_t b0 = 0; b0 < s.n_rows && ok; b0 += band_rows) {
            TcBand bt{};
            ok = build_band(*s.vtab, 0, b0, std::min(band_rows, s.n_rows - b0), sh, TC_GROUP_ROWS, tabs, tct, &bt);  // groups of exactly 32 rows
            g.bands.push_back(bt);
        }
        g.ok = ok;
        it = cache->geoms.emplace(key, std::move(g)).first;
    }
    const TcGeom &g = it->second;
    if (!g.ok) return FANLIN_EINVAL;
    const uint32_t n_e = s.in_w * s.c;
    for (const TcBand &bt : g.bands) {
        BlurVTcItem f{};
        f.src = src; f.dst = dst; f.src_pitch = src_pitch; f.src_h = s.in_h;
        f.n_e = n_e; f.n_chunks = (n_e + TC_M - 1) / TC_M;
        f.band_r0 = bt.r0; f.band_rows = bt.rows;
        f.grp_off = bt.grp_off; f.n_groups = bt.n_groups; f.kg_max = bt.kg_max;
        f.n_a = 4;
        while (f.n_a > 2 && blur_v_tc_smem_bytes(f.kg_max, f.n_a) > TC_SMEM_LIMIT) f.n_a--;
        f.scale = g.scale;
        items->push_back(f);
    }
    return FANLIN_OK;
}

}  // namespace fanlin
