// Context, buffer pools, prepared (device-resident) batches and the blocking
// host-buffer entry point.  No CPU fallback: every path ends in a CUDA launch or
// an error status.
#include "runtime.h"

#include "fused.h"
#include "blur.h"
#include "fused_tc.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>

using namespace fanlin;

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            set_error(std::string("fanlin: CUDA error: ") + cudaGetErrorString(e_) + " in " #expr); \
            return FANLIN_ECUDA;                                                               \
        }                                                                                      \
    } while (0)

namespace fanlin {

// ---- pinned pool --------------------------------------------------------------

static size_t size_class(size_t b) {
    size_t c = 64 << 10;
    while (c < b) c <<= 1;
    return c;
}

PinnedPool::~PinnedPool() {
    for (auto &kv : free_)
        for (void *p : kv.second) cudaFreeHost(p);
}

void *PinnedPool::alloc(size_t bytes) {
    const size_t cls = size_class(bytes);
    {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = free_.find(cls);
        if (it != free_.end() && !it->second.empty()) {
            void *p = it->second.back();
            it->second.pop_back();
            cached_ -= cls;
            live_[p] = cls;
            return p;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, cls, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(mu_);
    live_[p] = cls;
    return p;
}

void PinnedPool::free(void *p) {
    if (!p) return;
    size_t cls = 0;
    {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = live_.find(p);
        if (it == live_.end()) return;
        cls = it->second;
        live_.erase(it);
        if (cached_ + cls <= limit_) {
            free_[cls].push_back(p);
            cached_ += cls;
            return;
        }
    }
    cudaFreeHost(p);
}

bool PinnedPool::owns(const void *p, size_t bytes) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.upper_bound(p);
    if (it == live_.begin()) return false;
    --it;
    const char *base = static_cast<const char *>(it->first);
    const char *q = static_cast<const char *>(p);
    return q >= base && q + bytes <= base + it->second;
}

TableBuf::~TableBuf() {
    if (!d_w && !d_info && !d_b) return;
    cudaSetDevice(ordinal);
    cudaFree(d_w);
    cudaFree(d_info);
    cudaFree(d_b);
}

TableGen::TableGen(bool allow_hmma) : fcache(fused_cache_new()), tcache(fused_tc_cache_new(allow_hmma)), btcache(blur_tc_cache_new()) {}
TableGen::~TableGen() {
    fused_cache_free(fcache);
    fused_tc_cache_free(tcache);
    blur_tc_cache_free(btcache);
    if (uploaded) cudaEventDestroy(uploaded);
}

}  // namespace fanlin

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Brings the device copy of a generation's arenas up to date on `st` (under dev->tab_mu): grows the buffers when needed,
// copies what was appended since the last call and orders `st` behind every earlier upload, whatever stream made it.
static int upload_tables(fanlin_ctx *ctx, DeviceState *dev, TableGen *g, cudaStream_t st) {
    const size_t nw = g->ftabs.w.size() * 4, ni = g->ftabs.info.size() * 4, nb = g->tctabs.b.size();
    if (!g->uploaded && cudaEventCreateWithFlags(&g->uploaded, cudaEventDisableTiming) != cudaSuccess) { set_error("fanlin: cudaEventCreate failed"); return FANLIN_ECUDA; }
    if (!g->buf || nw > g->buf->cap_w || ni > g->buf->cap_info || nb > g->buf->cap_b) {
        auto grow = [](size_t need, size_t floor_) { size_t c = floor_; while (c < need) c <<= 1; return c; };
        std::shared_ptr<TableBuf> nbuf(new TableBuf());
        nbuf->ordinal = dev->ordinal;
        nbuf->cap_w = grow(nw, size_t(4) << 20); nbuf->cap_info = grow(ni, size_t(4) << 20); nbuf->cap_b = grow(nb, size_t(32) << 20);
        if (cudaMalloc(&nbuf->d_w, nbuf->cap_w) != cudaSuccess || cudaMalloc(&nbuf->d_info, nbuf->cap_info) != cudaSuccess ||
            cudaMalloc(&nbuf->d_b, nbuf->cap_b) != cudaSuccess) {
            cudaGetLastError();
            set_error("fanlin: device allocation for the filter tables failed");
            return FANLIN_ENOMEM;
        }
        g->buf = nbuf;  // the old buffers live on in the batches that launch from them
        g->up_w = g->up_info = g->up_b = 0;
    } else if (cudaStreamWaitEvent(st, g->uploaded, 0) != cudaSuccess) {
        set_error("fanlin: cudaStreamWaitEvent failed");
        return FANLIN_ECUDA;
    }
    cudaError_t e = cudaSuccess;
    if (g->ftabs.w.size() > g->up_w)
        e = cudaMemcpyAsync(static_cast<float *>(g->buf->d_w) + g->up_w, g->ftabs.w.data() + g->up_w, (g->ftabs.w.size() - g->up_w) * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && g->ftabs.info.size() > g->up_info)
        e = cudaMemcpyAsync(static_cast<uint32_t *>(g->buf->d_info) + g->up_info, g->ftabs.info.data() + g->up_info, (g->ftabs.info.size() - g->up_info) * 4,
                            cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && g->tctabs.b.size() > g->up_b)
        e = cudaMemcpyAsync(static_cast<uint8_t *>(g->buf->d_b) + g->up_b, g->tctabs.b.data() + g->up_b, g->tctabs.b.size() - g->up_b, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { set_error(std::string("fanlin: table upload failed: ") + cudaGetErrorString(e)); return FANLIN_ECUDA; }
    ctx->table_bytes += (g->ftabs.w.size() - g->up_w) * 4 + (g->ftabs.info.size() - g->up_info) * 4 + (g->tctabs.b.size() - g->up_b);
    g->up_w = g->ftabs.w.size(); g->up_info = g->ftabs.info.size(); g->up_b = g->tctabs.b.size();
    if (cudaEventRecord(g->uploaded, st) != cudaSuccess) { set_error("fanlin: cudaEventRecord failed"); return FANLIN_ECUDA; }
    return FANLIN_OK;
}

// ---- context --------------------------------------------------------------------

extern "C" int fanlin_abi_version(void) { return FANLIN_ABI_VERSION; }

extern "C" const char *fanlin_last_error(void) { return get_error(); }

extern "C" int fanlin_plan_job(const fanlin_job *job, fanlin_plan *plan) {
    if (!job || !plan) { set_error("fanlin: null argument"); return FANLIN_EINVAL; }
    JobPlan p;
    const int rc = plan_job(*job, &p, false);
    if (rc == FANLIN_OK) *plan = p.pub;
    return rc;
}

static void batcher_main(fanlin_ctx *ctx, int dev_index);

extern "C" int fanlin_init(const int *device_ids, int n_devices, const fanlin_config *cfg, fanlin_ctx **out) {
    if (!out) { set_error("fanlin: null out pointer"); return FANLIN_EINVAL; }
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        set_error("fanlin: no CUDA device available (this path has no CPU fallback)");
        return FANLIN_ENODEVICE;
    }
    std::vector<int> ids;
    if (n_devices <= 0 || !device_ids) {
        for (int i = 0; i < count; i++) ids.push_back(i);
    } else {
        for (int i = 0; i < n_devices; i++) {
            if (device_ids[i] < 0 || device_ids[i] >= count) { set_error("fanlin: bad device ordinal"); return FANLIN_EINVAL; }
            ids.push_back(device_ids[i]);
        }
    }
    std::unique_ptr<fanlin_ctx> ctx(new fanlin_ctx());
    if (cfg) std::memcpy(&ctx->cfg, cfg, std::min<size_t>(cfg->struct_size ? cfg->struct_size : sizeof(*cfg), sizeof(*cfg)));
    if (!ctx->cfg.device_scratch_bytes) ctx->cfg.device_scratch_bytes = uint64_t(8) << 30;
    if (!ctx->cfg.pinned_bytes) ctx->cfg.pinned_bytes = uint64_t(2) << 30;
    ctx->adaptive_window = ctx->cfg.batch_window_us == 0;  // default: no idle wait in the request batcher (batcher_main)
    if (!ctx->cfg.batch_window_us) ctx->cfg.batch_window_us = 200;
    if (!ctx->cfg.max_batch_jobs) ctx->cfg.max_batch_jobs = 4096;
    ctx->pinned.set_limit(ctx->cfg.pinned_bytes);
    for (int id : ids) {
        CUDA_TRY(cudaSetDevice(id));
        std::unique_ptr<DeviceState> d(new DeviceState());
        d->ordinal = id;
        CUDA_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&d->copy_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&d->copy_out, cudaStreamNonBlocking));
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, id) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;  // keep freed blocks cached: the device side of the buffer pool
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        ctx->devs.push_back(std::move(d));
    }
    fanlin_ctx *raw = ctx.release();
    for (size_t i = 0; i < raw->devs.size(); i++) raw->devs[i]->worker = std::thread(batcher_main, raw, int(i));
    *out = raw;
    return FANLIN_OK;
}

extern "C" void fanlin_shutdown(fanlin_ctx *ctx) {
    if (!ctx) return;
    ctx->down = true;
    for (auto &d : ctx->devs) {
        {
            std::lock_guard<std::mutex> lk(d->qmu);
            d->stop = true;
        }
        d->qcv.notify_all();
        if (d->worker.joinable()) d->worker.join();
        cudaSetDevice(d->ordinal);
        cudaStreamSynchronize(d->stream);
        d->gen.reset();
        cudaStreamDestroy(d->stream);
        cudaStreamDestroy(d->copy_in);
        cudaStreamDestroy(d->copy_out);
    }
    delete ctx;
}

extern "C" int fanlin_device_count(const fanlin_ctx *ctx) { return ctx ? int(ctx->devs.size()) : 0; }

extern "C" int fanlin_get_stats(const fanlin_ctx *ctx, fanlin_stats *out) {
    if (!ctx || !out) return FANLIN_EINVAL;
    out->kernel_launches = ctx->kernel_launches.load();
    out->jobs = ctx->jobs.load();
    out->batches = ctx->batches.load();
    out->h2d_bytes = ctx->h2d_bytes.load();
    out->d2h_bytes = ctx->d2h_bytes.load();
    out->table_bytes = ctx->table_bytes.load();
    return FANLIN_OK;
}

extern "C" void *fanlin_host_alloc(fanlin_ctx *ctx, size_t bytes) {
    if (!ctx || !bytes) return nullptr;
    return ctx->pinned.alloc(bytes);
}
extern "C" void fanlin_host_free(fanlin_ctx *ctx, void *p) {
    if (ctx) ctx->pinned.free(p);
}

// ---- prepared batches -------------------------------------------------------------

namespace {

struct JobScratch {
    size_t pre = 0, inter = 0, tmp = 0, fin = 0, lor = 0;                     // bytes
    size_t pre_off = 0, inter_off = 0, tmp_off = 0, fin_off = 0, lor_off = 0;  // offsets inside the chunk's scratch
};

void fill_desc(StageDesc *d, const StagePlan &s, const fanlin_job &job, uint8_t *inter, float *tmp,
               const std::map<const AxisTable *, uint32_t> &tab_base, bool last_stage) {
    std::memset(d, 0, sizeof(*d));
    if (s.src_is_input) {
        d->src = job.src;
        d->src_pitch = job.src_pitch ? job.src_pitch : job.src_w * job.src_channels * sample_bytes(job.src_sample);
    } else {
        d->src = inter;
        d->src_pitch = s.in_pitch ? s.in_pitch : s.in_w * s.c_mem * sample_bytes(s.s_in);
    }
    d->s_in = s.s_in; d->s_out = s.s_out;
    d->dst = last_stage ? job.dst : inter;
    d->tmp = tmp;
    d->src_w = s.in_w; d->src_h = s.in_h;
    d->c_mem = s.c_mem; d->c = s.c; d->color_op = s.color_op;
    d->v_tab = s.vtab ? tab_base.at(s.vtab.get()) : NO_TABLE;
    d->h_tab = s.htab ? tab_base.at(s.htab.get()) : NO_TABLE;
    d->v_max_taps = s.vtab ? s.vtab->max_taps : 0;
    d->h_max_taps = s.htab ? s.htab->max_taps : 0;
    d->oy0 = s.oy0; d->n_rows = s.n_rows; d->ox0 = s.ox0; d->n_cols = s.n_cols;
    d->sx0 = s.sx0; d->n_sx = s.n_sx; d->sy0 = s.sy0; d->n_sy = s.n_sy;
    d->tmp_pitch = s.n_sx * s.c;
    d->dst_pitch = s.canvas_pitch ? s.canvas_pitch : s.canvas_w * s.c_out * sample_bytes(s.s_out); d->c_out = s.c_out;
    d->canvas_w = s.canvas_w; d->canvas_h = s.canvas_h;
    d->dst_x = s.dst_x; d->dst_y = s.dst_y; d->epi = s.epi; d->fill = s.fill;
}

void geom_add(LaunchGeom *g, const StageDesc &d) {
    g->n_jobs++;
    g->max_n_rows = std::max(g->max_n_rows, d.n_rows);
    g->max_n_sx = std::max(g->max_n_sx, d.n_sx);
    g->max_n_cols = std::max(g->max_n_cols, d.n_cols);
    g->max_canvas_w = std::max(g->max_canvas_w, d.canvas_w);
    g->max_canvas_h = std::max(g->max_canvas_h, d.canvas_h);
}

}  // namespace

extern "C" void fanlin_batch_free(fanlin_batch *b) {
    if (!b) return;
    if (b->dev) cudaSetDevice(b->dev->ordinal);
    if (b->d_meta) cudaFreeAsync(b->d_meta, b->alloc_stream);
    if (b->d_scratch) cudaFreeAsync(b->d_scratch, b->alloc_stream);
    if (b->h_meta && b->ctx) b->ctx->pinned.free(b->h_meta);
    for (cudaEvent_t e : b->events) cudaEventDestroy(e);
    delete b;
}

static int prepare_on_stream(fanlin_ctx *ctx, int device_index, const fanlin_job *jobs, uint32_t n_jobs, fanlin_plan *plans_out,
                             fanlin_batch **out, cudaStream_t up_stream);

extern "C" int fanlin_batch_prepare(fanlin_ctx *ctx, int device_index, const fanlin_job *jobs, uint32_t n_jobs,
                                    fanlin_plan *plans_out, fanlin_batch **out) {
    if (!ctx || !out || (!jobs && n_jobs)) { set_error("fanlin: null argument"); return FANLIN_EINVAL; }
    if (device_index < 0 || device_index >= int(ctx->devs.size())) { set_error("fanlin: bad device index"); return FANLIN_EINVAL; }
    DeviceState *dev = ctx->devs[device_index].get();
    const int rc = prepare_on_stream(ctx, device_index, jobs, n_jobs, plans_out, out, dev->stream);
    if (rc != FANLIN_OK) return rc;
    // the caller may launch on any stream: make the uploaded descriptors visible first
    if (cudaStreamSynchronize(dev->stream) != cudaSuccess) { set_error("fanlin: CUDA error while uploading batch descriptors"); return FANLIN_ECUDA; }
    return FANLIN_OK;
}

// Plans the jobs, builds tables and descriptors and enqueues their upload on `up_stream`
// (launches on that stream are ordered after it; no host synchronisation).
static int prepare_on_stream(fanlin_ctx *ctx, int device_index, const fanlin_job *jobs, uint32_t n_jobs, fanlin_plan *plans_out,
                             fanlin_batch **out, cudaStream_t up_stream) {
    if (!ctx || !out || (!jobs && n_jobs)) { set_error("fanlin: null argument"); return FANLIN_EINVAL; }
    *out = nullptr;
    if (ctx->down) { set_error("fanlin: context is shut down"); return FANLIN_ESHUTDOWN; }
    if (device_index < 0 || device_index >= int(ctx->devs.size())) { set_error("fanlin: bad device index"); return FANLIN_EINVAL; }
    DeviceState *dev = ctx->devs[device_index].get();
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    std::unique_ptr<fanlin_batch, void (*)(fanlin_batch *)> b(new fanlin_batch(), fanlin_batch_free);
    b->ctx = ctx;
    b->dev = dev;
    b->n_jobs = n_jobs;
    b->plans.resize(n_jobs);
    const bool exact = ctx->cfg.exact != 0;
    // Geometry caches and table arenas live in the device's table generation (TableGen): what an earlier batch built is
    // found again, what this batch adds is uploaded behind the loop.  The generation is replaced -- never edited -- when
    // it has grown past a bound; batches in flight keep their TableBuf.
    std::unique_lock<std::mutex> tab_lock(dev->tab_mu);
    const bool allow_hmma = ctx->cfg.vertical_path == 3 || ctx->cfg.vertical_path == 0;  // both Lanczos3 passes on the tensor cores where the geometry allows
    // (FANLIN_TABLE_GEN_LIMIT_MB: the bound, for tests of the roll-over)
    static const size_t gen_limit = [] { const char *e = std::getenv("FANLIN_TABLE_GEN_LIMIT_MB"); const long v = e ? std::atol(e) : 0; return size_t(v > 0 ? v : 768) << 20; }();
    if (!dev->gen || dev->gen->host_bytes() > gen_limit) dev->gen = std::make_shared<TableGen>(allow_hmma);
    std::shared_ptr<TableGen> gen = dev->gen;
    FusedCache *const fcache_p = gen->fcache;
    FusedTables &ftabs = gen->ftabs;
    std::vector<uint8_t> fused_a(n_jobs, 0);  // stage A: 1 = fused resample kernel, 2 = its tensor-core variant
    std::vector<uint8_t> gather_a(n_jobs, 0); // stage A is a Nearest resample: the compose kernel gathers through the tap tables
    std::vector<StagePlan> a_pre(n_jobs);     // present: stage A as the tensor-core kernel sees it behind a colour-op pass
    // EXIF orientation AFTER the resample (plan.h stored_axes_stage): stage A runs on the image as stored and writes a small
    // plain image to scratch; that image is oriented (the pass that used to run over the full-resolution source) and, where
    // the oriented stage had a letterbox / to_rgba8 epilogue, composed onto the canvas.  late_a = the oriented stage A.
    std::vector<uint8_t> late(n_jobs, 0);
    std::vector<StagePlan> late_a(n_jobs);
    static const bool no_gray_canvas = [] { const char *e = std::getenv("FANLIN_GRAY_CANVAS"); return e && e[0] == '0'; }();  // (A/B switch)
    static const bool no_rgb8_epilogue = [] { const char *e = std::getenv("FANLIN_RGB8_EPILOGUE"); return e && e[0] == '0'; }();  // (A/B switch: 0 keeps the to_rgb8 pass)
    static const bool late_orient_on = [] { const char *e = std::getenv("FANLIN_LATE_ORIENT"); return !(e && e[0] == '0'); }();
    FusedTcCache *const tcache_p = gen->tcache;
    FusedTcTables &tctabs = gen->tctabs;
    const bool use_tc = ctx->cfg.vertical_path != 1;  // 2: tensor-core vertical pass, CUDA-core horizontal stage
    // 16-bit / f32 subpixels (fanlin_job.src_sample): the stages that read them take the generic kernels of kernels_deep.cu
    // (crate operation order); a blur behind a letterbox reads the Rgba<u8> canvas and is an ordinary u8 stage
    std::vector<uint8_t> deep_a(n_jobs, 0), deep_b(n_jobs, 0);
    std::vector<uint8_t> fast_b(n_jobs, 0);  // stage B takes the fast blur kernels
    std::vector<uint8_t> tc_b(n_jobs, 0);    // ... both passes on the tensor cores, no f32 intermediate (kernels_blur_tc.cu)
    BlurTables &btabs = gen->btabs;
    std::vector<BlurItem> bitems;
    std::vector<BlurTcItem> btitems;
    BlurTcCache *const btcache_p = gen->btcache;

    // 1. plans + table arena
    std::map<const AxisTable *, uint32_t> tab_base;
    std::vector<TapEntry> entries;
    std::vector<float> weights;
    auto add_table = [&](const std::shared_ptr<const AxisTable> &t) {
        if (!t || tab_base.count(t.get())) return;
        tab_base[t.get()] = uint32_t(entries.size());
        const uint32_t wb = uint32_t(weights.size());
        for (const TapEntry &e : t->entries) entries.push_back(TapEntry{e.left, e.count, e.woff + wb});
        weights.insert(weights.end(), t->weights.begin(), t->weights.end());
    };
    std::vector<fanlin_job> ej(jobs, jobs + n_jobs);  // the request as stages a / b see it: behind the orientation pass, if any
    for (uint32_t i = 0; i < n_jobs; i++) {
        const int rc = plan_job(jobs[i], &b->plans[i], true);
        if (rc != FANLIN_OK) return rc;
        if (b->plans[i].pre.present) ej[i] = b->plans[i].pre.job;  // src: its scratch image, set once the scratch is laid out
        // Gray canvas (EPI_GRAY, common.h): a one-channel image letterboxed onto a gray fill colour with a blur behind it -- the
        // default fill is (32, 32, 32), so every grayscale + fit + blur request -- is blurred as ONE plane and expanded to
        // (l, l, l, 255) by the last pass.  Tried first, taken back where stage A falls to a kernel without the epilogue.
        const bool gray_try = [&] {
            const JobPlan &p = b->plans[i];
            if (no_gray_canvas || exact || !use_tc || jobs[i].orientation >= 2 || jobs[i].src_sample != SAMPLE_U8) return false;
            if (!p.a.present || !p.b.present || !p.a.separable || p.a.epi != EPI_BLEND_FILL || p.a.c != 1 || p.a.v_kind != KIND_LANCZOS3) return false;
            if (p.b.epi != EPI_PLAIN || p.b.c != 4 || p.b.color_op != COLOR_NONE) return false;
            const uint32_t f = p.a.fill;
            return (f & 0xffu) == ((f >> 8) & 0xffu) && (f & 0xffu) == ((f >> 16) & 0xffu);
        }();
        const uint32_t post_c_in_keep = b->plans[i].post_c_in;
        auto set_gray = [&](bool on) {
            JobPlan &p = b->plans[i];
            p.a.c_out = on ? 1u : 4u;
            p.a.epi = on ? uint32_t(EPI_BLEND_FILL) | EPI_GRAY : uint32_t(EPI_BLEND_FILL);
            p.b.c_mem = p.b.c = p.b.c_out = on ? 1u : 4u;
            p.post_c_in = on ? 1u : post_c_in_keep;                      // FANLIN_TO_RGB8 / _YCBCR: the pass they have anyway reads the plane
            p.post_c_out = on && !post_c_in_keep ? 4u : 3u;              // else it writes (l, l, l, 255)
            if (p.a.present && p.b.present) {  // the canvas between the stages: rows on a 16-byte stride (TMA reads it)
                p.a.canvas_pitch = uint32_t(align_up(size_t(p.a.canvas_w) * p.a.c_out * sample_bytes(p.a.s_out), 16));
                p.b.in_pitch = p.a.canvas_pitch;
            }
        };
        if (gray_try) set_gray(true);
        else if (b->plans[i].a.present && b->plans[i].b.present) {  // the canvas between the stages: rows on a 16-byte stride (TMA reads it)
            StagePlan &sa = b->plans[i].a;
            sa.canvas_pitch = uint32_t(align_up(size_t(sa.canvas_w) * sa.c_out * sample_bytes(sa.s_out), 16));
            b->plans[i].b.in_pitch = sa.canvas_pitch;
        }
        if (!jobs[i].src || !jobs[i].dst) { set_error("fanlin: null src or dst"); return FANLIN_EINVAL; }
        if (jobs[i].dst_capacity < b->plans[i].pub.out_bytes) {
            set_error("fanlin: dst_capacity smaller than the planned output");
            return FANLIN_ECAPACITY;
        }
        if (plans_out) plans_out[i] = b->plans[i].pub;
        if (n_jobs < 74 && b->plans[i].a.present && b->plans[i].a.separable && b->plans[i].a.v_kind == KIND_LANCZOS3) {
            // a handful of requests: cut each image into bands, so that its resample runs on 148 / n_jobs SMs instead of one
            // (a single C2 request spent 0.19 of its 0.86 ms in a one-CTA kernel); bands of >= 24 rows, so that a band is about a group
            const uint32_t by_rows = std::max(1u, b->plans[i].a.n_rows / 24u);
            b->plans[i].a.min_bands = std::max(1u, std::min({148u / n_jobs, by_rows, 48u}));
        }
        deep_a[i] = b->plans[i].a.present && b->plans[i].a.s_in != SAMPLE_U8;
        deep_b[i] = b->plans[i].b.present && b->plans[i].b.s_in != SAMPLE_U8;
        if (late_orient_on && !exact && use_tc && jobs[i].src_sample == SAMPLE_U8 && b->plans[i].pre.present && b->plans[i].a.present && b->plans[i].a.separable &&
            b->plans[i].a.v_kind == KIND_LANCZOS3 && b->plans[i].a.n_rows && b->plans[i].a.n_cols) {
            const StagePlan in = stored_axes_stage(b->plans[i].a, jobs[i], jobs[i].orientation);
            StagePlan probe = in;  // as the tensor-core kernel would see it (behind the colour pass when grayscale is asked for)
            if (probe.color_op == COLOR_GRAY) { probe.color_op = COLOR_NONE; probe.c_mem = probe.c; probe.src_is_input = false; probe.in_pitch = uint32_t(align_up(size_t(probe.in_w) * probe.c, 16)); }
            // only where the stage on the stored image keeps the tensor-core kernels (rows on a 16-byte stride: the oriented
            // scratch image of the orientation pass always has them, a caller's device batch may not)
            if (fused_tc_eligible(probe, jobs[i])) {
                late[i] = 1;
                late_a[i] = b->plans[i].a;
                b->plans[i].a = in;
                b->plans[i].pre.present = false;  // no orientation pass over the source
                ej[i] = jobs[i];
                ej[i].orientation = 0;
            }
        }
        for (const auto &t : {b->plans[i].a.vtab, b->plans[i].a.htab, b->plans[i].b.vtab, b->plans[i].b.htab})
            if (t) gen->keep.insert(t);
        // FANLIN_TO_RGB8 as an epilogue: where stage A writes the final image and its kernel can leave
        // the alpha byte behind, the to_rgb8 pass over the output disappears -- tried first, taken back when the stage
        // then falls to a kernel that cannot (the CUDA-core fused kernel, the tensor-core kernel with the CUDA-core
        // horizontal stage).  With a blur behind it, or the orientation applied after the resample, the pass stays.
        // Request-sized batches only (< 74 jobs, where a launch less is what counts): three-byte pixels leave the tensor-core
        // kernels through their byte-staging / unaligned write-out paths, which on a large batch costs more than the pass
        // over the small outputs it saves (measured per image: C2 1.36 -> 1.56 us, C1 0.44 -> 0.75 us; the pass: ~0.1 us).
        const bool rgb8_try = !no_rgb8_epilogue && n_jobs < 74 && b->plans[i].post_c_in && !b->plans[i].post_ycbcr && b->plans[i].post_s_in == SAMPLE_U8 && b->plans[i].a.present &&
                              !b->plans[i].b.present && !late[i] && b->plans[i].a.s_out == SAMPLE_U8;
        const StagePlan a_keep = b->plans[i].a;
        if (rgb8_try) {
            StagePlan &ta = b->plans[i].a;
            ta.epi = (ta.epi == EPI_PLAIN ? uint32_t(EPI_TO_RGBA) : ta.epi) | EPI_RGB8;
            ta.c_out = 3;
        }
      for (int attempt = 0; attempt < 2; attempt++) {
        const JobPlan &p = b->plans[i];
        a_pre[i] = StagePlan();
        fused_a[i] = 0;
        gather_a[i] = !deep_a[i] && p.a.present && p.a.separable && p.a.v_kind == KIND_NEAREST && p.a.h_kind == KIND_NEAREST;  // one tap per output: a gather
        if (!exact && use_tc && !gather_a[i] && !deep_a[i]) {
            // (inverse rides on the vertical pass of the both-passes kernels only: the kernel with the CUDA-core horizontal
            // stage keeps the colour pass in front of it)
            if (fused_tc_eligible(p.a, ej[i]) && fused_tc_geometry_ok(p.a, tcache_p, &ftabs, &tctabs) &&
                (p.a.color_op == COLOR_NONE || fused_tc_uses_hmma(p.a, tcache_p, &ftabs, &tctabs) || fused_tc_uses_ring(p.a, tcache_p, &ftabs, &tctabs))) {
                fused_a[i] = 2;
            } else if (p.a.present && p.a.separable && p.a.color_op != COLOR_NONE && p.a.src_is_input) {
                // Grayscale / inverse keep the tensor-core path: the source bytes reach the tensor core
                // straight from memory, so the colour op runs first as a pass of its own into scratch
                // (1 + 2 c / c_mem times the source bytes instead of once, at several times the speed of
                // the CUDA-core resample that would apply it on load).
                StagePlan m = p.a;
                m.color_op = COLOR_NONE; m.c_mem = m.c; m.src_is_input = false;
                m.in_pitch = uint32_t(align_up(size_t(m.in_w) * m.c, 16));
                if (fused_tc_eligible(m, ej[i]) && fused_tc_geometry_ok(m, tcache_p, &ftabs, &tctabs)) {
                    fused_a[i] = 2;
                    a_pre[i] = m;
                }
            }
        }
        if (!fused_a[i] && !gather_a[i] && !deep_a[i]) fused_a[i] = !exact && fused_eligible(p.a, ej[i]) && fused_geometry_ok(p.a, fcache_p, &ftabs);
        if (attempt == 0 && gray_try) {
            const StagePlan &ta = a_pre[i].present ? a_pre[i] : p.a;
            const bool both_passes_tc = fused_a[i] == 2 && (fused_tc_uses_hmma(ta, tcache_p, &ftabs, &tctabs) || fused_tc_uses_ring(ta, tcache_p, &ftabs, &tctabs));
            if (fused_a[i] == 0 || both_passes_tc) break;  // generic / both-passes kernels: they know the one-byte canvas
            set_gray(false);
            continue;
        }
        if (attempt == 0 && rgb8_try) {
            const StagePlan &ta = a_pre[i].present ? a_pre[i] : p.a;
            const bool both_passes_tc = fused_a[i] == 2 && (fused_tc_uses_hmma(ta, tcache_p, &ftabs, &tctabs) || fused_tc_uses_ring(ta, tcache_p, &ftabs, &tctabs));
            if (fused_a[i] == 0 || both_passes_tc) { b->plans[i].post_c_in = 0; break; }  // compose / generic / deep / both-passes kernels: folded
            b->plans[i].a = a_keep;  // the kernel it fell to has no such epilogue: plan A as it was, the pass stays
            continue;
        }
        break;
      }
        const JobPlan &p = b->plans[i];
        if (!fused_a[i]) { add_table(p.a.vtab); add_table(p.a.htab); }
        fast_b[i] = !exact && !deep_b[i] && blur_eligible(p.b);
        if (fast_b[i] && use_tc && ctx->cfg.blur_path != 1) {
            // rows of the blur's input: the caller's image, or scratch (orientation pass / the canvas of stage A: 256-byte
            // aligned blocks, rows on a 16-byte stride)
            const bool from_caller = p.b.src_is_input && !p.pre.present;
            const uint8_t *bsrc = from_caller ? jobs[i].src : nullptr;
            const uint32_t bpitch = p.b.src_is_input ? (ej[i].src_pitch ? ej[i].src_pitch : ej[i].src_w * ej[i].src_channels)
                                                     : (p.b.in_pitch ? p.b.in_pitch : p.b.in_w * p.b.c);  // (u8 here: fast_b excludes deep_b)
            tc_b[i] = blur_tc_eligible(p.b, bpitch, bsrc);
        }
        if (!fast_b[i]) { add_table(p.b.vtab); add_table(p.b.htab); }
    }

    // 2. scratch layout, chunked so one chunk fits the scratch budget
    std::vector<JobScratch> js(n_jobs);
    std::vector<uint32_t> chunk_end;
    size_t scratch_bytes = 0;
    {
        size_t cur = 0;
        for (uint32_t i = 0; i < n_jobs; i++) {
            const JobPlan &p = b->plans[i];
            if (a_pre[i].present) js[i].pre = align_up(size_t(a_pre[i].in_pitch) * a_pre[i].in_h, 256);
            if (p.pre.present) js[i].pre = align_up(size_t(p.pre.job.src_pitch) * p.pre.job.src_h, 256);
            if (late[i]) {  // the stored-axes stage's plain output, and its oriented copy where a compose step follows
                js[i].lor = align_up(size_t(p.a.n_cols) * p.a.n_rows * p.a.c, 256);
                if (p.a.present && p.b.present) js[i].inter = align_up(size_t(late_a[i].canvas_pitch) * late_a[i].canvas_h, 256);
            }
            if (!late[i] && p.a.present && p.b.present) js[i].inter = align_up(size_t(p.a.canvas_pitch) * p.a.canvas_h, 256);
            size_t ta = 0, tb = 0;
            if (p.a.present && p.a.separable && !fused_a[i] && !gather_a[i]) ta = size_t(p.a.n_rows) * p.a.n_sx * p.a.c * 4;
            if (p.b.present && !tc_b[i]) tb = size_t(p.b.n_rows) * p.b.n_sx * p.b.c * 4;  // f32 intermediate of the two-kernel blur paths
            js[i].tmp = align_up(std::max(ta, tb), 256);
            if (p.post_c_in) js[i].fin = align_up(size_t(p.pub.out_w) * p.pub.out_h * p.post_c_in * sample_bytes(p.post_s_in), 256);  // the final image before to_rgb8
            const size_t need = js[i].pre + js[i].inter + js[i].tmp + js[i].fin + js[i].lor;
            if (cur && cur + need > ctx->cfg.device_scratch_bytes) {
                chunk_end.push_back(i);
                cur = 0;
            }
            js[i].pre_off = cur;
            js[i].inter_off = cur + js[i].pre;
            js[i].tmp_off = cur + js[i].pre + js[i].inter;
            js[i].fin_off = js[i].tmp_off + js[i].tmp;
            js[i].lor_off = js[i].fin_off + js[i].fin;
            cur += need;
            scratch_bytes = std::max(scratch_bytes, cur);
        }
        chunk_end.push_back(n_jobs);
    }
    b->alloc_stream = up_stream;
    if (scratch_bytes) CUDA_TRY(cudaMallocAsync(&b->d_scratch, scratch_bytes, up_stream));  // stream-ordered pool: no device sync
    for (uint32_t i = 0; i < n_jobs; i++)
        if (b->plans[i].pre.present) ej[i].src = static_cast<const uint8_t *>(b->d_scratch) + js[i].pre_off;
    for (uint32_t i = 0; i < n_jobs; i++)  // FANLIN_TO_RGB8: the stages write the final image to scratch, a last pass converts it into dst
        if (b->plans[i].post_c_in) ej[i].dst = static_cast<uint8_t *>(b->d_scratch) + js[i].fin_off;

    // where stage A of job i writes: its scratch image when the orientation follows, the canvas in front of the blur, the output
    auto a_dst = [&](uint32_t i) -> uint8_t * {
        if (late[i]) return static_cast<uint8_t *>(b->d_scratch) + js[i].lor_off;
        return b->plans[i].b.present ? static_cast<uint8_t *>(b->d_scratch) + js[i].inter_off : ej[i].dst;
    };

    // 3. descriptors per chunk and stage kind
    std::vector<StageDesc> descs;
    std::vector<FusedItem> fitems;
    std::vector<FusedTcItem> tcitems;
    std::vector<BlurVTcItem> bvitems;
    struct HostStep { int kind; size_t first; LaunchGeom g; uint32_t variant, n_items, max_band; size_t smem; uint32_t n_paired = 0; };
    std::vector<HostStep> hsteps;
    uint32_t begin = 0;
    for (uint32_t end : chunk_end) {
        // stage A through the fused kernel, one launch per (channels, colour op) variant
        std::map<uint32_t, std::vector<uint32_t>> by_variant;
        std::map<uint32_t, std::vector<uint32_t>> tc_by_c;
        for (uint32_t i = begin; i < end; i++) {
            if (fused_a[i] == 1) by_variant[fused_variant(b->plans[i].a)].push_back(i);
            if (fused_a[i] == 2) {  // key: channels | 8 when the horizontal stage runs on the tensor cores too (another kernel)
                const StagePlan &ta = a_pre[i].present ? a_pre[i] : b->plans[i].a;
                tc_by_c[ta.c | (fused_tc_uses_hmma(ta, tcache_p, &ftabs, &tctabs) ? 8u : 0u) |
                        (fused_tc_uses_ring(ta, tcache_p, &ftabs, &tctabs) ? 16u : 0u) | (ta.color_op == COLOR_INVERT ? 32u : 0u)].push_back(i);
            }
        }
        {  // orientation passes: the stored image turned (and its colour op applied) into scratch, in front of everything
          for (int deep = 0; deep < 2; deep++) {  // (the passes of 16-bit / f32 images: kind 12)
            HostStep hs{deep ? 12 : 6, descs.size(), LaunchGeom{}, 0, 0, 0, 0};
            for (uint32_t i = begin; i < end; i++) {
                const JobPlan &p = b->plans[i];
                if (!p.pre.present || (p.pre.sample != SAMPLE_U8) != (deep != 0)) continue;
                StageDesc d;
                std::memset(&d, 0, sizeof(d));
                d.src = jobs[i].src;
                d.src_pitch = jobs[i].src_pitch ? jobs[i].src_pitch : jobs[i].src_w * jobs[i].src_channels * sample_bytes(jobs[i].src_sample);
                d.src_w = jobs[i].src_w; d.src_h = jobs[i].src_h;
                d.s_in = d.s_out = p.pre.sample;
                d.c_mem = p.pre.c_mem; d.c = p.pre.c; d.color_op = p.pre.color_op; d.orient = p.pre.orient;
                d.v_tab = d.h_tab = NO_TABLE;
                d.oy0 = p.pub.src_y0; d.n_rows = p.pub.src_y1 - p.pub.src_y0; d.ox0 = 0; d.n_cols = ej[i].src_w;
                d.dst = const_cast<uint8_t *>(ej[i].src) + size_t(d.oy0) * ej[i].src_pitch;
                d.dst_pitch = ej[i].src_pitch; d.c_out = p.pre.c; d.canvas_w = ej[i].src_w; d.canvas_h = d.n_rows; d.epi = EPI_PLAIN;
                geom_add(&hs.g, d);
                descs.push_back(d);
            }
            if (hs.g.n_jobs) hsteps.push_back(hs);
          }
        }
        {  // colour-op passes in front of the tensor-core resample: the needed source rows, op applied, into scratch
            HostStep hs{5, descs.size(), LaunchGeom{}, 0, 0, 0, 0};
            for (uint32_t i = begin; i < end; i++) {
                if (!a_pre[i].present) continue;
                const StagePlan &a = b->plans[i].a;
                StageDesc d;
                std::memset(&d, 0, sizeof(d));
                d.src = jobs[i].src;
                d.src_pitch = jobs[i].src_pitch ? jobs[i].src_pitch : jobs[i].src_w * jobs[i].src_channels;
                d.src_w = a.in_w; d.src_h = a.in_h; d.c_mem = a.c_mem; d.c = a.c; d.color_op = a.color_op;
                d.v_tab = d.h_tab = NO_TABLE;
                d.oy0 = a.sy0; d.n_rows = a.n_sy; d.ox0 = 0; d.n_cols = a.in_w;
                d.dst = static_cast<uint8_t *>(b->d_scratch) + js[i].pre_off + size_t(a.sy0) * a_pre[i].in_pitch;
                d.dst_pitch = a_pre[i].in_pitch; d.c_out = a.c; d.canvas_w = a.in_w; d.canvas_h = a.n_sy; d.epi = EPI_PLAIN;
                geom_add(&hs.g, d);
                descs.push_back(d);
            }
            if (hs.g.n_jobs) hsteps.push_back(hs);
        }
        for (auto &kv : tc_by_c) {
            HostStep hs{3, tcitems.size(), LaunchGeom{}, kv.first, 0, 0, 0};
            for (uint32_t i : kv.second) {
                const JobPlan &p = b->plans[i];
                uint8_t *inter = js[i].inter ? static_cast<uint8_t *>(b->d_scratch) + js[i].inter_off : nullptr;
                const bool pre = a_pre[i].present;
                const uint8_t *tsrc = pre ? static_cast<const uint8_t *>(b->d_scratch) + js[i].pre_off : ej[i].src;
                const uint32_t pitch = pre ? a_pre[i].in_pitch : ej[i].src_pitch ? ej[i].src_pitch : ej[i].src_w * ej[i].src_channels;
                const int rc = fused_tc_build(pre ? a_pre[i] : p.a, ej[i], tsrc, pitch, a_dst(i), tcache_p, &ftabs,
                                              &tctabs, &tcitems);
                if (rc != FANLIN_OK) { set_error("fanlin: internal: tensor-core tables"); return rc; }
            }
            hs.n_items = uint32_t(tcitems.size() - hs.first);
            for (size_t k = hs.first; k < tcitems.size(); k++)
                hs.smem = std::max(hs.smem, fused_tc_item_smem(tcitems[k]));
            if (hs.n_items) hsteps.push_back(hs);
        }
        for (auto &kv : by_variant) {
            HostStep hs{2, fitems.size(), LaunchGeom{}, kv.first, 0, 0, 0};
            for (uint32_t i : kv.second) {
                const JobPlan &p = b->plans[i];
                uint8_t *inter = js[i].inter ? static_cast<uint8_t *>(b->d_scratch) + js[i].inter_off : nullptr;
                const uint32_t pitch = ej[i].src_pitch ? ej[i].src_pitch : ej[i].src_w * ej[i].src_channels;
                const int rc = fused_build(p.a, ej[i].src, pitch, a_dst(i), fcache_p, &ftabs, &fitems);
                if (rc != FANLIN_OK) { set_error("fanlin: internal: fused tables"); return rc; }
            }
            hs.n_items = uint32_t(fitems.size() - hs.first);
            for (size_t k = hs.first; k < fitems.size(); k++) hs.max_band = std::max(hs.max_band, fitems[k].band_rows);
            if (hs.n_items) hsteps.push_back(hs);
        }
        for (int pass = 0; pass < 3; pass++) {  // 0: A separable (generic), 1: A compose, 2: B separable
            if (pass == 2) {
                // orientation after the resample: the stored-axes stage's small image turned into the oriented one -- by the
                // orientation pass straight onto the canvas where the oriented stage's epilogue is plain, else by a compose step
                // that reads its source through the orientation (letterbox / to_rgba8 in the same kernel)
                HostStep ho{6, descs.size(), LaunchGeom{}, 0, 0, 0, 0};
                for (uint32_t i = begin; i < end; i++) {
                    if (!late[i] || late_a[i].epi != EPI_PLAIN) continue;
                    const JobPlan &p = b->plans[i];
                    const StagePlan &oa = late_a[i];
                    uint8_t *lo = static_cast<uint8_t *>(b->d_scratch) + js[i].lor_off;
                    uint8_t *final_dst = p.b.present ? static_cast<uint8_t *>(b->d_scratch) + js[i].inter_off : ej[i].dst;
                    const uint32_t cpitch = oa.canvas_pitch ? oa.canvas_pitch : oa.canvas_w * oa.c_out;
                    StageDesc d;
                    std::memset(&d, 0, sizeof(d));
                    d.src = lo; d.src_pitch = p.a.n_cols * p.a.c; d.src_w = p.a.n_cols; d.src_h = p.a.n_rows;
                    d.c_mem = d.c = d.c_out = p.a.c; d.color_op = COLOR_NONE; d.orient = jobs[i].orientation;
                    d.v_tab = d.h_tab = NO_TABLE;
                    d.oy0 = 0; d.n_rows = oa.n_rows; d.ox0 = 0; d.n_cols = oa.n_cols; d.canvas_w = oa.n_cols; d.canvas_h = oa.n_rows; d.epi = EPI_PLAIN;
                    d.dst = final_dst + size_t(oa.dst_y) * cpitch + size_t(oa.dst_x) * oa.c_out;
                    d.dst_pitch = cpitch;
                    geom_add(&ho.g, d);
                    descs.push_back(d);
                }
                if (ho.g.n_jobs) hsteps.push_back(ho);
                HostStep hc{1, descs.size(), LaunchGeom{}, 0, 0, 0, 0};
                for (uint32_t i = begin; i < end; i++) {
                    if (!late[i] || late_a[i].epi == EPI_PLAIN) continue;
                    const JobPlan &p = b->plans[i];
                    StagePlan cs = late_a[i];  // the oriented stage's canvas, placement, epilogue and fill; its input is the oriented small image
                    cs.separable = false; cs.src_is_input = false; cs.vtab = nullptr; cs.htab = nullptr; cs.color_op = COLOR_NONE;
                    cs.in_w = p.a.n_cols; cs.in_h = p.a.n_rows; cs.c_mem = cs.c = p.a.c; cs.in_pitch = p.a.n_cols * p.a.c;  // the small image as stored
                    cs.ox0 = cs.oy0 = 0; cs.sx0 = cs.sy0 = 0; cs.n_sx = cs.n_cols; cs.n_sy = cs.n_rows;
                    uint8_t *lo = static_cast<uint8_t *>(b->d_scratch) + js[i].lor_off;
                    StageDesc d;
                    fill_desc(&d, cs, ej[i], lo, nullptr, tab_base, true);
                    d.orient = jobs[i].orientation;
                    d.dst = p.b.present ? static_cast<uint8_t *>(b->d_scratch) + js[i].inter_off : ej[i].dst;
                    geom_add(&hc.g, d);
                    descs.push_back(d);
                }
                if (hc.g.n_jobs) hsteps.push_back(hc);
            }
          for (int deep = 0; deep < 2; deep++) {  // (stages on 16-bit / f32 subpixels: kinds 10 / 11)
            HostStep hs{(pass == 1 ? 1 : 0) + (deep ? 10 : 0), descs.size(), LaunchGeom{}, 0, 0, 0, 0};
            for (uint32_t i = begin; i < end; i++) {
                const JobPlan &p = b->plans[i];
                const StagePlan &s = pass == 2 ? p.b : p.a;
                if (!s.present) continue;
                if ((pass == 2 ? deep_b[i] : deep_a[i]) != deep) continue;
                if (pass == 0 && (!s.separable || fused_a[i] || gather_a[i])) continue;
                if (pass == 1 && s.separable && !gather_a[i]) continue;
                if (pass == 2 && fast_b[i]) continue;
                uint8_t *inter = js[i].inter ? static_cast<uint8_t *>(b->d_scratch) + js[i].inter_off : nullptr;
                float *tmp = js[i].tmp ? reinterpret_cast<float *>(static_cast<uint8_t *>(b->d_scratch) + js[i].tmp_off) : nullptr;
                StageDesc d;
                const bool last = pass == 2 || !p.b.present;
                fill_desc(&d, s, ej[i], inter, tmp, tab_base, last);
                if (pass < 2 && late[i]) d.dst = a_dst(i);
                geom_add(&hs.g, d);
                descs.push_back(d);
            }
            if (hs.g.n_jobs) hsteps.push_back(hs);
          }
        }
        // stage B through the fast blur kernels, one launch pair per (channels, sigma); the vertical pass
        // on the tensor cores where the rows allow TMA (16-byte stride), else on the CUDA cores
        std::map<std::tuple<uint32_t, uint32_t, uint32_t>, std::vector<uint32_t>> blur_groups;
        auto b_src = [&](uint32_t i, const uint8_t **src, uint32_t *pitch) {
            const JobPlan &p = b->plans[i];
            uint8_t *inter = js[i].inter ? static_cast<uint8_t *>(b->d_scratch) + js[i].inter_off : nullptr;
            *src = p.b.src_is_input ? ej[i].src : inter;
            *pitch = p.b.src_is_input ? (ej[i].src_pitch ? ej[i].src_pitch : ej[i].src_w * ej[i].src_channels)
                                      : (p.b.in_pitch ? p.b.in_pitch : p.b.in_w * p.b.c);
        };
        {  // both blur passes on the tensor cores: one launch for all such jobs of the chunk
            HostStep hb{9, btitems.size(), LaunchGeom{}, 0, 0, 0, 0};
            for (uint32_t i = begin; i < end; i++) {
                if (!tc_b[i]) continue;
                const JobPlan &p = b->plans[i];
                const uint8_t *src; uint32_t pitch;
                b_src(i, &src, &pitch);
                const int rc = blur_tc_build(p.b, src, pitch, ej[i].dst, p.b.in_w * p.b.c, btcache_p, &btabs, &ftabs, &tctabs, &btitems);
                if (rc != FANLIN_OK) { set_error("fanlin: internal: tensor-core blur tables"); return rc; }
            }
            hb.n_items = uint32_t(btitems.size() - hb.first);
            for (size_t k = hb.first; k < btitems.size(); k++) hb.smem = std::max(hb.smem, blur_tc_smem_bytes(btitems[k].box_rows, btitems[k].kg_max, btitems[k].n_win));
            if (hb.n_items) hsteps.push_back(hb);
        }
        for (uint32_t i = begin; i < end; i++)
            if (fast_b[i] && !tc_b[i]) {
                uint32_t sb;
                std::memcpy(&sb, &b->plans[i].b.sigma, 4);
                const uint8_t *src; uint32_t pitch;
                b_src(i, &src, &pitch);
                const uint32_t vtc = use_tc && blur_v_tc_eligible(b->plans[i].b, pitch, src) ? 1u : 0u;
                blur_groups[std::make_tuple(b->plans[i].b.c, sb, vtc)].push_back(i);
            }
        for (auto &kv : blur_groups) {
            const bool vtc = std::get<2>(kv.first) != 0;
            HostStep hv{7, bvitems.size(), LaunchGeom{}, 0, 0, 0, 0};
            HostStep hs{4, bitems.size(), LaunchGeom{}, 0, 0, 0, 0};
            hs.n_paired = vtc ? 1 : 0;  // kind 4: the vertical pass was done by the step before
            for (uint32_t i : kv.second) {
                const JobPlan &p = b->plans[i];
                BlurItem bi{};
                blur_build(p.b, &btabs, &ftabs.w, &bi);
                b_src(i, &bi.src, &bi.src_pitch);
                bi.dst = ej[i].dst;
                bi.tmp = reinterpret_cast<float *>(static_cast<uint8_t *>(b->d_scratch) + js[i].tmp_off);
                bi.aligned4 = (bi.src_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(bi.src) & 3) == 0);
                hs.g.max_canvas_w = std::max(hs.g.max_canvas_w, bi.w);
                hs.g.max_canvas_h = std::max(hs.g.max_canvas_h, bi.h);
                hs.variant = bi.c; hs.n_items++; hs.max_band = bi.radius; hs.smem = bi.taps_pad;
                bitems.push_back(bi);
                if (vtc) {
                    const int rc = blur_v_tc_build(p.b, bi.src, bi.src_pitch, bi.tmp, tcache_p, &ftabs, &tctabs, &bvitems);
                    if (rc != FANLIN_OK) { set_error("fanlin: internal: tensor-core blur tables"); return rc; }
                }
            }
            if (vtc) {
                hv.n_items = uint32_t(bvitems.size() - hv.first);
                for (size_t k = hv.first; k < bvitems.size(); k++) hv.smem = std::max(hv.smem, blur_v_tc_smem_bytes(bvitems[k].kg_max, bvitems[k].n_a));
                hsteps.push_back(hv);
            }
            hsteps.push_back(hs);
        }
        {  // FANLIN_TO_RGB8: the last pass, scratch -> dst
          for (int deep = 0; deep < 2; deep++) {
            HostStep hs{deep ? 13 : 8, descs.size(), LaunchGeom{}, 0, 0, 0, 0};
            for (uint32_t i = begin; i < end; i++) {
                const JobPlan &p = b->plans[i];
                if (!p.post_c_in || (p.post_s_in != SAMPLE_U8) != (deep != 0)) continue;
                StageDesc d;
                std::memset(&d, 0, sizeof(d));
                d.src = ej[i].dst; d.dst = jobs[i].dst;
                d.s_in = p.post_s_in; d.s_out = SAMPLE_U8;
                d.c_mem = p.post_c_in; d.c = 3; d.c_out = p.post_ycbcr ? 3u : p.post_c_out;
                d.epi = p.post_ycbcr ? 1u : 0u;  // (this pass only: 1 = planar Y, Cb, Cr instead of interleaved RGB)
                d.canvas_w = p.pub.out_w; d.canvas_h = p.pub.out_h;
                d.v_tab = d.h_tab = NO_TABLE;
                geom_add(&hs.g, d);
                descs.push_back(d);
            }
            if (hs.g.n_jobs) hsteps.push_back(hs);
          }
        }
        begin = end;
    }

    // 4. upload descriptors + tables in one block
    const size_t off_tab = align_up(descs.size() * sizeof(StageDesc), 256);
    const size_t off_w = off_tab + align_up(entries.size() * sizeof(TapEntry), 256);
    const size_t off_fi = off_w + align_up(weights.size() * sizeof(float), 256);
    const size_t off_ti = off_fi + align_up(fitems.size() * sizeof(FusedItem), 256);
    const size_t off_tm = off_ti + align_up(tcitems.size() * sizeof(FusedTcItem), 256);
    const size_t off_bi = off_tm + align_up(tcitems.size() * 128, 256);
    const size_t off_bvi = off_bi + align_up(bitems.size() * sizeof(BlurItem), 256);
    const size_t off_bvm = off_bvi + align_up(bvitems.size() * sizeof(BlurVTcItem), 256);
    const size_t off_bti = off_bvm + align_up(bvitems.size() * 128, 256);
    const size_t off_btm = off_bti + align_up(btitems.size() * sizeof(BlurTcItem), 256);
    const size_t meta_bytes = off_btm + align_up(btitems.size() * 128, 256) + 256;
    b->h_meta = ctx->pinned.alloc(meta_bytes);  // pinned, kept until the batch is freed: the upload is asynchronous
    if (!b->h_meta) { set_error("fanlin: pinned allocation failed"); return FANLIN_ENOMEM; }
    struct MetaView {
        uint8_t *p;
        uint8_t *data() { return p; }
    } meta{static_cast<uint8_t *>(b->h_meta)};
    std::memset(meta.data(), 0, meta_bytes);
    if (!tcitems.empty()) std::memcpy(meta.data() + off_ti, tcitems.data(), tcitems.size() * sizeof(FusedTcItem));
    if (!bitems.empty()) std::memcpy(meta.data() + off_bi, bitems.data(), bitems.size() * sizeof(BlurItem));
    if (!bvitems.empty()) std::memcpy(meta.data() + off_bvi, bvitems.data(), bvitems.size() * sizeof(BlurVTcItem));
    for (size_t k = 0; k < bvitems.size(); k++) {
        if (!encode_row_tile_map(meta.data() + off_bvm + k * 128, bvitems[k].src, bvitems[k].src_pitch, bvitems[k].src_h, bvitems[k].kg_max)) {
            set_error("fanlin: cuTensorMapEncodeTiled failed");
            return FANLIN_ECUDA;
        }
    }
    if (!btitems.empty()) std::memcpy(meta.data() + off_bti, btitems.data(), btitems.size() * sizeof(BlurTcItem));
    for (size_t k = 0; k < btitems.size(); k++) {  // the row bytes beyond n_e (pitch padding) must read as zeros: the map's width is n_e
        if (!encode_row_tile_map(meta.data() + off_btm + k * 128, btitems[k].src, btitems[k].src_pitch, btitems[k].src_h, btitems[k].box_rows, btitems[k].n_e)) {
            set_error("fanlin: cuTensorMapEncodeTiled failed");
            return FANLIN_ECUDA;
        }
    }
    for (size_t k = 0; k < tcitems.size(); k++) {  // one TMA tensor map per (image, band): box rows = the band's kg_max
        if (!encode_row_tile_map(meta.data() + off_tm + k * 128, tcitems[k].src, tcitems[k].src_pitch, tcitems[k].src_h, tcitems[k].kg_max)) {
            set_error("fanlin: cuTensorMapEncodeTiled failed");
            return FANLIN_ECUDA;
        }
    }
    if (!fitems.empty()) std::memcpy(meta.data() + off_fi, fitems.data(), fitems.size() * sizeof(FusedItem));
    if (!descs.empty()) std::memcpy(meta.data(), descs.data(), descs.size() * sizeof(StageDesc));
    if (!entries.empty()) std::memcpy(meta.data() + off_tab, entries.data(), entries.size() * sizeof(TapEntry));
    if (!weights.empty()) std::memcpy(meta.data() + off_w, weights.data(), weights.size() * sizeof(float));
    CUDA_TRY(cudaMallocAsync(&b->d_meta, meta_bytes, up_stream));
    CUDA_TRY(cudaMemcpyAsync(b->d_meta, meta.data(), meta_bytes, cudaMemcpyHostToDevice, up_stream));
    const uint8_t *mbase = static_cast<const uint8_t *>(b->d_meta);
    b->d_tab = reinterpret_cast<const TapEntry *>(mbase + off_tab);
    b->d_w = reinterpret_cast<const float *>(mbase + off_w);
    {  // the generation's tables: upload what this batch appended, then launch from its device copy
        const int urc = upload_tables(ctx, dev, gen.get(), up_stream);
        if (urc != FANLIN_OK) return urc;
        b->tbuf = gen->buf;
        b->d_fw = static_cast<const float *>(b->tbuf->d_w);
        b->d_finfo = static_cast<const uint32_t *>(b->tbuf->d_info);
        b->d_tb = static_cast<const uint8_t *>(b->tbuf->d_b);
        tab_lock.unlock();
    }
    for (const HostStep &hs : hsteps) {
        if (hs.kind == 4) {
            for (uint32_t o = 0; o < hs.n_items; o += 65535) {  // grid.z carries the job index
                fanlin_batch::Step st{};
                st.kind = 4;
                st.blur_items = reinterpret_cast<const BlurItem *>(mbase + off_bi) + hs.first + o;
                st.n_items = std::min<uint32_t>(65535, hs.n_items - o);
                st.max_w = hs.g.max_canvas_w; st.max_h = hs.g.max_canvas_h;
                st.c = hs.variant; st.radius = hs.max_band; st.taps_pad = uint32_t(hs.smem);
                st.n_paired = hs.n_paired;  // 1: vertical pass done on the tensor cores by the step before
                b->steps.push_back(st);
                b->launches_per_run += hs.n_paired ? 1 : 2;
            }
            continue;
        }
        if (hs.kind == 9) {
            fanlin_batch::Step st{};
            st.kind = 9;
            st.bt_items = reinterpret_cast<const BlurTcItem *>(mbase + off_bti) + hs.first;
            st.tmaps = mbase + off_btm + hs.first * 128;
            st.n_items = hs.n_items;
            st.smem = hs.smem;
            b->steps.push_back(st);
            b->launches_per_run += 1;
            continue;
        }
        if (hs.kind == 7) {
            fanlin_batch::Step st{};
            st.kind = 7;
            st.bv_items = reinterpret_cast<const BlurVTcItem *>(mbase + off_bvi) + hs.first;
            st.tmaps = mbase + off_bvm + hs.first * 128;
            st.n_items = hs.n_items;
            st.smem = hs.smem;
            b->steps.push_back(st);
            b->launches_per_run += 1;
            continue;
        }
        if (hs.kind == 3) {
            fanlin_batch::Step st{};
            st.kind = 3;
            st.tc_items = reinterpret_cast<const FusedTcItem *>(mbase + off_ti) + hs.first;
            st.tmaps = mbase + off_tm + hs.first * 128;
            st.n_items = hs.n_items;
            st.variant = hs.variant;
            st.smem = hs.smem;
            b->steps.push_back(st);
            b->launches_per_run += 1;
            continue;
        }
        if (hs.kind == 2) {
            fanlin_batch::Step st{};
            st.kind = 2;
            st.items = reinterpret_cast<const FusedItem *>(mbase + off_fi) + hs.first;
            st.n_items = hs.n_items;
            st.variant = hs.variant;
            st.max_band = hs.max_band;
            b->steps.push_back(st);
            b->launches_per_run += 1;
            continue;
        }
        // grid.y carries the job index: split launches above the 65535 limit
        for (uint32_t o = 0; o < hs.g.n_jobs; o += 65535) {
            fanlin_batch::Step st{};
            st.kind = hs.kind;
            st.descs = reinterpret_cast<const StageDesc *>(mbase) + hs.first + o;
            st.geom = hs.g;
            st.geom.n_jobs = std::min<uint32_t>(65535, hs.g.n_jobs - o);
            b->steps.push_back(st);
            b->launches_per_run += (st.kind == 0 || st.kind == 10) ? 2 : 1;
        }
    }
    *out = b.release();
    return FANLIN_OK;
}

extern "C" int fanlin_batch_launch(fanlin_batch *b, void *cuda_stream) {
    if (!b) { set_error("fanlin: null batch"); return FANLIN_EINVAL; }
    CUDA_TRY(cudaSetDevice(b->dev->ordinal));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : b->dev->stream;
    int n = 0;
    LaunchCtx lc;
    lc.st = st;
    if (b->timing) {  // events accumulate over launches until fanlin_batch_kernel_times reads them
        lc.events = &b->events;
        lc.names = &b->ev_names;
        lc.used = b->ev_used;
    }
    for (const fanlin_batch::Step &s : b->steps) {
        if (s.kind == 4) {
            n += launch_blur(s.blur_items, s.n_items, s.max_w, s.max_h, s.c, s.radius, s.taps_pad, b->d_fw, s.n_paired != 0, lc);
        } else if (s.kind == 9) {
            n += launch_blur_tc(s.bt_items, s.tmaps, s.n_items, s.smem, b->d_tb, b->d_finfo, b->d_fw, lc);
        } else if (s.kind == 7) {
            n += launch_blur_v_tc(s.bv_items, s.tmaps, s.n_items, s.smem, b->d_tb, b->d_finfo, lc);
        } else if (s.kind == 3) {
            const int k = launch_fused_tc(s.tc_items, s.tmaps, s.n_items, s.variant, s.smem, b->d_tb, b->d_fw, b->d_finfo, lc);
            if (k < 0) { set_error("fanlin: internal: no tensor-core kernel variant"); return FANLIN_EINVAL; }
            n += k;
        } else if (s.kind == 2) {
            const int k = launch_fused(s.items, s.n_items, s.variant, s.max_band, b->d_fw, b->d_finfo, lc);
            if (k < 0) { set_error("fanlin: internal: no fused kernel variant"); return FANLIN_EINVAL; }
            n += k;
        } else if (s.kind == 10) n += launch_sep_deep(s.descs, b->d_tab, b->d_w, s.geom, lc);
        else if (s.kind == 11) n += launch_compose_deep(s.descs, s.geom, lc);
        else if (s.kind == 12) n += launch_orient_deep(s.descs, s.geom, lc);
        else if (s.kind == 13) n += launch_to_rgb8_deep(s.descs, s.geom, lc);
        else if (s.kind == 8) n += launch_to_rgb8(s.descs, s.geom, lc);
        else if (s.kind == 6) n += launch_orient_pass(s.descs, s.geom, lc);
        else if (s.kind == 5) n += launch_color_pass(s.descs, s.geom, lc);
        else if (s.kind == 1) n += launch_compose(s.descs, b->d_tab, s.geom, lc);
        else n += launch_sep_exact(s.descs, b->d_tab, b->d_w, s.geom, lc);
    }
    if (b->timing) b->ev_used = lc.used;
    CUDA_TRY(cudaGetLastError());
    b->ctx->kernel_launches += uint64_t(n);
    b->ctx->jobs += b->n_jobs;
    b->ctx->batches += 1;
    return FANLIN_OK;
}

extern "C" int fanlin_batch_launch_count(const fanlin_batch *b) { return b ? b->launches_per_run : 0; }

extern "C" int fanlin_batch_set_timing(fanlin_batch *b, int enable) {
    if (!b) return FANLIN_EINVAL;
    b->timing = enable != 0;
    return FANLIN_OK;
}

extern "C" int fanlin_batch_kernel_times(fanlin_batch *b, const char **names, float *ms, int cap) {
    if (!b) return 0;
    const int n = int(b->ev_used / 2);
    for (int i = 0; i < n && i < cap; i++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, b->events[2 * i], b->events[2 * i + 1]) != cudaSuccess) { cudaGetLastError(); t = -1.f; }
        if (ms) ms[i] = t;
        if (names) names[i] = b->ev_names[i];
    }
    b->ev_used = 0;
    b->ev_names.clear();
    return n;
}

// ---- host-buffer entry point -------------------------------------------------------

namespace {

constexpr uint32_t BATCH_STAGING_MIN_JOBS = 32;  // (= BATCHER_DIRECT_JOBS: calls this large are batches)

// True when the copy engine can read / write p directly: memory from fanlin_host_alloc, or anything else the caller
// pinned (cudaHostAlloc / cudaHostRegister).  A decoder's Vec<u8> is pageable: cudaMemcpyAsync from it is staged by the
// driver, synchronously and at a fraction of the link rate, so large batches of pageable buffers go through the
// library's own pinned staging (run_on_device).
bool is_pinned(fanlin_ctx *ctx, const void *p, size_t bytes) {
    if (ctx->pinned.owns(p, bytes)) return true;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// Copies `n` blocks on up to `threads` host threads (a single thread moves ~10 GB/s, the link five times that).
struct CopyOp { void *dst; const void *src; size_t bytes; };
void parallel_copy(const std::vector<CopyOp> &ops, unsigned threads) {
    size_t total = 0;
    for (const CopyOp &o : ops) total += o.bytes;
    threads = unsigned(std::min<size_t>(threads, std::max<size_t>(1, total >> 22)));  // at least 4 MB per thread
    if (threads <= 1 || ops.size() < 2) {
        for (const CopyOp &o : ops) std::memcpy(o.dst, o.src, o.bytes);
        return;
    }
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (size_t i = next++; i < ops.size(); i = next++) std::memcpy(ops[i].dst, ops[i].src, ops[i].bytes);
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < threads; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}

// One sub-batch in flight on one stream: device staging buffers, pinned staging for pageable callers, the prepared batch.
struct Flight {
    uint8_t *d_in = nullptr, *d_out = nullptr;
    uint8_t *h_in = nullptr, *h_out = nullptr;  // pinned staging (from the context's pool) when the caller's buffers are pageable
    std::vector<CopyOp> out_ops;                // h_out -> caller dst, once the stream has drained
    fanlin_ctx *ctx = nullptr;
    fanlin_batch *batch = nullptr;
    cudaStream_t st = nullptr;
    void finish_outputs(unsigned threads) {  // call after the stream has been synchronised
        if (!out_ops.empty()) parallel_copy(out_ops, threads);
        out_ops.clear();
    }
    void release() {
        if (batch) fanlin_batch_free(batch);
        if (d_in) cudaFreeAsync(d_in, st);
        if (d_out) cudaFreeAsync(d_out, st);
        if (h_in) ctx->pinned.free(h_in);
        if (h_out) ctx->pinned.free(h_out);
        batch = nullptr;
        d_in = d_out = nullptr;
        h_in = h_out = nullptr;
    }
};

// Runs jobs (host pointers) on one device.  The batch is cut into sub-batches that alternate
// between two streams, so the H2D copies of one overlap the kernels and D2H copies of the other.
int run_on_device(fanlin_ctx *ctx, int dev_index, const fanlin_job *jobs, uint32_t n, fanlin_plan *plans) {
    DeviceState *dev = ctx->devs[dev_index].get();
    std::lock_guard<std::mutex> lk(dev->mu);
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    std::vector<fanlin_plan> pl(n);
    for (uint32_t i = 0; i < n; i++) {
        const int rc = fanlin_plan_job(&jobs[i], &pl[i]);
        if (rc != FANLIN_OK) return rc;
        if (!jobs[i].src || !jobs[i].dst) { set_error("fanlin: null src or dst"); return FANLIN_EINVAL; }
        if (jobs[i].dst_capacity < pl[i].out_bytes) { set_error("fanlin: dst_capacity smaller than the planned output"); return FANLIN_ECAPACITY; }
    }
    const uint32_t per = std::max<uint32_t>(64, (n + 15) / 16);  // <= 16 sub-batches, each at least 64 jobs (fewer where max_bytes cuts them)
    size_t max_bytes = size_t(2) << 30;
    Flight fl[2];
    fl[0].st = dev->copy_in;
    fl[1].st = dev->copy_out;
    fl[0].ctx = fl[1].ctx = ctx;
    // Pageable callers (the Vec<u8> of a decoder, src/handler.rs:219): batches stage through pinned buffers of the context's
    // pool -- the host copy of sub-batch k + 1 runs while sub-batch k is on the link.  Judged on the first job (a batch comes
    // from one producer); single requests keep the driver's own staging, which costs the same as ours for one image.
    // (FANLIN_COPY_THREADS: experiments)
    static const unsigned copy_threads_env = [] { const char *e = std::getenv("FANLIN_COPY_THREADS"); const long v = e ? std::atol(e) : 0; return unsigned(v > 0 ? v : 0); }();
    const unsigned copy_threads = copy_threads_env ? copy_threads_env
                                                   : std::max(1u, std::min(16u, std::thread::hardware_concurrency() / unsigned(std::max<size_t>(1, ctx->devs.size()))));
    const bool stage_in = n >= BATCH_STAGING_MIN_JOBS && !is_pinned(ctx, jobs[0].src, 1);
    const bool stage_out = n >= BATCH_STAGING_MIN_JOBS && !is_pinned(ctx, jobs[0].dst, 1);
    if (stage_in) max_bytes = size_t(256) << 20;  // staged sub-batches: small enough that two of them stay in the pinned pool's cache
    int rc = FANLIN_OK;
    uint32_t begin = 0, k = 0;
    std::vector<fanlin_job> djobs;
    std::vector<size_t> in_off, out_off;
    while (begin < n && rc == FANLIN_OK) {
        Flight &f = fl[k & 1];
        if (f.batch || f.d_in) {  // the sub-batch that used this slot two rounds ago must have drained
            if (cudaStreamSynchronize(f.st) != cudaSuccess) { set_error("fanlin: CUDA error in a sub-batch"); rc = FANLIN_ECUDA; break; }
            f.finish_outputs(copy_threads);
            f.release();
        }
        uint32_t end = begin;
        size_t in_bytes = 0, out_bytes = 0;
        in_off.clear();
        out_off.clear();
        while (end < n && end - begin < per) {
            // device rows start every 16 bytes whatever the width: the tensor-core path (TMA) needs that stride
            const size_t ib = align_up(align_up(size_t(jobs[end].src_w) * jobs[end].src_channels * sample_bytes(jobs[end].src_sample), 16) * jobs[end].src_h, 256);
            if (end > begin && in_bytes + ib > max_bytes) break;
            in_off.push_back(in_bytes);
            out_off.push_back(out_bytes);
            in_bytes += ib;
            out_bytes += align_up(pl[end].out_bytes, 256);
            end++;
        }
        const uint32_t m = end - begin;
        if (cudaMallocAsync(reinterpret_cast<void **>(&f.d_in), in_bytes + 256, f.st) != cudaSuccess ||
            cudaMallocAsync(reinterpret_cast<void **>(&f.d_out), out_bytes + 256, f.st) != cudaSuccess) {
            cudaGetLastError();
            set_error("fanlin: device allocation failed");
            rc = FANLIN_ENOMEM;
            break;
        }
        djobs.assign(jobs + begin, jobs + end);
        if (stage_in || stage_out) {
            if (stage_in) f.h_in = static_cast<uint8_t *>(ctx->pinned.alloc(in_bytes + 256));
            if (stage_out) f.h_out = static_cast<uint8_t *>(ctx->pinned.alloc(out_bytes + 256));
            if ((stage_in && !f.h_in) || (stage_out && !f.h_out)) { set_error("fanlin: pinned staging allocation failed"); rc = FANLIN_ENOMEM; break; }
        }
        // Only the source rows the output depends on cross the link (fanlin_plan.src_y0 .. src_y1: a crop=true request on a
        // 4000x3000 image reads 2484 of its 3000 rows): they land at their own place in the device image, the rows around
        // them stay unwritten and unread.  Stored-rotated images (EXIF >= 2) are copied whole.
        auto rows_of = [&](uint32_t idx, uint32_t *y0, uint32_t *y1) {
            const fanlin_job &j = jobs[idx];
            *y0 = 0; *y1 = j.src_h;
            if (j.orientation < 2 && pl[idx].src_y1 > pl[idx].src_y0 && pl[idx].src_y1 <= j.src_h) { *y0 = pl[idx].src_y0; *y1 = pl[idx].src_y1; }
        };
        if (stage_in) {  // caller rows -> pinned staging, in the layout of the device buffer (rows on a 16-byte stride)
            std::vector<CopyOp> ops;
            for (uint32_t i = 0; i < m; i++) {
                const fanlin_job &j = jobs[begin + i];
                const size_t row = size_t(j.src_w) * j.src_channels * sample_bytes(j.src_sample);
                const size_t pitch = j.src_pitch ? j.src_pitch : row, dpitch = align_up(row, 16);
                uint32_t y0, y1;
                rows_of(begin + i, &y0, &y1);
                if (pitch == row && dpitch == row) {
                    ops.push_back(CopyOp{f.h_in + in_off[i] + y0 * row, j.src + y0 * row, row * (y1 - y0)});
                } else {
                    for (uint32_t y = y0; y < y1; y++) ops.push_back(CopyOp{f.h_in + in_off[i] + y * dpitch, j.src + y * pitch, row});
                }
            }
            parallel_copy(ops, copy_threads);
            const cudaError_t e = cudaMemcpyAsync(f.d_in, f.h_in, in_bytes, cudaMemcpyHostToDevice, f.st);  // one copy per sub-batch
            if (e != cudaSuccess) { set_error(std::string("fanlin: H2D failed: ") + cudaGetErrorString(e)); rc = FANLIN_ECUDA; break; }
        }
        for (uint32_t i = 0; i < m && rc == FANLIN_OK; i++) {
            const fanlin_job &j = jobs[begin + i];
            const size_t row = size_t(j.src_w) * j.src_channels * sample_bytes(j.src_sample);
            const size_t pitch = j.src_pitch ? j.src_pitch : row, dpitch = align_up(row, 16);
            uint32_t y0, y1;
            rows_of(begin + i, &y0, &y1);
            // ... and of those rows only the columns it depends on (src_x0 .. src_x1), where that saves at least a twentieth of
            // the row: a 2-D copy to the same place in the device image (what lies around it is read with zero weights or not at all)
            size_t xb0 = 0, xb1 = row;
            {
                const fanlin_plan &pp = pl[begin + i];
                const size_t pxb = size_t(j.src_channels) * sample_bytes(j.src_sample);
                if (!stage_in && j.orientation < 2 && pp.src_x1 > pp.src_x0 && pp.src_x1 <= j.src_w && (size_t(pp.src_x1 - pp.src_x0) * pxb) * 20 <= row * 19) {
                    xb0 = size_t(pp.src_x0) * pxb;
                    xb1 = size_t(pp.src_x1) * pxb;
                }
            }
            const bool whole_rows = xb0 == 0 && xb1 == row;
            const cudaError_t e = stage_in ? cudaSuccess
                                  : whole_rows && pitch == row && dpitch == row
                                      ? cudaMemcpyAsync(f.d_in + in_off[i] + y0 * row, j.src + y0 * row, row * (y1 - y0), cudaMemcpyHostToDevice, f.st)
                                      : cudaMemcpy2DAsync(f.d_in + in_off[i] + y0 * dpitch + xb0, dpitch, j.src + y0 * pitch + xb0, pitch, xb1 - xb0, y1 - y0, cudaMemcpyHostToDevice, f.st);
            if (e != cudaSuccess) { set_error(std::string("fanlin: H2D failed: ") + cudaGetErrorString(e)); rc = FANLIN_ECUDA; }
            djobs[i].src = f.d_in + in_off[i];
            djobs[i].src_pitch = uint32_t(dpitch);
            djobs[i].dst = f.d_out + out_off[i];
            djobs[i].dst_capacity = pl[begin + i].out_bytes;
            ctx->h2d_bytes += (xb1 - xb0) * (y1 - y0);
        }
        if (rc != FANLIN_OK) break;
        rc = prepare_on_stream(ctx, dev_index, djobs.data(), m, nullptr, &f.batch, f.st);
        if (rc != FANLIN_OK) break;
        rc = fanlin_batch_launch(f.batch, f.st);
        if (rc != FANLIN_OK) break;
        if (stage_out) {  // one copy per sub-batch into pinned staging; handed to the caller's buffers when the stream has drained
            const cudaError_t e = cudaMemcpyAsync(f.h_out, f.d_out, out_bytes, cudaMemcpyDeviceToHost, f.st);
            if (e != cudaSuccess) { set_error(std::string("fanlin: D2H failed: ") + cudaGetErrorString(e)); rc = FANLIN_ECUDA; break; }
            for (uint32_t i = 0; i < m; i++) {
                f.out_ops.push_back(CopyOp{jobs[begin + i].dst, f.h_out + out_off[i], size_t(pl[begin + i].out_bytes)});
                ctx->d2h_bytes += pl[begin + i].out_bytes;
            }
        } else {
            for (uint32_t i = 0; i < m; i++) {
                const cudaError_t e = cudaMemcpyAsync(jobs[begin + i].dst, f.d_out + out_off[i], pl[begin + i].out_bytes, cudaMemcpyDeviceToHost, f.st);
                if (e != cudaSuccess) { set_error(std::string("fanlin: D2H failed: ") + cudaGetErrorString(e)); rc = FANLIN_ECUDA; break; }
                ctx->d2h_bytes += pl[begin + i].out_bytes;
            }
        }
        begin = end;
        k++;
    }
    for (Flight &f : fl) {
        const cudaError_t se = cudaStreamSynchronize(f.st);
        if (rc == FANLIN_OK && se != cudaSuccess) {
            set_error(std::string("fanlin: CUDA error: ") + cudaGetErrorString(se));
            rc = FANLIN_ECUDA;
        }
        if (rc == FANLIN_OK) f.finish_outputs(copy_threads);
        f.release();
    }
    if (rc == FANLIN_OK && plans) std::copy(pl.begin(), pl.end(), plans);
    return rc;
}

}  // namespace

extern "C" void fanlin_shard_range(uint32_t n_jobs, uint32_t n_shards, uint32_t shard, uint32_t *lo, uint32_t *hi) {
    if (n_shards == 0) n_shards = 1;
    const uint32_t q = n_jobs / n_shards, r = n_jobs % n_shards;  // the first r shards take one more
    const uint32_t a = shard < n_shards ? shard * q + std::min(shard, r) : n_jobs;
    const uint32_t b = shard < n_shards ? a + q + (shard < r ? 1 : 0) : n_jobs;
    if (lo) *lo = a;
    if (hi) *hi = b;
}

// ---- YCCK -> CMYK (decode side, src/handler.rs:420-439) ---------------------------------------

extern "C" int fanlin_ycck_to_cmyk_device(fanlin_ctx *ctx, int device_index, const uint8_t *src, uint8_t *dst, uint64_t n_pixels,
                                          void *cuda_stream) {
    if (!ctx || (n_pixels && (!src || !dst))) { set_error("fanlin: null argument"); return FANLIN_EINVAL; }
    if (ctx->down) { set_error("fanlin: context is shut down"); return FANLIN_ESHUTDOWN; }
    if (device_index < 0 || device_index >= int(ctx->devs.size())) { set_error("fanlin: bad device index"); return FANLIN_EINVAL; }
    DeviceState *dev = ctx->devs[device_index].get();
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    LaunchCtx lc;
    lc.st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : dev->stream;
    const int n = launch_ycck_to_cmyk(src, dst, size_t(n_pixels), lc);
    CUDA_TRY(cudaGetLastError());
    ctx->kernel_launches += uint64_t(n);
    return FANLIN_OK;
}

extern "C" int fanlin_ycck_to_cmyk(fanlin_ctx *ctx, const uint8_t *src, uint8_t *dst, uint64_t n_pixels) {
    if (!ctx || (n_pixels && (!src || !dst))) { set_error("fanlin: null argument"); return FANLIN_EINVAL; }
    if (ctx->down) { set_error("fanlin: context is shut down"); return FANLIN_ESHUTDOWN; }
    if (n_pixels == 0) return FANLIN_OK;
    const int dev_index = int(ctx->rr++ % uint32_t(ctx->devs.size()));
    DeviceState *dev = ctx->devs[dev_index].get();
    std::lock_guard<std::mutex> lk(dev->mu);
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    const uint64_t chunk_px = uint64_t(8) << 20;  // 32 MB per chunk, two in flight
    cudaStream_t sts[2] = {dev->copy_in, dev->copy_out};
    uint8_t *d_buf[2] = {nullptr, nullptr};
    int rc = FANLIN_OK;
    for (int k = 0; k < 2 && rc == FANLIN_OK; k++)
        if (cudaMallocAsync(reinterpret_cast<void **>(&d_buf[k]), size_t(std::min(chunk_px, n_pixels)) * 4, sts[k]) != cudaSuccess) {
            cudaGetLastError();
            set_error("fanlin: device allocation failed");
            rc = FANLIN_ENOMEM;
        }
    uint32_t k = 0;
    for (uint64_t p0 = 0; p0 < n_pixels && rc == FANLIN_OK; p0 += chunk_px, k ^= 1u) {
        const uint64_t np = std::min(chunk_px, n_pixels - p0);
        LaunchCtx lc;
        lc.st = sts[k];  // stream order: the chunk that used this buffer two rounds ago has left it
        if (cudaMemcpyAsync(d_buf[k], src + 4 * p0, size_t(np) * 4, cudaMemcpyHostToDevice, sts[k]) != cudaSuccess) { rc = FANLIN_ECUDA; break; }
        ctx->kernel_launches += uint64_t(launch_ycck_to_cmyk(d_buf[k], d_buf[k], size_t(np), lc));
        if (cudaMemcpyAsync(dst + 4 * p0, d_buf[k], size_t(np) * 4, cudaMemcpyDeviceToHost, sts[k]) != cudaSuccess) { rc = FANLIN_ECUDA; break; }
        ctx->h2d_bytes += np * 4;
        ctx->d2h_bytes += np * 4;
    }
    for (int q = 0; q < 2; q++) {
        const cudaError_t se = cudaStreamSynchronize(sts[q]);
        if (se != cudaSuccess && rc == FANLIN_OK) rc = FANLIN_ECUDA;
        if (d_buf[q]) cudaFreeAsync(d_buf[q], sts[q]);
    }
    if (rc == FANLIN_ECUDA) set_error(std::string("fanlin: CUDA error in ycck_to_cmyk: ") + cudaGetErrorString(cudaGetLastError()));
    return rc;
}

// Request batcher: one collector thread per device.  Small fanlin_run calls (a request = one
// image, src/main.rs:179 calls process_image from up to max_clients tokio workers at once) queue
// here; the collector waits batch_window_us after the first arrival, merges what came in into one
// ragged batch, runs it, and wakes the callers.  A failing merged batch is re-run request by
// request so that one bad image fails alone.
constexpr uint32_t BATCHER_DIRECT_JOBS = 32;  // calls this large are batches already

static void batcher_main(fanlin_ctx *ctx, int dev_index) {
    DeviceState *dev = ctx->devs[dev_index].get();
    for (;;) {
        std::vector<Request *> take;
        {
            std::unique_lock<std::mutex> lk(dev->qmu);
            dev->qcv.wait(lk, [&] { return dev->stop || !dev->queue.empty(); });
            if (dev->stop && dev->queue.empty()) return;
            const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(ctx->cfg.batch_window_us);
            uint32_t jobs = 0;
            for (;;) {
                while (!dev->queue.empty() && jobs < ctx->cfg.max_batch_jobs) {
                    jobs += dev->queue.front()->n;
                    take.push_back(dev->queue.front());
                    dev->queue.pop_front();
                }
                if (jobs >= ctx->cfg.max_batch_jobs || dev->stop) break;
                // default (batch_window_us left 0): no idle wait -- what is queued goes now, what arrives while this batch runs is
                // merged into the next one (queues form exactly when the device is busy).  A lone client used to see the whole
                // window, 200 us, on every request.  An explicit batch_window_us keeps the fixed window.
                if (ctx->adaptive_window) break;
                if (dev->qcv.wait_until(lk, deadline) == std::cv_status::timeout && dev->queue.empty()) break;
            }
        }
        std::vector<fanlin_job> merged;
        std::vector<fanlin_plan> plans;
        for (Request *r : take) merged.insert(merged.end(), r->jobs, r->jobs + r->n);
        plans.resize(merged.size());
        int rc = run_on_device(ctx, dev_index, merged.data(), uint32_t(merged.size()), plans.data());
        size_t off = 0;
        for (Request *r : take) {
            int rrc = rc;
            std::string err;
            if (rc != FANLIN_OK && take.size() > 1) {  // isolate the failure
                rrc = run_on_device(ctx, dev_index, r->jobs, r->n, r->plans);
                if (rrc != FANLIN_OK) err = get_error();
            } else if (rc != FANLIN_OK) {
                err = get_error();
            } else if (r->plans) {
                std::copy(plans.begin() + off, plans.begin() + off + r->n, r->plans);
            }
            off += r->n;
            std::lock_guard<std::mutex> lk(r->m);
            r->rc = rrc;
            r->err = err;
            r->done = true;
            r->cv.notify_one();
        }
    }
}

extern "C" int fanlin_run(fanlin_ctx *ctx, const fanlin_job *jobs, uint32_t n_jobs, fanlin_plan *plans) {
    if (!ctx || (!jobs && n_jobs)) { set_error("fanlin: null argument"); return FANLIN_EINVAL; }
    if (ctx->down) { set_error("fanlin: context is shut down"); return FANLIN_ESHUTDOWN; }
    if (n_jobs == 0) return FANLIN_OK;
    const int nd = int(ctx->devs.size());
    if (n_jobs < BATCHER_DIRECT_JOBS) {  // a request or a short GIF: through the batcher of one device
        DeviceState *dev = ctx->devs[ctx->rr++ % uint32_t(nd)].get();
        Request r;
        r.jobs = jobs;
        r.n = n_jobs;
        r.plans = plans;
        {
            std::lock_guard<std::mutex> lk(dev->qmu);
            if (dev->stop) { set_error("fanlin: context is shut down"); return FANLIN_ESHUTDOWN; }
            dev->queue.push_back(&r);
        }
        dev->qcv.notify_all();
        std::unique_lock<std::mutex> lk(r.m);
        r.cv.wait(lk, [&] { return r.done; });
        if (r.rc != FANLIN_OK) set_error(r.err);
        return r.rc;
    }
    if (nd == 1) return run_on_device(ctx, 0, jobs, n_jobs, plans);
    // shard by image index: contiguous blocks, one host thread per device, no collective
    std::vector<std::thread> th;
    std::vector<int> rcs(nd, FANLIN_OK);
    std::vector<std::string> errs(nd);
    for (int d = 0; d < nd; d++) {
        uint32_t lo, hi;
        fanlin_shard_range(n_jobs, uint32_t(nd), uint32_t(d), &lo, &hi);
        if (lo == hi) continue;
        th.emplace_back([&, d, lo, hi] {
            rcs[d] = run_on_device(ctx, d, jobs + lo, hi - lo, plans ? plans + lo : nullptr);
            if (rcs[d] != FANLIN_OK) errs[d] = get_error();
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < nd; d++)
        if (rcs[d] != FANLIN_OK) { set_error(errs[d]); return rcs[d]; }
    return FANLIN_OK;
}
