// Context, per-device state and prepared batches.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "blur.h"
#include "fused.h"
#include "fused_tc.h"
#include "kernels.h"
#include "plan.h"

namespace fanlin {

// Size-class pool of pinned host buffers (cudaHostAlloc is milliseconds; reuse).
class PinnedPool {
   public:
    ~PinnedPool();
    void *alloc(size_t bytes);
    void free(void *p);
    bool owns(const void *p, size_t bytes);
    void set_limit(size_t bytes) { limit_ = bytes; }

   private:
    std::mutex mu_;
    std::map<size_t, std::vector<void *>> free_;  // class size -> buffers
    std::map<const void *, size_t> live_;         // base -> class size
    size_t cached_ = 0, limit_ = size_t(2) << 30;
};

// One blocked fanlin_run call waiting in a device's batcher queue.
struct Request {
    const fanlin_job *jobs = nullptr;
    uint32_t n = 0;
    fanlin_plan *plans = nullptr;
    int rc = 0;
    std::string err;
    bool done = false;
    std::mutex m;
    std::condition_variable cv;
};

// Filter tables of a device that outlive a batch: the geometry caches of the fused / tensor-core / blur builders, the
// host arenas they append to (weights, info words, weight tiles) and their device copies.  A second batch of a geometry
// this context has seen uploads no table bytes (fanlin_stats.table_bytes), which is what lets a single request take the
// kernels whose tables cost ~1 MB per geometry to build.  Offsets handed to the kernels are relative to the arenas and stay
// valid for the life of the generation; when the device copy has to grow, a new TableBuf is allocated and the old one
// lives on until the last batch that launches from it is freed.
struct TableBuf {
    int ordinal = 0;
    void *d_w = nullptr, *d_info = nullptr, *d_b = nullptr;
    size_t cap_w = 0, cap_info = 0, cap_b = 0;  // bytes
    ~TableBuf();
};
struct FusedCache;
struct FusedTcCache;
struct TableGen {
    FusedCache *fcache = nullptr;
    FusedTcCache *tcache = nullptr;
    BlurTcCache *btcache = nullptr;
    FusedTables ftabs;
    FusedTcTables tctabs;
    BlurTables btabs;
    std::set<std::shared_ptr<const AxisTable>> keep;  // the caches are keyed by table address: held here, an address stays one table
    std::shared_ptr<TableBuf> buf;
    size_t up_w = 0, up_info = 0, up_b = 0;           // elements (floats, words, bytes) already on the device
    cudaEvent_t uploaded = nullptr;                   // recorded behind the latest upload
    explicit TableGen(bool allow_hmma);
    ~TableGen();
    size_t host_bytes() const { return ftabs.w.size() * 4 + ftabs.info.size() * 4 + tctabs.b.size(); }
};

struct DeviceState {
    int ordinal = 0;
    cudaStream_t stream = nullptr;   // context stream (prepare uploads, default launches)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    std::mutex mu;                   // serialises host-path batches on this device
    std::mutex tab_mu;               // guards gen (prepare may run on several threads)
    std::shared_ptr<TableGen> gen;
    // request batcher
    std::mutex qmu;
    std::condition_variable qcv;
    std::deque<Request *> queue;
    std::thread worker;
    bool stop = false;
};

}  // namespace fanlin

struct fanlin_ctx {
    std::vector<std::unique_ptr<fanlin::DeviceState>> devs;
    fanlin_config cfg{};
    fanlin::PinnedPool pinned;
    std::atomic<uint64_t> kernel_launches{0}, jobs{0}, batches{0}, h2d_bytes{0}, d2h_bytes{0}, table_bytes{0};
    std::atomic<uint32_t> rr{0};
    std::atomic<bool> down{false};
    bool adaptive_window = false;  // batch_window_us was left at its default: the request batcher does not wait when its queue is empty
};

struct fanlin_batch {
    fanlin_ctx *ctx = nullptr;
    fanlin::DeviceState *dev = nullptr;
    uint32_t n_jobs = 0;
    std::vector<fanlin::JobPlan> plans;
    struct Step {
        int kind;  // 0 separable generic (exact), 1 compose, 2 fused resample, 3 fused resample (tensor cores), 4 blur, 5 colour pass, 6 orientation pass, 7 vertical blur (tensor cores), 8 to_rgb8, 9 blur with both passes on the tensor cores; 10 / 11 / 12 / 13 = 0 / 1 / 6 / 8 for 16-bit and f32 subpixels (kernels_deep.cu)
        const fanlin::BlurItem *blur_items;
        uint32_t max_w, max_h, c, radius, taps_pad;
        const fanlin::FusedTcItem *tc_items;
        const fanlin::BlurVTcItem *bv_items;
        const fanlin::BlurTcItem *bt_items;
        const void *tmaps;  // CUtensorMap per tc item
        size_t smem;
        const fanlin::StageDesc *descs;
        fanlin::LaunchGeom geom;
        const fanlin::FusedItem *items;
        uint32_t n_items, variant, max_band;
        uint32_t n_paired;  // kind 4: 1 = the vertical pass was done by a kind-7 step
    };
    std::vector<Step> steps;
    std::shared_ptr<fanlin::TableBuf> tbuf;  // the device tables this batch launches from
    void *d_meta = nullptr;     // descriptors + tables
    cudaStream_t alloc_stream = nullptr;  // stream the device blocks were allocated on (stream-ordered pool)
    void *h_meta = nullptr;     // their pinned host copy (the upload is asynchronous)
    void *d_scratch = nullptr;  // intermediates (reused across chunks)
    const fanlin::TapEntry *d_tab = nullptr;
    const float *d_w = nullptr;
    const float *d_fw = nullptr;        // fused scatter tables
    const uint32_t *d_finfo = nullptr;
    const uint8_t *d_tb = nullptr;      // tensor-core weight digit tiles
    int launches_per_run = 0;
    bool timing = false;
    std::vector<cudaEvent_t> events;
    std::vector<const char *> ev_names;
    size_t ev_used = 0;
};
