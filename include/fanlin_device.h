/*
 * fanlin_device.h -- C ABI of the B200-native pixel-transform stage of fanlin-rs.
 *
 * The reference has no FFI for this path; the seam is internal Rust.  Each entry
 * point below names the reference lines it replaces (paths under the reference
 * repository).  A Rust `extern "C"` block binding exactly these symbols is shown
 * in INTEGRATION.md.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * Pixel buffers are row-major, interleaved channels (1 = Luma, 2 = LumaA,
 * 3 = Rgb, 4 = Rgba), i.e. image::ImageBuffer::as_raw() (tight) unless a pitch
 * is given.  Subpixels are u8 (the north_star's path: every tensor-core kernel)
 * or, for the DynamicImage variants a 16-bit PNG / TIFF or an HDR / EXR file
 * decodes to (src/handler.rs:219), u16 / f32 in native byte order
 * (fanlin_job.src_sample; generic kernels in the crate's operation order).
 */
#ifndef FANLIN_DEVICE_H
#define FANLIN_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FANLIN_ABI_VERSION 2 /* 2: fanlin_job.src_sample, fanlin_plan.out_sample (16-bit and f32 DynamicImage variants) */

#if defined(__GNUC__)
#define FANLIN_API __attribute__((visibility("default")))
#else
#define FANLIN_API
#endif

typedef struct fanlin_ctx fanlin_ctx;
typedef struct fanlin_batch fanlin_batch;

enum fanlin_status {
    FANLIN_OK = 0,
    FANLIN_EINVAL = 1,     /* bad job / argument */
    FANLIN_ENOMEM = 2,     /* host or device allocation failed */
    FANLIN_ECAPACITY = 3,  /* dst_capacity smaller than the planned output */
    FANLIN_ECUDA = 4,      /* CUDA runtime error (message in fanlin_last_error) */
    FANLIN_ENODEVICE = 5,  /* no usable CUDA device: there is NO CPU fallback */
    FANLIN_ESHUTDOWN = 6
};

/* image::imageops::FilterType as used by the reference: Lanczos3 for stills
 * (src/handler.rs:233,235), Nearest for GIF frames (src/handler.rs:338,340). */
enum fanlin_filter { FANLIN_FILTER_NEAREST = 0, FANLIN_FILTER_LANCZOS3 = 1 };

/* Subpixel type of a DynamicImage variant: ImageLuma8 .. ImageRgba8, ImageLuma16 .. ImageRgba16, ImageRgb32F /
 * ImageRgba32F (f32 has no Luma variants: src_channels 3 or 4). */
enum fanlin_sample { FANLIN_SAMPLE_U8 = 0, FANLIN_SAMPLE_U16 = 1, FANLIN_SAMPLE_F32 = 2 };

enum fanlin_flags {
    FANLIN_GRAYSCALE = 1u << 0, /* Query::grayscale()  src/query.rs:64-66; wins over INVERSE (handler.rs:224-228) */
    FANLIN_INVERSE = 1u << 1,   /* Query::inverse()    src/query.rs:68-70 */
    FANLIN_HAS_DIMS = 1u << 2,  /* Query::dimensions() is Some((req_w, req_h))  src/query.rs:28-33 */
    FANLIN_CROP = 1u << 3,      /* Query::cropping()   src/query.rs:55-57 -> resize_to_fill */
    FANLIN_TO_RGBA8 = 1u << 4,  /* GIF frames end with img.to_rgba8()  src/handler.rs:355; also the WebP branch's into_rgba8() (:287) */
    FANLIN_TO_RGB8 = 1u << 5,   /* the JPEG branch: the encoder works on RGB8 (src/handler.rs:274-278), DynamicImage::to_rgb8 --
                                   alpha dropped, luma replicated; not together with FANLIN_TO_RGBA8.  For request-sized batches
                                   folded into the epilogue of the kernel that writes the final image (no extra pass) */
    FANLIN_TO_YCBCR = 1u << 6   /* the JPEG branch, one step further: the result as the three full-resolution
                                   planes Y, Cb, Cr the JPEG encoder derives from it -- image-0.25.6 codecs/jpeg/encoder.rs
                                   rgb_to_ycbcr on to_rgb8() of the result (f32, JFIF coefficients x 255, truncating cast);
                                   dst holds [3][out_h][out_w] u8, out_channels = 3, fanlin_plan.stages bit 6.  For a host
                                   encoder that accepts planar input; excludes TO_RGBA8 / TO_RGB8 */
};

/* One image (still) or one composited GIF frame and the request parameters the
 * stage reads.  Replaces the owned DynamicImage + &query::Query that
 * src/handler.rs:224-255 and :329-355 operate on. */
typedef struct fanlin_job {
    const uint8_t *src;    /* host pointer for fanlin_run, device pointer for fanlin_batch_* */
    uint32_t src_w, src_h;
    uint32_t src_channels; /* 1..4 */
    uint32_t src_pitch;    /* bytes per row; 0 = src_w * src_channels */
    uint32_t flags;        /* enum fanlin_flags */
    uint32_t filter;       /* enum fanlin_filter */
    uint32_t req_w, req_h; /* used when FANLIN_HAS_DIMS */
    uint8_t fill_rgb[3];   /* Query::fill_color()  src/query.rs:35-49 */
    uint8_t orientation;   /* EXIF orientation of a decoded still: 0 / 1 = none, 2..8 as ImageDecoder::orientation()
                              reports it; replaces img.apply_orientation(o) at src/handler.rs:206,221-223.  The source
                              fields describe the image as stored; the stage sees it oriented.  Leave it 0 on the GIF path
                              (process_gif never reads EXIF). */
    float blur_sigma;      /* Query::blur(): 0 = off, else already clamped to [10,20]  src/query.rs:59-62 */
    uint8_t *dst;          /* host (fanlin_run) or device (fanlin_batch_*) pointer, tight rows */
    uint64_t dst_capacity; /* bytes available at dst */
    uint32_t src_sample;   /* enum fanlin_sample of src (0 = u8).  src_pitch stays in bytes; src must be aligned to the
                              subpixel size.  The result keeps the subpixel type except behind a letterbox (the reference's
                              canvas is Rgba<u8>, handler.rs:240-247) and TO_RGBA8 / TO_RGB8: fanlin_plan.out_sample says which */
    uint32_t reserved;     /* 0 */
} fanlin_job;

/* What the stage will produce for a job: lets the caller allocate dst and pick
 * the DynamicImage variant to rebuild. */
typedef struct fanlin_plan {
    uint32_t out_w, out_h, out_channels;
    uint32_t resized_w, resized_h; /* resize_dimensions() result before crop; 0 if no resample */
    uint32_t crop_x, crop_y;       /* resize_to_fill crop origin inside the resized image */
    uint32_t overlay_x, overlay_y; /* letterbox offset (handler.rs:244-245) */
    uint32_t src_x0, src_y0, src_x1, src_y1; /* source window the output depends on (in the oriented image) */
    uint32_t stages;               /* bit0 colour op, bit1 resample, bit2 letterbox, bit3 blur, bit4 to_rgba8, bit5 to_rgb8, bit6 planar YCbCr */
    uint64_t out_bytes;
    uint64_t algorithmic_bytes;    /* src window bytes + out_bytes */
    uint32_t out_sample;           /* enum fanlin_sample of the output (out_bytes = out_w * out_h * out_channels * its size) */
    uint32_t reserved;
} fanlin_plan;

typedef struct fanlin_config {
    uint32_t struct_size;        /* sizeof(fanlin_config), for forward compatibility */
    uint32_t exact;              /* 1: crate operation order, no FMA contraction (bit-exact to the CPU path, slower) */
    uint64_t device_scratch_bytes; /* per-device arena for staging + intermediates; 0 = default */
    uint64_t pinned_bytes;       /* per-device pinned staging ring; 0 = default */
    uint32_t batch_window_us;    /* request batcher: 0 = default, no idle wait (what is queued is dispatched at once; requests that
                                    arrive while a batch runs are merged into the next one); > 0: the collector waits this long
                                    after the first arrival before it dispatches */
    uint32_t max_batch_jobs;     /* 0 = default */
    uint32_t vertical_path;      /* 0 = tensor cores, both Lanczos3 passes where the geometry allows (tables are cached per geometry in the context, so single requests take them too); 1 = CUDA cores only; 2 = tensor cores for the vertical pass only; 3 = same as 0 (kept from ABI 1, where 0 needed >= 256 jobs per batch) */
    uint32_t blur_path;          /* 0 = both blur passes on the tensor cores where eligible (no f32 intermediate in HBM); 1 = the two-kernel blur (vertical pass on the tensor cores, horizontal on the CUDA cores) */
} fanlin_config;

typedef struct fanlin_stats {
    uint64_t kernel_launches;
    uint64_t jobs;
    uint64_t batches;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t table_bytes; /* filter-table bytes uploaded so far: a batch whose geometries the context has seen adds none */
} fanlin_stats;

FANLIN_API int fanlin_abi_version(void);

/* Pure host.  Output geometry of one job: restates the guards of
 * src/handler.rs:231,238,252 and image's resize_dimensions / resize_to_fill. */
FANLIN_API int fanlin_plan_job(const fanlin_job *job, fanlin_plan *plan);

/* Create a context on the given CUDA device ordinals (n_devices = 0: all visible
 * devices).  Fails with FANLIN_ENODEVICE when there is none.  Called once at
 * start-up, next to State::new (src/main.rs:63-66); the handle is Send + Sync. */
FANLIN_API int fanlin_init(const int *device_ids, int n_devices, const fanlin_config *cfg, fanlin_ctx **out);
FANLIN_API void fanlin_shutdown(fanlin_ctx *ctx);
FANLIN_API int fanlin_device_count(const fanlin_ctx *ctx);

/* Blocking, re-entrant.  Runs n_jobs host-resident jobs: replaces the body of
 * src/handler.rs:224-255 (n_jobs = 1 per request; concurrent callers are merged
 * by the batcher) and the per-frame closure of :321-357 (n_jobs = frame count).
 * plans may be NULL. */
FANLIN_API int fanlin_run(fanlin_ctx *ctx, const fanlin_job *jobs, uint32_t n_jobs, fanlin_plan *plans);

/* Device-resident batches (src/dst are device pointers on device_index):
 * prepare once (plans, weight tables and ragged descriptors uploaded), launch
 * many times on a caller-provided cudaStream_t (NULL = the context's stream). */
FANLIN_API int fanlin_batch_prepare(fanlin_ctx *ctx, int device_index, const fanlin_job *jobs, uint32_t n_jobs,
                         fanlin_plan *plans, fanlin_batch **out);
FANLIN_API int fanlin_batch_launch(fanlin_batch *batch, void *cuda_stream);
/* Kernels one launch enqueues, and the name-independent index of the dominant one. */
FANLIN_API int fanlin_batch_launch_count(const fanlin_batch *batch);
/* The caller must have synchronised every stream it launched the batch on. */
FANLIN_API void fanlin_batch_free(fanlin_batch *batch);
/* Per-kernel device times of a batch (benchmarks): when enabled, fanlin_batch_launch
 * brackets every kernel with CUDA events on the launching stream;
 * fanlin_batch_kernel_times returns, after the caller synchronised that stream,
 * the kernels launched since the previous call (name + milliseconds), up to cap,
 * and resets the record; the return value is the number of kernels recorded. */
FANLIN_API int fanlin_batch_set_timing(fanlin_batch *batch, int enable);
FANLIN_API int fanlin_batch_kernel_times(fanlin_batch *batch, const char **names, float *ms, int cap);

/* YCCK -> CMYK of a decoded Adobe JPEG, the float loop of convert_jpeg_color_if_needed
 * (src/handler.rs:420-439): n_pixels 4-byte pixels {Y, Cb, Cr, K} -> {C', M', Y', 255 - K} with
 * the reference's f32 expression order, clamp and truncating cast -- bit-exact.  src == dst is allowed (the reference
 * converts in place).  The lcms2 CMYK -> RGB transform that follows (:469-493) stays on the host.
 * fanlin_ycck_to_cmyk: host buffers, blocking (chunks alternate between two streams so copies overlap the kernel);
 * fanlin_ycck_to_cmyk_device: device pointers on device_index, asynchronous on cuda_stream (NULL = the context's). */
FANLIN_API int fanlin_ycck_to_cmyk(fanlin_ctx *ctx, const uint8_t *src, uint8_t *dst, uint64_t n_pixels);
FANLIN_API int fanlin_ycck_to_cmyk_device(fanlin_ctx *ctx, int device_index, const uint8_t *src, uint8_t *dst, uint64_t n_pixels,
                                          void *cuda_stream);

/* Pinned host buffers from the context's pool, so decoders can write pixels
 * where the copy engine can read them without a staging memcpy. */
FANLIN_API void *fanlin_host_alloc(fanlin_ctx *ctx, size_t bytes);
FANLIN_API void fanlin_host_free(fanlin_ctx *ctx, void *p);

FANLIN_API int fanlin_get_stats(const fanlin_ctx *ctx, fanlin_stats *out);

/* How fanlin_run (and bench.py across ranks) shards n_jobs independent images over
 * n_shards devices: contiguous blocks by image index, no collective.
 * Pure host. */
FANLIN_API void fanlin_shard_range(uint32_t n_jobs, uint32_t n_shards, uint32_t shard, uint32_t *lo, uint32_t *hi);

/* Thread-local message of the last failure on this thread. */
FANLIN_API const char *fanlin_last_error(void);

/* ---- host-side mirror of query::Query (src/query.rs:3-94) ---------------- */

typedef struct fanlin_query {
    uint32_t has_w, w, has_h, h;
    uint32_t has_rgb;
    char rgb[64];
    uint32_t has_quality, quality;
    uint32_t has_crop, crop;
    uint32_t has_blur, blur;
    uint32_t has_grayscale, grayscale;
    uint32_t has_inverse, inverse;
    uint32_t has_avif, avif;
    uint32_t has_webp, webp;
} fanlin_query;

/* Parses "w=300&h=200&rgb=32,32,32" (or a URL with '?').  Returns FANLIN_EINVAL
 * where serde's Query extractor errors (src/query.rs:128-135, :294-301). */
FANLIN_API int fanlin_query_parse(const char *query_string, fanlin_query *q);
FANLIN_API int fanlin_query_dimensions(const fanlin_query *q, uint32_t *w, uint32_t *h); /* 1 = Some */
FANLIN_API void fanlin_query_fill_color(const fanlin_query *q, uint8_t rgb[3]);
FANLIN_API float fanlin_query_blur(const fanlin_query *q);
FANLIN_API int fanlin_query_as_is(const fanlin_query *q);
FANLIN_API int fanlin_query_unsupported_scale_size(const fanlin_query *q);

/* Fills the request fields of a job from a query, for a still image
 * (gif = 0: Lanczos3, blur honoured) or a GIF frame (gif = 1: Nearest, no blur,
 * TO_RGBA8) -- the two call patterns of src/handler.rs:224-255 and :329-355. */
FANLIN_API void fanlin_job_from_query(const fanlin_query *q, int gif, fanlin_job *job);

#ifdef __cplusplus
}
#endif
#endif /* FANLIN_DEVICE_H */
