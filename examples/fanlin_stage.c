/*
 * fanlin_stage.c -- the C ABI of include/fanlin_device.h used from plain C, the way a host binding
 * (the Rust crate of INTEGRATION.md, cgo, JNI ...) drives it: what src/handler.rs:224-255 of fanlin-rs
 * does between decode and encode, for one image.
 *
 *   gcc -std=c99 -Iinclude examples/fanlin_stage.c -Lfanlin-rs_b200 -lfanlin_device -Wl,-rpath,$PWD/fanlin-rs_b200 -o fanlin_stage
 *   ./fanlin_stage in.ppm "w=300&h=200&rgb=32,32,32" out.pam     # the README's bench request of the reference
 *   ./fanlin_stage --plan 1920 1080 3 "w=300&h=200"               # host only: the output geometry, no device needed
 *
 * Input: binary PPM (P6, maxval 255) or PGM (P5).  Output: PAM (P7), which holds 1-4 channels -- the stage returns
 * Luma / LumaA / Rgb / Rgba exactly as the DynamicImage variant the reference would hold at that point.
 * There is no CPU fallback: without a CUDA device fanlin_init fails with FANLIN_ENODEVICE and so does this program
 * (exit status = the fanlin_status).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fanlin_device.h"

static int fail(const char *what, int rc) {
    fprintf(stderr, "%s: status %d: %s\n", what, rc, fanlin_last_error());
    return rc ? rc : 1;
}

static unsigned char *read_pnm(const char *path, uint32_t *w, uint32_t *h, uint32_t *c) {
    FILE *f = fopen(path, "rb");
    char magic[3] = {0};
    unsigned maxval = 0;
    unsigned char *px = NULL;
    if (!f) return NULL;
    if (fscanf(f, "%2s", magic) == 1 && (!strcmp(magic, "P6") || !strcmp(magic, "P5"))) {
        int ch;
        *c = magic[1] == '6' ? 3u : 1u;
        /* header fields, '#' comments allowed between them */
        for (int field = 0; field < 3; field++) {
            unsigned v = 0;
            while ((ch = fgetc(f)) != EOF) {
                if (ch == '#') { while ((ch = fgetc(f)) != EOF && ch != '\n') {} continue; }
                if (ch >= '0' && ch <= '9') { ungetc(ch, f); break; }
            }
            if (fscanf(f, "%u", &v) != 1) { fclose(f); return NULL; }
            if (field == 0) *w = v; else if (field == 1) *h = v; else maxval = v;
        }
        fgetc(f); /* the single whitespace byte in front of the raster */
        if (maxval == 255 && *w && *h) {
            size_t n = (size_t)*w * *h * *c;
            px = (unsigned char *)malloc(n);
            if (px && fread(px, 1, n, f) != n) { free(px); px = NULL; }
        }
    }
    fclose(f);
    return px;
}

static int write_pam(const char *path, const unsigned char *px, uint32_t w, uint32_t h, uint32_t c) {
    static const char *tupl[] = {"", "GRAYSCALE", "GRAYSCALE_ALPHA", "RGB", "RGB_ALPHA"};
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    fprintf(f, "P7\nWIDTH %u\nHEIGHT %u\nDEPTH %u\nMAXVAL 255\nTUPLTYPE %s\nENDHDR\n", w, h, c, tupl[c]);
    fwrite(px, 1, (size_t)w * h * c, f);
    return fclose(f);
}

/* query string -> the request fields of a job: what handler.rs reads off &query::Query */
static int job_from_query_string(const char *qs, fanlin_job *job) {
    fanlin_query q;
    int rc = fanlin_query_parse(qs, &q);
    if (rc != FANLIN_OK) return rc;
    memset(job, 0, sizeof *job);
    fanlin_job_from_query(&q, /*gif=*/0, job);
    return FANLIN_OK;
}

static void print_plan(const fanlin_plan *p) {
    printf("out %ux%u x %u channels (%llu bytes); resized %ux%u, crop at (%u, %u), overlay at (%u, %u); "
           "source window x [%u, %u) y [%u, %u); stages 0x%x; algorithmic bytes %llu\n",
           p->out_w, p->out_h, p->out_channels, (unsigned long long)p->out_bytes, p->resized_w, p->resized_h, p->crop_x,
           p->crop_y, p->overlay_x, p->overlay_y, p->src_x0, p->src_x1, p->src_y0, p->src_y1, p->stages,
           (unsigned long long)p->algorithmic_bytes);
}

int main(int argc, char **argv) {
    fanlin_job job;
    fanlin_plan plan;
    int rc;
    if (fanlin_abi_version() != FANLIN_ABI_VERSION) {
        fprintf(stderr, "header is ABI %d, library is ABI %d\n", FANLIN_ABI_VERSION, fanlin_abi_version());
        return 1;
    }
    if (argc == 6 && !strcmp(argv[1], "--plan")) { /* host only */
        if ((rc = job_from_query_string(argv[5], &job)) != FANLIN_OK) return fail("query", rc);
        job.src_w = (uint32_t)atoi(argv[2]);
        job.src_h = (uint32_t)atoi(argv[3]);
        job.src_channels = (uint32_t)atoi(argv[4]);
        if ((rc = fanlin_plan_job(&job, &plan)) != FANLIN_OK) return fail("plan", rc);
        print_plan(&plan);
        return 0;
    }
    if (argc != 4) {
        fprintf(stderr, "usage: %s in.ppm|in.pgm \"w=300&h=200&...\" out.pam\n       %s --plan W H CHANNELS \"query\"\n", argv[0], argv[0]);
        return 1;
    }
    uint32_t w = 0, h = 0, c = 0;
    unsigned char *src = read_pnm(argv[1], &w, &h, &c);
    if (!src) { fprintf(stderr, "%s: not a binary PPM / PGM with maxval 255\n", argv[1]); return 1; }
    if ((rc = job_from_query_string(argv[2], &job)) != FANLIN_OK) return fail("query", rc);
    job.src = src; job.src_w = w; job.src_h = h; job.src_channels = c;
    if ((rc = fanlin_plan_job(&job, &plan)) != FANLIN_OK) return fail("plan", rc);
    print_plan(&plan);

    fanlin_ctx *ctx = NULL;
    if ((rc = fanlin_init(NULL, 0, NULL, &ctx)) != FANLIN_OK) return fail("fanlin_init", rc); /* no device: no fallback */
    unsigned char *dst = (unsigned char *)malloc(plan.out_bytes);
    job.dst = dst; job.dst_capacity = plan.out_bytes;
    rc = fanlin_run(ctx, &job, 1, NULL);
    if (rc == FANLIN_OK) {
        fanlin_stats st;
        fanlin_get_stats(ctx, &st);
        printf("%llu kernel launch(es), %llu bytes to the device, %llu back\n", (unsigned long long)st.kernel_launches,
               (unsigned long long)st.h2d_bytes, (unsigned long long)st.d2h_bytes);
        if (write_pam(argv[3], dst, plan.out_w, plan.out_h, plan.out_channels)) { fprintf(stderr, "cannot write %s\n", argv[3]); rc = 1; }
    } else {
        fail("fanlin_run", rc);
    }
    fanlin_shutdown(ctx);
    free(dst);
    free(src);
    return rc;
}
