/*
 * tcv_caller.c -- a plain C caller written against libtcvideo's public interface only (tcvideo.h), shaped like transcode's
 * do_process_frame (src/video_trans.c:192-426): one YUV420P frame and one RGB24 frame go through clip, deinterlace,
 * resize, reduce, flips, gamma, antialias and tcv_convert, every PROCESS_FRAME stage once per plane.  The SAME source is
 * built twice by tests/test_tcv_shim.py:
 *     tcv_caller      against include/tcvideo.h,                  linked with libtcvgpu + libacgpu   (this repo)
 *     tcv_caller_ref  against /root/reference/libtcvideo/tcvideo.h, linked with oracle/_ref/libtcv_ref (the reference)
 * and the two must print the same digests: that is the drop-in claim for the tcv_* interface.
 * With "time N" it also times N frames of `-I 5`-style linear blend on a 1920x1080 luma plane through whichever library
 * it was linked with.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "tcvideo.h"

int ac_init(int accel);      /* aclib/ac.h:59; both link targets export it */

static unsigned long long fnv(const uint8_t *p, size_t n)
{
    unsigned long long h = 1469598103934665603ull;
    size_t i;
    for (i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

static void fill(uint8_t *p, size_t n, unsigned seed)
{
    size_t i;
    unsigned long long z = seed;
    for (i = 0; i < n; i++) {
        z = z * 6364136223846793005ull + 1442695040888963407ull;
        p[i] = (uint8_t)(z >> 56);
    }
}

#define CHECK(call) do { if (!(call)) { fprintf(stderr, "failed: %s\n", #call); return 1; } } while (0)

/* one PROCESS_FRAME-style stage over the three planes of a YUV420P frame (video_trans.c:37-46) */
#define PLANES420(w, h) const int pw[3] = {(w), (w) / 2, (w) / 2}, ph[3] = {(h), (h) / 2, (h) / 2}

int main(int argc, char **argv)
{
    TCVHandle tcv;
    int w = 352, h = 288, i;
    uint8_t *a, *b, *t;
    size_t cap;

    if (argc > 2 && strcmp(argv[1], "time") == 0) {
        const int n = atoi(argv[2]), W = 1920, H = 1080;
        uint8_t *s = malloc((size_t)W * H), *d = malloc((size_t)W * H);
        struct timespec t0, t1;
        ac_init(-1);
        CHECK(tcv = tcv_init());
        fill(s, (size_t)W * H, 9);
        CHECK(tcv_deinterlace(tcv, s, d, W, H, 1, TCV_DEINTERLACE_LINEAR_BLEND));
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (i = 0; i < n; i++) CHECK(tcv_deinterlace(tcv, s, d, W, H, 1, TCV_DEINTERLACE_LINEAR_BLEND));
        clock_gettime(CLOCK_MONOTONIC, &t1);
        printf("linear_blend_1080p_y frames_per_s %.1f\n", n / ((t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9));
        tcv_free(tcv);
        return 0;
    }
    if (argc > 2) { w = atoi(argv[1]); h = atoi(argv[2]); }
    cap = (size_t)(w + 64) * (h + 64) * 4;
    a = malloc(cap); b = malloc(cap);
    ac_init(-1);                                   /* src/transcode.c:2495 */
    CHECK(tcv = tcv_init());

    /* ---- a YUV420P frame through -j, -I 4, -B, -r, -z, -l, -G, -C ------------------------------------------------ */
    fill(a, (size_t)w * h * 3 / 2, 1);
    memset(b, 0x55, cap);
    {
        int cw = w, ch = h, nw, nh;
        size_t so, dofs;
        /* clip 8 px left/right, 16 rows top/bottom */
        nw = cw - 16; nh = ch - 32;
        { PLANES420(cw, ch); so = dofs = 0;
          for (i = 0; i < 3; i++) {
              CHECK(tcv_clip(tcv, a + so, b + dofs, pw[i], ph[i], 1, 8 / (i ? 2 : 1), 8 / (i ? 2 : 1), 16 / (i ? 2 : 1), 16 / (i ? 2 : 1), i ? 128 : 0));
              so += (size_t)pw[i] * ph[i]; dofs += (size_t)(i ? nw / 2 : nw) * (i ? nh / 2 : nh);
          } }
        cw = nw; ch = nh; t = a; a = b; b = t;
        printf("clip %016llx\n", fnv(a, (size_t)cw * ch * 3 / 2));
        /* drop the bottom field on every plane */
        nh = ch / 2;
        { PLANES420(cw, ch); so = dofs = 0;
          for (i = 0; i < 3; i++) {
              CHECK(tcv_deinterlace(tcv, a + so, b + dofs, pw[i], ph[i], 1, TCV_DEINTERLACE_DROP_FIELD_BOTTOM));
              so += (size_t)pw[i] * ph[i]; dofs += (size_t)pw[i] * (i ? nh / 2 : nh);
          } }
        ch = nh; t = a; a = b; b = t;
        printf("dropfield %016llx\n", fnv(a, (size_t)cw * ch * 3 / 2));
        /* rows +16 (resize_h = 2), then columns -16 (resize_w = -2) */
        nh = ch + 16;
        { PLANES420(cw, ch); so = dofs = 0;
          for (i = 0; i < 3; i++) {
              CHECK(tcv_resize(tcv, a + so, b + dofs, pw[i], ph[i], 1, 0, 2, i ? 4 : 8, i ? 4 : 8));
              so += (size_t)pw[i] * ph[i]; dofs += (size_t)pw[i] * (i ? nh / 2 : nh);
          } }
        ch = nh; t = a; a = b; b = t;
        nw = cw - 16;
        { PLANES420(cw, ch); so = dofs = 0;
          for (i = 0; i < 3; i++) {
              CHECK(tcv_resize(tcv, a + so, b + dofs, pw[i], ph[i], 1, -2, 0, i ? 4 : 8, i ? 4 : 8));
              so += (size_t)pw[i] * ph[i]; dofs += (size_t)(i ? nw / 2 : nw) * ph[i];
          } }
        cw = nw; t = a; a = b; b = t;
        printf("resize %016llx\n", fnv(a, (size_t)cw * ch * 3 / 2));
        /* flip both ways, in place as video_trans.c does via the second buffer; here src == dest on purpose */
        { PLANES420(cw, ch); so = 0;
          for (i = 0; i < 3; i++) {
              CHECK(tcv_flip_v(tcv, a + so, a + so, pw[i], ph[i], 1));
              CHECK(tcv_flip_h(tcv, a + so, a + so, pw[i], ph[i], 1));
              so += (size_t)pw[i] * ph[i];
          } }
        printf("flips %016llx\n", fnv(a, (size_t)cw * ch * 3 / 2));
        /* gamma on luma in place (video_trans.c:393-395), antialias luma into the second buffer */
        CHECK(tcv_gamma_correct(tcv, a, a, cw, ch, 1, 0.8));
        CHECK(tcv_antialias(tcv, a, b, cw, ch, 1, 0.333, 0.5));
        printf("gamma_aa %016llx %016llx\n", fnv(a, (size_t)cw * ch), fnv(b, (size_t)cw * ch));
        /* reduce luma 2x2 */
        CHECK(tcv_reduce(tcv, a, b, cw, ch, 1, 2, 2));
        printf("reduce %016llx\n", fnv(b, (size_t)(cw / 2) * (ch / 2)));
    }

    /* ---- an RGB24 frame: -I 5, tcv_convert both ways and in place, -K as the two conversions ---------------------- */
    fill(a, (size_t)w * h * 3, 2);
    CHECK(tcv_deinterlace(tcv, a, b, w, h, 3, TCV_DEINTERLACE_INTERPOLATE));
    printf("interpolate_rgb %016llx\n", fnv(b, (size_t)w * h * 3));
    fill(a, (size_t)w * h * 3, 2);
    CHECK(tcv_deinterlace(tcv, a, b, w, h, 3, TCV_DEINTERLACE_LINEAR_BLEND));
    printf("blend_rgb %016llx\n", fnv(b, (size_t)w * h * 3));
    CHECK(tcv_convert(tcv, b, a, w, h, IMG_RGB24, IMG_YUV420P));
    printf("rgb_yuv420p %016llx\n", fnv(a, (size_t)w * h * 3 / 2));
    CHECK(tcv_convert(tcv, a, a, w, h, IMG_YUV420P, IMG_RGB24));          /* in place through a temporary */
    printf("yuv420p_rgb_inplace %016llx\n", fnv(a, (size_t)w * h * 3));
    CHECK(tcv_convert(tcv, a, b, w, h, IMG_RGB24, IMG_GRAY8));
    CHECK(tcv_convert(tcv, b, a, w, h, IMG_GRAY8, IMG_RGB24));
    printf("decolor %016llx\n", fnv(a, (size_t)w * h * 3));
    CHECK(tcv_convert(tcv, a, b, w, h, IMG_RGB24, IMG_RGB24));
    printf("copy %016llx\n", fnv(b, (size_t)w * h * 3));

    /* ---- rejections return 0 (tcvideo.c:192-202, 690-699, 848-851, 1008-1015) ------------------------------------- */
    printf("rejects %d %d %d %d %d\n",
           tcv_clip(tcv, a, b, w, h, 1, w, 0, 0, 0, 0), tcv_reduce(tcv, a, b, w, h, 1, 0, 1),
           tcv_gamma_correct(tcv, a, b, w, h, 1, 0.0), tcv_convert(NULL, a, b, w, h, IMG_RGB24, IMG_BGR24),
           tcv_resize(tcv, a, b, w, h, 1, 0, 1, 3, 8));
    printf("zoom_names %s %s %d %d\n", tcv_zoom_filter_to_string(TCV_ZOOM_DEFAULT), tcv_zoom_filter_to_string(TCV_ZOOM_B_SPLINE),
           (int)tcv_zoom_filter_from_string("cubic_keys4"), (int)tcv_zoom_filter_from_string("nope"));
    tcv_free(tcv);
    free(a); free(b);
    return 0;
}
