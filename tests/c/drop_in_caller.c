/*
 * drop_in_caller.c -- a plain C caller written against aclib's public interface only (ac.h + imgconvert.h), the way
 * transcode's modules use it (e.g. libtcvideo/tcvideo.c:1058-1060, import/decode_lavc.c:310-312).  It is compiled
 * against include/ and linked against libacgpu.so unchanged: that is the drop-in claim.  It prints one FNV-1a digest
 * per operation; tests/test_drop_in_c.py compares them with digests of the checker's output for the same inputs.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ac.h"
#include "imgconvert.h"

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
    for (i = 0; i < n; i++) {                     /* splitmix-like bytes, mirrored in the Python test */
        z = z * 6364136223846793005ull + 1442695040888963407ull;
        p[i] = (uint8_t)(z >> 56);
    }
}

int main(int argc, char **argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 320, h = argc > 2 ? atoi(argv[2]) : 240;
    int accel = AC_ALL & ac_cpuinfo();
    uint8_t *yuv, *rgb, *yuv2, *line;
    uint8_t *src[3], *dst[3];
    size_t yuvsz = (size_t)w * h + 2 * UV_PLANE_SIZE(IMG_YUV420P, w, h);

    printf("accel %s\n", ac_flagstotext(accel));
    if (!ac_init(accel)) { fprintf(stderr, "ac_init failed\n"); return 2; }
    if (!IS_YUV_FORMAT(IMG_YUV420P) || !IS_RGB_FORMAT(IMG_RGB24) || IMG_YUV_DEFAULT != IMG_YUV420P) return 3;

    yuv = malloc(yuvsz); rgb = malloc((size_t)w * h * 4); yuv2 = malloc((size_t)w * h * 2); line = malloc((size_t)w * 3);
    fill(yuv, yuvsz, 1);

    YUV_INIT_PLANES(src, yuv, IMG_YUV420P, w, h);
    dst[0] = rgb;
    if (!ac_imgconvert(src, IMG_YUV420P, dst, IMG_RGB24, w, h)) return 4;
    printf("yuv420p_rgb24 %016llx\n", fnv(rgb, (size_t)w * h * 3));

    src[0] = rgb;
    YUV_INIT_PLANES(dst, yuv2, IMG_YUV422P, w, h);
    if (!ac_imgconvert(src, IMG_RGB24, dst, IMG_YUV422P, w, h)) return 5;
    printf("rgb24_yuv422p %016llx\n", fnv(yuv2, (size_t)w * h * 2));

    /* YV12 alias and an unknown pair */
    YUV_INIT_PLANES(src, yuv, IMG_YV12, w, h);
    dst[0] = rgb;
    if (!ac_imgconvert(src, IMG_YV12, dst, IMG_BGR24, w, h)) return 6;
    printf("yv12_bgr24 %016llx\n", fnv(rgb, (size_t)w * h * 3));
    if (ac_imgconvert(src, IMG_UNKNOWN, dst, IMG_RGB24, w, h)) return 7;

    /* deinterlace-style line ops (libtcvideo/tcvideo.c:353-362, :464-475) */
    ac_average(rgb, rgb + (size_t)w * 6, line, w * 3);
    printf("average %016llx\n", fnv(line, (size_t)w * 3));
    ac_rescale(rgb, rgb + (size_t)w * 3, line, w * 3, 49152, 16384);
    printf("rescale %016llx\n", fnv(line, (size_t)w * 3));
    ac_rescale(rgb, NULL, line, w * 3, 65536, 0);          /* copy branch must not touch src2 */
    printf("rescale_copy %016llx\n", fnv(line, (size_t)w * 3));
    ac_memcpy(rgb, rgb + 1, (size_t)w * 3);                 /* ascending overlapping copy */
    printf("memcpy %016llx\n", fnv(rgb, (size_t)w * 3));
    return 0;
}
