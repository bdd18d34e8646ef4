/*
 * tcv_shim.c -- libtcvgpu.so: libtcvideo's C interface (include/tcvideo.h, reference libtcvideo/tcvideo.h:54-98) as a
 * thin layer over libacgpu's frame-granular entry points.  Plain C, no CUDA in this file: every function checks what the
 * reference checks first (handle), then makes ONE libacgpu call for the whole plane.  The argument rules, error returns
 * and results of that call are the reference's (csrc/host_tcv.cu cites tcvideo.c line by line).
 */
#include <stdio.h>
#include <stdlib.h>
#include <strings.h>

#include "acgpu.h"
#include "tcvideo.h"

struct tcvhandle_ {
    int magic;
};

static void complain(const char *who)
{
    /* the reference logs every rejection through tc_log_error("libtcvideo", ...) (tcvideo.c:192-202 etc.) */
    fprintf(stderr, "[libtcvgpu] %s: %s\n", who, acgpu_last_error());
}

static int done(const char *who, int ok)
{
    /* the reference is synchronous: the caller may read dest as soon as the function returns */
    if (ok && !acgpu_stream_sync(NULL)) ok = 0;
    if (!ok) complain(who);
    return ok ? 1 : 0;
}

TCVHandle tcv_init(void)
{
    TCVHandle h;
    if (!(ac_cpuinfo() & AC_CUDA) || !ac_init(AC_CUDA)) {
        fprintf(stderr, "[libtcvgpu] tcv_init: no usable CUDA device (libtcvgpu has no CPU implementation)\n");
        return NULL;
    }
    h = calloc(1, sizeof(*h));
    if (h) h->magic = 0x74637667;
    return h;
}

void tcv_free(TCVHandle handle) { free(handle); }

int tcv_clip(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
             int clip_left, int clip_right, int clip_top, int clip_bottom, uint8_t black_pixel)
{
    (void)handle;
    return done("tcv_clip", acgpu_clip_batch(src, dest, width, height, Bpp, clip_left, clip_right, clip_top, clip_bottom,
                                             black_pixel, 0, 0, 1, NULL));
}

int tcv_deinterlace(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp, TCVDeinterlaceMode mode)
{
    int m;
    (void)handle;
    switch (mode) {       /* tcvideo.c:297-310 */
    case TCV_DEINTERLACE_DROP_FIELD_TOP:    m = ACGPU_DEINT_DROP_FIELD_TOP; break;
    case TCV_DEINTERLACE_DROP_FIELD_BOTTOM: m = ACGPU_DEINT_DROP_FIELD_BOTTOM; break;
    case TCV_DEINTERLACE_INTERPOLATE:       m = ACGPU_DEINT_INTERPOLATE; break;
    case TCV_DEINTERLACE_LINEAR_BLEND:      m = ACGPU_DEINT_LINEAR_BLEND; break;
    default:                                m = -1; break;      /* rejected below as "invalid mode" */
    }
    return done("tcv_deinterlace", acgpu_deinterlace_batch(src, dest, width, height, Bpp, m, 0, 0, 1, NULL));
}

int tcv_resize(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
               int resize_w, int resize_h, int scale_w, int scale_h)
{
    (void)handle;      /* the reference keeps its weight tables in the handle (tcvideo.c:1115-1165); libacgpu caches them per thread */
    return done("tcv_resize", acgpu_resize_batch(src, dest, width, height, Bpp, resize_w, resize_h, scale_w, scale_h, 0, 0, 1, NULL));
}

int tcv_reduce(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int reduce_w, int reduce_h)
{
    (void)handle;
    return done("tcv_reduce", acgpu_reduce_batch(src, dest, width, height, Bpp, reduce_w, reduce_h, 0, 0, 1, NULL));
}

int tcv_flip_v(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp)
{
    (void)handle;
    return done("tcv_flip_v", acgpu_flip_v_batch(src, dest, width, height, Bpp, 0, 0, 1, NULL));
}

int tcv_flip_h(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp)
{
    (void)handle;
    return done("tcv_flip_h", acgpu_flip_h_batch(src, dest, width, height, Bpp, 0, 0, 1, NULL));
}

int tcv_gamma_correct(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double gamma)
{
    (void)handle;
    return done("tcv_gamma_correct", acgpu_gamma_correct_batch(src, dest, width, height, Bpp, gamma, 0, 0, 1, NULL));
}

int tcv_antialias(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double weight, double bias)
{
    (void)handle;
    return done("tcv_antialias", acgpu_antialias_batch(src, dest, width, height, Bpp, weight, bias, 0, 0, 1, NULL));
}

int tcv_convert(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, ImageFormat srcfmt, ImageFormat destfmt)
{
    if (!handle) {        /* tcvideo.c:1008-1011: the one function that tests its handle */
        fprintf(stderr, "[libtcvgpu] tcv_convert(): No handle given!\n");
        return 0;
    }
    return done("tcv_convert", acgpu_convert_batch(src, dest, width, height, srcfmt, destfmt, 0, 0, 1, NULL));
}

int tcv_zoom(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
             int new_w, int new_h, TCVZoomFilter filter)
{
    (void)handle; (void)src; (void)dest; (void)width; (void)height; (void)Bpp; (void)new_w; (void)new_h; (void)filter;
    fprintf(stderr, "[libtcvgpu] tcv_zoom: not provided (the filtered resampler of libtcvideo/zoom.c is outside the aclib path); "
                    "build transcode with the reference's zoom.c for -Z / -I 3\n");
    return 0;
}

static const struct { TCVZoomFilter id; const char *name; } zoom_names[] = {
    {TCV_ZOOM_BELL, "Bell"}, {TCV_ZOOM_BOX, "Box"}, {TCV_ZOOM_B_SPLINE, "B_spline"}, {TCV_ZOOM_HERMITE, "Hermite"},
    {TCV_ZOOM_LANCZOS3, "Lanczos3"}, {TCV_ZOOM_MITCHELL, "Mitchell"}, {TCV_ZOOM_TRIANGLE, "Triangle"},
    {TCV_ZOOM_CUBIC_KEYS4, "Cubic_Keys4"}, {TCV_ZOOM_SINC8, "Sinc8"},
};

const char *tcv_zoom_filter_to_string(TCVZoomFilter filter)
{
    size_t i;
    if (filter == TCV_ZOOM_DEFAULT) return "Lanczos3";          /* zoom.c:100-101 */
    for (i = 0; i < sizeof(zoom_names) / sizeof(zoom_names[0]); i++)
        if (zoom_names[i].id == filter) return zoom_names[i].name;
    return NULL;
}

TCVZoomFilter tcv_zoom_filter_from_string(const char *name)
{
    size_t i;
    if (!name) return TCV_ZOOM_NULL;
    if (strcasecmp(name, "default") == 0) return TCV_ZOOM_LANCZOS3;       /* zoom.c:140-141 */
    for (i = 0; i < sizeof(zoom_names) / sizeof(zoom_names[0]); i++)
        if (strcasecmp(name, zoom_names[i].name) == 0) return zoom_names[i].id;
    return TCV_ZOOM_NULL;
}
