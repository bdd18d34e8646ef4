/*
 * include/tcvideo.h -- libtcvgpu: the interface of transcode's libtcvideo (libtcvideo/tcvideo.h:27-98) over libacgpu.
 *
 * Same names, argument lists, enum values and return convention (non-zero = success, 0 = rejected) as the reference
 * header, so libtcvideo's callers (src/video_trans.c:192-426, the filter modules, import/export modules calling
 * tcv_convert) compile and link unchanged against libtcvgpu + libacgpu instead of libtcvideo.la + libac.la.
 * Every function works on one plane (Bpp 1 or 3) or one image held in HOST memory -- what those callers pass -- with one
 * upload, one launch and one download per call instead of one aclib call per row; device pointers from acgpu_malloc are
 * accepted as well and processed in place in HBM.  Calls return after the result has landed (the reference is synchronous).
 *
 * Differences from the reference, all documented in DESIGN.md: tcv_deinterlace's linear blend leaves `src` intact (the
 * reference destroys it, tcvideo.c:368-389); tcv_zoom is not provided (see below).
 */
#ifndef LIBTCVGPU_TCVIDEO_H
#define LIBTCVGPU_TCVIDEO_H

#include <stdint.h>

#include "imgconvert.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque per-caller state (libtcvideo/tcvideo.h:24-27).  libtcvgpu keeps its tables inside libacgpu, per calling
 * thread, so one handle may be shared by threads; it only records that tcv_init() was called. */
typedef struct tcvhandle_ *TCVHandle;

/* libtcvideo/tcvideo.h:29-35 */
typedef enum {
    TCV_DEINTERLACE_DROP_FIELD_TOP,
    TCV_DEINTERLACE_DROP_FIELD_BOTTOM,
    TCV_DEINTERLACE_INTERPOLATE,
    TCV_DEINTERLACE_LINEAR_BLEND
} TCVDeinterlaceMode;

/* libtcvideo/tcvideo.h:37-50 */
typedef enum {
    TCV_ZOOM_DEFAULT = 0,
    TCV_ZOOM_HERMITE = 1,
    TCV_ZOOM_BOX,
    TCV_ZOOM_TRIANGLE,
    TCV_ZOOM_BELL,
    TCV_ZOOM_B_SPLINE,
    TCV_ZOOM_LANCZOS3,
    TCV_ZOOM_MITCHELL,
    TCV_ZOOM_CUBIC_KEYS4,
    TCV_ZOOM_SINC8,
    TCV_ZOOM_NULL
} TCVZoomFilter;

TCVHandle tcv_init(void);                 /* 0 when no usable device: there is no CPU implementation behind it */
void tcv_free(TCVHandle handle);

int tcv_clip(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
             int clip_left, int clip_right, int clip_top, int clip_bottom, uint8_t black_pixel);
int tcv_deinterlace(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp, TCVDeinterlaceMode mode);
int tcv_resize(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
               int resize_w, int resize_h, int scale_w, int scale_h);
int tcv_reduce(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int reduce_w, int reduce_h);
int tcv_flip_v(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp);
int tcv_flip_h(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp);
int tcv_gamma_correct(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double gamma);
int tcv_antialias(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double weight, double bias);
int tcv_convert(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, ImageFormat srcfmt, ImageFormat destfmt);

/* The filtered resampler behind -Z and -I 3 (libtcvideo/zoom.c) is not a caller of the aclib path and has no device twin:
 * tcv_zoom returns 0 and says so.  A transcode build that needs -Z keeps the reference's zoom.c and the tcv_zoom wrapper
 * (tcvideo.c:558-652) as a source file of its own next to libtcvgpu. */
int tcv_zoom(TCVHandle handle, uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
             int new_w, int new_h, TCVZoomFilter filter);
const char *tcv_zoom_filter_to_string(TCVZoomFilter filter);        /* names as libtcvideo/zoom.c:79-105 */
TCVZoomFilter tcv_zoom_filter_from_string(const char *name);        /* case-insensitive, TCV_ZOOM_NULL if unknown */

#ifdef __cplusplus
}
#endif
#endif /* LIBTCVGPU_TCVIDEO_H */
