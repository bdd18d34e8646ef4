/*
 * include/imgconvert.h -- libacgpu's copy of the aclib image-conversion interface.
 *
 * Drop-in for aclib/imgconvert.h:16-90: same ImageFormat ids, helper macros and the two entry
 * points.  ac_imgconvert() accepts host OR device plane pointers (classified per call with
 * cudaPointerGetAttributes): host planes are staged through per-thread pinned buffers, device planes
 * are converted in place on the GPU with no copies.  All 15x15 pairs of aclib/imgconvert.c's table
 * plus the YV12 plane-swap alias (imgconvert.c:40-56) are implemented; an unknown pair returns 0
 * exactly like imgconvert.c:63.
 */
#ifndef ACGPU_IMGCONVERT_H
#define ACGPU_IMGCONVERT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    IMG_UNKNOWN  = 0,
    IMG_YUV_BASE = 0x1000,
    IMG_YUV420P  = 0x1001,  /* planar, chroma 2x2 subsampled            */
    IMG_YV12     = 0x1002,  /* YUV420P with the U and V planes exchanged */
    IMG_YUV411P  = 0x1003,  /* planar, chroma 4x1 subsampled            */
    IMG_YUV422P  = 0x1004,  /* planar, chroma 2x1 subsampled            */
    IMG_YUV444P  = 0x1005,  /* planar, full-resolution chroma           */
    IMG_YUY2     = 0x1006,  /* packed Y U Y V                           */
    IMG_UYVY     = 0x1007,  /* packed U Y V Y                           */
    IMG_YVYU     = 0x1008,  /* packed Y V Y U                           */
    IMG_Y8       = 0x1009,  /* luma only                                */
    IMG_YUV_LAST = 0x100A,
    IMG_RGB_BASE = 0x2000,
    IMG_RGB24    = 0x2001,
    IMG_BGR24    = 0x2002,
    IMG_RGBA32   = 0x2003,
    IMG_ABGR32   = 0x2004,
    IMG_ARGB32   = 0x2005,
    IMG_BGRA32   = 0x2006,
    IMG_GRAY8    = 0x2007,
    IMG_RGB_LAST = 0x2008
} ImageFormat;

#define IMG_NONE        IMG_UNKNOWN
#define IMG_YUV_DEFAULT IMG_YUV420P
#define IMG_RGB_DEFAULT IMG_RGB24

#define IS_YUV_FORMAT(f) ((f) > IMG_YUV_BASE && (f) < IMG_YUV_LAST)
#define IS_RGB_FORMAT(f) ((f) > IMG_RGB_BASE && (f) < IMG_RGB_LAST)

/* Chroma plane bytes of a planar YUV frame; the luma plane is always w*h (aclib/imgconvert.h:54-59). */
#define UV_PLANE_SIZE(f, w, h)                                            \
    (((f) == IMG_YUV420P || (f) == IMG_YV12) ? ((w) / 2) * ((h) / 2)      \
     : (f) == IMG_YUV411P                    ? ((w) / 4) * (h)            \
     : (f) == IMG_YUV422P                    ? ((w) / 2) * (h)            \
     : (f) == IMG_YUV444P                    ? (w) * (h)                  \
                                             : 0)

/* Plane pointers of one tightly packed frame buffer (aclib/imgconvert.h:62-65). */
#define YUV_INIT_PLANES(planes, buffer, f, w, h)                          \
    ((planes)[0] = (buffer),                                              \
     (planes)[1] = (planes)[0] + (w) * (h),                               \
     (planes)[2] = (planes)[1] + UV_PLANE_SIZE((f), (w), (h)))

/* 1 on success, 0 on failure. */
int ac_imgconvert_init(int accel);
int ac_imgconvert(uint8_t **src, ImageFormat srcfmt, uint8_t **dest, ImageFormat destfmt,
                  int width, int height);

#ifdef __cplusplus
}
#endif
#endif /* ACGPU_IMGCONVERT_H */
