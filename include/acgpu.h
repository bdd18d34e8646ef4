/*
 * include/acgpu.h -- entry points that libacgpu ADDS to the aclib interface.
 *
 * The reference has no equivalent of these: aclib processes one host frame (or one line) per call
 * (aclib/imgconvert.c:34-64, aclib/rescale.c:23-32, aclib/average.c:22-26).  A GPU needs many
 * frames per launch and data that stays in HBM, so libacgpu offers
 *   - batched, device-resident forms of ac_imgconvert / ac_rescale / ac_average, and
 *   - frame-granular forms of their direct libtcvideo callers, tcv_deinterlace
 *     (libtcvideo/tcvideo.c:290-389) and tcv_resize (libtcvideo/tcvideo.c:427-531),
 *   - the small amount of device plumbing a C caller needs (device selection, memory, streams, events).
 * Plain C ABI: pointers, sizes and opaque handles only.
 *
 * Conventions: functions returning int give 1 on success and 0 on failure (aclib's convention,
 * aclib/ac.h:56-57); acgpu_last_error() then describes the failure for the calling thread.
 * Every call acts on the calling thread's current device (acgpu_set_device; default 0, or
 * $ACGPU_DEVICE).  `stream` may be NULL = the calling thread's private stream.  Batched calls are
 * asynchronous with respect to the host; use acgpu_stream_sync() or events.
 *
 * Ordering: a batched call given HOST memory on either side runs its upload / kernels / download on the thread's
 * private stream, ordered after everything queued so far on the stream the caller named, and returns when the result
 * has landed.  The legacy aclib calls (ac_imgconvert, ac_average, ac_rescale, ac_memcpy) have no stream argument: with
 * device pointers they order only against the thread's private stream -- synchronise other streams first.
 * ac_memcpy treats every pointer as host memory until the process has called an acgpu_* function that can hand out or
 * select device memory (acgpu_malloc, acgpu_set_device, acgpu_stream_create, any batched call): unmodified callers pay no
 * driver query per copy.
 * Pitches: a frame pitch of 0 means tightly packed frames; a pitch smaller than a frame is rejected when nframes > 1
 * (acgpu_imgconvert_batch and acgpu_rowops_run take separate plane / row offsets and need explicit pitches).
 * In place: src == dest is accepted where the reference's sequential result is well defined and equal to the out-of-place
 * one -- flips, gamma, conversion (through a temporary, tcvideo.c:1044-1064), cropping clip, reduce, drop-field and
 * interpolating deinterlace, shrinking resize -- and rejected elsewhere (growing clip, enlarging resize, linear blend,
 * antialias); partially overlapping src / dest are always rejected (tcvideo.c:180 "src and dest do not overlap").
 */
#ifndef ACGPU_H
#define ACGPU_H

#include <stddef.h>
#include <stdint.h>

#include "ac.h"
#include "imgconvert.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct acgpu_stream_s *acgpu_stream_t;   /* a CUDA stream */
typedef struct acgpu_event_s  *acgpu_event_t;    /* a CUDA event  */

/* ---- library / device ------------------------------------------------------------------- */
const char *acgpu_version(void);
const char *acgpu_last_error(void);
int  acgpu_device_count(void);
int  acgpu_set_device(int ordinal);
int  acgpu_get_device(void);
int  acgpu_device_sm_count(void);
/* Which kernel tier served the calling thread's last conversion: 0 none, 1 generic (any size or
 * alignment), 2 vectorised (16-byte aligned planes), 3 TMA-staged.  For tests and profiling. */
int  acgpu_last_kernel_tier(void);
/* Number of kernel launches issued by the calling thread since the last reset (bench.py's
 * `gpu_launches`). */
uint64_t acgpu_launch_count(int reset);
/* Force a tier for ac_imgconvert / acgpu_imgconvert_batch on this thread: 0 = automatic. */
void acgpu_force_tier(int tier);

/* ---- memory, streams, events ------------------------------------------------------------ */
void *acgpu_malloc(size_t bytes);                 /* device memory, 256-byte aligned */
void  acgpu_free(void *dptr);
void *acgpu_host_alloc(size_t bytes);             /* page-locked host memory */
void  acgpu_host_free(void *hptr);
int   acgpu_memcpy_h2d(void *dptr, const void *hptr, size_t bytes, acgpu_stream_t stream);
int   acgpu_memcpy_d2h(void *hptr, const void *dptr, size_t bytes, acgpu_stream_t stream);
int   acgpu_memcpy_d2d(void *dptr, const void *sptr, size_t bytes, acgpu_stream_t stream);
int   acgpu_memset(void *dptr, int value, size_t bytes, acgpu_stream_t stream);

/* Frame-buffer plumbing.  transcode allocates every frame buffer with tc_bufalloc / tc_buffree (libtcutil/memutils.c:89-123:
 * page-aligned malloc; libtc/tcframes.c:214-229 calls it for vframe_list_t.internal_video_buf_0 / _1).  acgpu_bufalloc /
 * acgpu_buffree keep that contract on PAGE-LOCKED memory, so the unmodified per-frame ac_* / tcv_* calls on such frames DMA
 * straight from and to them (INTEGRATION.md section 4 shows the two-line replacement).  A buffer that already exists is
 * page-locked once with acgpu_host_register (whole pages; undo with acgpu_host_unregister before freeing it). */
void *acgpu_bufalloc(size_t size);
void  acgpu_buffree(void *ptr);
int   acgpu_host_register(void *ptr, size_t size);
int   acgpu_host_unregister(void *ptr);
int   acgpu_pointer_kind(const void *ptr);        /* 0 pageable host, 1 page-locked host, 2 device */

acgpu_stream_t acgpu_stream_create(void);
void  acgpu_stream_destroy(acgpu_stream_t s);
int   acgpu_stream_sync(acgpu_stream_t s);
acgpu_event_t acgpu_event_create(void);
void  acgpu_event_destroy(acgpu_event_t e);
int   acgpu_event_record(acgpu_event_t e, acgpu_stream_t s);
int   acgpu_event_sync(acgpu_event_t e);
float acgpu_event_elapsed_ms(acgpu_event_t start, acgpu_event_t stop);

/* ---- batched ac_imgconvert --------------------------------------------------------------- */
/*
 * Converts `nframes` frames in one go.  src[p] / dest[p] are DEVICE pointers to plane p of frame 0
 * (planes[0] only for packed formats, exactly as ac_imgconvert); plane p of frame k lives
 * src_frame_pitch / dest_frame_pitch bytes further per frame.  Arithmetic, untouched bytes and the
 * YV12 alias are those of ac_imgconvert (aclib/imgconvert.c:34-64).  Unlike the reference's
 * UYVY/YVYU -> planar wrapper (aclib/img_yuv_mixed.c:24-36) the source is never modified.
 */
int acgpu_imgconvert_batch(uint8_t *const *src, ImageFormat srcfmt, size_t src_frame_pitch,
                           uint8_t *const *dest, ImageFormat destfmt, size_t dest_frame_pitch,
                           int width, int height, int nframes, acgpu_stream_t stream);

/*
 * Host-buffer form of the same: frames are tightly packed (YUV_INIT_PLANES layout) one after the
 * other in `src_frames` / `dest_frames` (best: acgpu_host_alloc memory).  Copies and kernels of
 * successive chunks are overlapped on internal streams; returns after everything has landed in
 * dest_frames.  This is the call bench.py's end-to-end number goes through.
 */
int acgpu_imgconvert_frames_host(const uint8_t *src_frames, ImageFormat srcfmt,
                                 uint8_t *dest_frames, ImageFormat destfmt,
                                 int width, int height, int nframes);
/*
 * The same over several GPUs of one host: frames are independent, so the run is shared out among `ndevices` devices
 * (devices 0 .. ndevices-1; ndevices <= 0 = every visible device) in grains of frames taken from a shared counter -- the
 * devices of a box do not all get the same share of the host's DMA bandwidth -- and each grain goes through its device's
 * pipeline on that device's internal host thread: one thread, stream set and staging per device, no exchange between devices.
 * What transcode's N frame threads do by hand (src/frame_threads.c:174-228), for callers that hold a run of frames.
 */
int acgpu_imgconvert_frames_host_multi(const uint8_t *src_frames, ImageFormat srcfmt,
                                       uint8_t *dest_frames, ImageFormat destfmt,
                                       int width, int height, int nframes, int ndevices);

/* ---- batched ac_average / ac_rescale ------------------------------------------------------ */
/*
 * One row operation: dest_row = blend(src1_row, src2_row) with ac_rescale's exact rules
 * (aclib/rescale.c:23-46: weight1 >= 65536 copies src1 and never reads src2; else weight2 >= 65536
 * copies src2; else (a*w1 + b*w2 + 32768) >> 16 in uint32, low byte kept), or ac_average's
 * (aclib/average.c:33-39) when `op` is ACGPU_ROW_AVERAGE.  Offsets are byte offsets from the frame
 * base pointers given to acgpu_rowops_run().
 */
enum { ACGPU_ROW_RESCALE = 0, ACGPU_ROW_AVERAGE = 1, ACGPU_ROW_COPY = 2, ACGPU_ROW_AVERAGE3 = 3 };
typedef struct {
    int64_t  src1_off, src2_off, src3_off, dest_off;
    uint32_t weight1, weight2;
    uint32_t op;       /* ACGPU_ROW_*; AVERAGE3 = average(src3, average(src1, src2)), the fused form of
                          tcv_deinterlace's linear blend (libtcvideo/tcvideo.c:368-389) */
    uint32_t reserved;
} acgpu_rowop;

/* Runs `nops` row operations of `row_bytes` bytes on each of `nframes` frames with one launch.
 * `ops` is a HOST array (copied to the device, cached by content).  Rows must not depend on rows
 * written by the same call. */
int acgpu_rowops_run(const uint8_t *src, size_t src_frame_pitch, uint8_t *dest, size_t dest_frame_pitch,
                     const acgpu_rowop *ops, int nops, int row_bytes, int nframes, acgpu_stream_t stream);

/* Device-pointer, single-call forms (the kernels behind the legacy ac_average / ac_rescale). */
int acgpu_average(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, size_t bytes, acgpu_stream_t stream);
int acgpu_rescale(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, size_t bytes,
                  uint32_t weight1, uint32_t weight2, acgpu_stream_t stream);

/* ---- frame-granular libtcvideo shapes ----------------------------------------------------------
 * Written for device-resident planes (asynchronous on `stream`).  src and/or dest may also be HOST memory, pageable or
 * page-locked -- what an unmodified libtcvideo caller holds: the batch is then uploaded once, processed on the device
 * and downloaded once through the calling thread's staging arena, and the call returns after the result has landed
 * (one round trip per frame instead of the one per ROW that tcv_deinterlace / tcv_resize make through ac_average /
 * ac_rescale).  The same holds for the element-wise operations further down. */
enum { ACGPU_DEINT_INTERPOLATE = 0, ACGPU_DEINT_LINEAR_BLEND = 1, ACGPU_DEINT_DROP_FIELD_TOP = 2, ACGPU_DEINT_DROP_FIELD_BOTTOM = 3 };
/* tcv_deinterlace (libtcvideo/tcvideo.c:290-389), all four modes, on nframes frames of width x height x Bpp bytes.
 * The drop-field modes write height/2 rows (tcvideo.c:326-338).  Unlike the reference's linear blend, src is left intact. */
int acgpu_deinterlace_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int mode,
                            size_t src_frame_pitch, size_t dest_frame_pitch, int nframes, acgpu_stream_t stream);
/* tcv_resize (libtcvideo/tcvideo.c:427-531): exactly one of resize_w / resize_h may be non-zero. */
int acgpu_resize_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
                       int resize_w, int resize_h, int scale_w, int scale_h,
                       size_t src_frame_pitch, size_t dest_frame_pitch, int nframes, acgpu_stream_t stream);

/* tcv_convert (libtcvideo/tcvideo.c:1001-1067) on device-resident frames: planes are laid out as YUV_INIT_PLANES
 * does, equal formats copy, and src == dest converts through an internal temporary exactly like the reference. */
int acgpu_convert_batch(uint8_t *src, uint8_t *dest, int width, int height, ImageFormat srcfmt, ImageFormat destfmt,
                        size_t src_frame_pitch, size_t dest_frame_pitch, int nframes, acgpu_stream_t stream);
/* transcode's -K on RGB24 frames (src/video_trans.c:381-388: RGB24 -> GRAY8 -> RGB24) fused into one in-place pass;
 * byte-identical to the two ac_imgconvert calls, half the memory traffic. */
int acgpu_decolor_rgb24_batch(uint8_t *frames, int width, int height, size_t frame_pitch, int nframes,
                              acgpu_stream_t stream);

/* ---- the remaining element-wise libtcvideo plane operations (device-resident) --------------------------------
 * Same arguments, checks and return values (1 = success, 0 = rejected) as the reference functions they replace; one
 * plane of width x height pixels with Bpp 1 or 3, nframes planes a fixed pitch apart, one launch per call. */
/* tcv_clip (libtcvideo/tcvideo.c:184-250): negative clips grow the frame with black_pixel.  dest holds
 * (width-clip_left-clip_right) x (height-clip_top-clip_bottom) pixels. */
int acgpu_clip_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
                     int clip_left, int clip_right, int clip_top, int clip_bottom, uint8_t black_pixel,
                     size_t src_frame_pitch, size_t dest_frame_pitch, int nframes, acgpu_stream_t stream);
/* tcv_reduce (tcvideo.c:681-717): keeps every reduce_w-th pixel of every reduce_h-th row. */
int acgpu_reduce_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int reduce_w, int reduce_h,
                       size_t src_frame_pitch, size_t dest_frame_pitch, int nframes, acgpu_stream_t stream);
/* tcv_flip_v / tcv_flip_h (tcvideo.c:739-766, 787-818); src == dest flips in place, as in the reference. */
int acgpu_flip_v_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
                       size_t src_frame_pitch, size_t dest_frame_pitch, int nframes, acgpu_stream_t stream);
int acgpu_flip_h_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
                       size_t src_frame_pitch, size_t dest_frame_pitch, int nframes, acgpu_stream_t stream);
/* tcv_gamma_correct (tcvideo.c:840-858): the 256-entry table is built on the host exactly as :1180-1189 does. */
int acgpu_gamma_correct_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double gamma,
                              size_t src_frame_pitch, size_t dest_frame_pitch, int nframes, acgpu_stream_t stream);
/* tcv_antialias (tcvideo.c:886-980), weight tables as :1209-1224; src and dest must not overlap. */
int acgpu_antialias_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double weight, double bias,
                          size_t src_frame_pitch, size_t dest_frame_pitch, int nframes, acgpu_stream_t stream);

/* ---- frame chains: several operations per PCIe round trip -----------------------------------------------------------
 * transcode applies a list of operations to every frame (do_process_frame, src/video_trans.c:192-426: clip, deinterlace,
 * resize, clip, reduce, flip, mirror, -k, -K, gamma, antialias) and its filter wrappers convert yuv -> rgb -> filter -> yuv
 * (filter/filter_ascii.c:367-373).  A chain runs such a list on the device copy of a frame: one upload, N kernels, one
 * download.  Frames are whole images in transcode's layouts (tightly packed, YUV_INIT_PLANES); every stage other than
 * CONVERT needs a YUV420P, YUV422P or RGB24 frame -- the layouts do_process_frame knows (video_trans.c:85-103); Y8 / GRAY8 pass
 * as its default single 8-bit plane (video_trans.c:77-84; no -k / -K there) -- and treats its planes as the reference does: PROCESS_FRAME stages run on every plane with the chroma planes' divided sizes and
 * arguments, the first-plane-only stages (-I 1/5, gamma, antialias) leave chroma alone.  Results are byte-identical to the
 * same libtcvideo / ac_imgconvert calls made one after the other on the host. */
enum {
    ACGPU_CHAIN_CONVERT = 1,   /* tcv_convert (tcvideo.c:1001-1067): p[0] = destination ImageFormat, any of the 225 pairs   */
    ACGPU_CHAIN_CLIP,          /* -j / -Y (video_trans.c:213-223, 327-337): p[0..3] = left, right, top, bottom in frame pixels
                                  (negative grows the frame with black: 0 / chroma 128); multiples of the chroma subsampling */
    ACGPU_CHAIN_DEINTERLACE,   /* -I (video_trans.c:227-279): p[0] = 1 interpolate, 5 linear blend (first plane), 4 drop field
                                  (all planes, half height), 2 nothing (left to the encoder); 3 needs tcv_zoom: rejected       */
    ACGPU_CHAIN_RESIZE,        /* -B / -X (video_trans.c:281-297): p[0] = resize_w, p[1] = resize_h in units of 8 pixels, rows first */
    ACGPU_CHAIN_REDUCE,        /* -r (video_trans.c:341-345): p[0] = reduce_w, p[1] = reduce_h                                */
    ACGPU_CHAIN_FLIP_V,        /* -z */
    ACGPU_CHAIN_FLIP_H,        /* -l */
    ACGPU_CHAIN_RGBSWAP,       /* -k (video_trans.c:349-366): R <-> B, or U <-> V planes                                      */
    ACGPU_CHAIN_DECOLOR,       /* -K (video_trans.c:370-386): RGB24 -> GRAY8 -> RGB24, or chroma planes = 128                 */
    ACGPU_CHAIN_GAMMA,         /* -G (video_trans.c:390-396): d[0] = gamma, first plane only                                  */
    ACGPU_CHAIN_ANTIALIAS      /* -C (video_trans.c:400-421): d[0] = weight, d[1] = bias, first plane only                    */
};
typedef struct {
    int32_t kind;      /* ACGPU_CHAIN_* */
    int32_t p[5];
    double  d[2];
} acgpu_chain_op;

/* Layout of the frames a chain produces (checks the whole list; 0 + acgpu_last_error() if a stage cannot apply). */
int acgpu_chain_output(ImageFormat fmt, int width, int height, const acgpu_chain_op *ops, int nops,
                       ImageFormat *out_fmt, int *out_width, int *out_height);
/* Device-resident: `nframes` frames a fixed pitch apart (0 = tightly packed), src is not modified, src and dest must not
 * overlap.  Intermediate frames live in the calling thread's temporary (at most $ACGPU_CHAIN_SCRATCH_BYTES, default 2 GiB;
 * longer batches are walked in sub-batches that fit).  Asynchronous on `stream`. */
int acgpu_chain_batch(const uint8_t *src, ImageFormat fmt, int width, int height, size_t src_frame_pitch,
                      uint8_t *dest, size_t dest_frame_pitch, const acgpu_chain_op *ops, int nops, int nframes,
                      acgpu_stream_t stream);
/* Host frames, tightly packed one after the other (best: acgpu_host_alloc memory): chunks of frames go through a four-slot
 * upload / chain / download pipeline; returns after everything has landed in dest_frames.  The _multi form shares the run
 * out among the devices in grains of frames, like acgpu_imgconvert_frames_host_multi. */
int acgpu_chain_frames_host(const uint8_t *src_frames, ImageFormat fmt, int width, int height, uint8_t *dest_frames,
                            const acgpu_chain_op *ops, int nops, int nframes);
int acgpu_chain_frames_host_multi(const uint8_t *src_frames, ImageFormat fmt, int width, int height, uint8_t *dest_frames,
                                  const acgpu_chain_op *ops, int nops, int nframes, int ndevices);
/* The same for frames that each live in a buffer of their own -- transcode's frame ring: one tc_bufalloc'd video_buf per
 * vframe_list_t (tccore/frame.h:215-253, libtc/tcframes.c:214-229) -- or at any stride inside a larger buffer (a YUV4MPEG2
 * stream under construction: "FRAME\n" + planes, encode/encode_yuv4mpeg.c:256-289).  src_frames[i] / dest_frames[i] point
 * at frame i (YUV_INIT_PLANES layout); dest_frames[i] == src_frames[i] processes frame i in place as do_process_frame does
 * (the frame is uploaded before its result comes back), any other overlap between the two lists is the caller's to avoid.
 * One copy per frame and direction instead of one per chunk; everything else is the pipeline of acgpu_chain_frames_host. */
int acgpu_chain_frame_list_host(const uint8_t *const *src_frames, ImageFormat fmt, int width, int height,
                                uint8_t *const *dest_frames, const acgpu_chain_op *ops, int nops, int nframes);
int acgpu_chain_frame_list_host_multi(const uint8_t *const *src_frames, ImageFormat fmt, int width, int height,
                                      uint8_t *const *dest_frames, const acgpu_chain_op *ops, int nops, int nframes, int ndevices);

/* Optional: stops the per-device host threads of the *_multi calls while the CUDA runtime is still alive.  Without it
 * they are abandoned at process exit (never joined from a static destructor). */
void acgpu_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* ACGPU_H */
