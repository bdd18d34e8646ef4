/*
 * frame_caller.c -- transcode's video frame buffers on libacgpu's allocator.
 *
 * The caller below is what libtc/tcframes.c does, in miniature: a frame owns two page-aligned buffers from the buffer
 * allocator (tc_alloc_video_frame, :214-229: internal_video_buf_0 / _1 via tc_bufalloc), its plane pointers are carved
 * out of them (tc_init_video_frame, :106-130), and frame operations ping-pong between the two (src/video_trans.c:130-150).
 * The ONLY change from the reference is the allocator: tc_bufalloc / tc_buffree are defined as acgpu_bufalloc /
 * acgpu_buffree, the two-line patch INTEGRATION.md proposes for libtcutil/memutils.h.  Every legacy call on such a frame
 * then takes the page-locked path.  A third buffer shows the other route: an existing malloc'ed frame registered once.
 * Last, a segment of the frame ring -- four frames, each in its own buffers -- goes through ONE frame-list chain call in
 * place, as do_process_frame works on ptr->video_buf (-I 5 -G 0.8: src/video_trans.c:267-277, 390-396).
 * Prints the pointer kind of each buffer and digests of the results; tests/test_frame_plumbing.py checks both.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ac.h"
#include "acgpu.h"
#include "imgconvert.h"
#include "tcvideo.h"

#define tc_bufalloc(size) acgpu_bufalloc(size)      /* libtcutil/memutils.h: the whole integration */
#define tc_buffree(ptr)   acgpu_buffree(ptr)

typedef struct {                /* the buffer-related fields of vframe_list_t (tccore/frame.h:215-253) */
    int v_width, v_height, video_size, free;
    uint8_t *video_buf, *video_buf2;
    uint8_t *internal_video_buf_0, *internal_video_buf_1;
    uint8_t *video_buf_Y[2], *video_buf_U[2], *video_buf_V[2];
} frame_t;

static int frame_alloc(frame_t *f, int w, int h)      /* tc_alloc_video_frame + tc_init_video_frame for an RGB-sized frame */
{
    const size_t size = (size_t)w * h * 3 + 128;      /* TC_FRAME_EXTRA_SIZE */
    int i;
    memset(f, 0, sizeof(*f));
    f->internal_video_buf_0 = tc_bufalloc(size);
    f->internal_video_buf_1 = tc_bufalloc(size);
    if (!f->internal_video_buf_0 || !f->internal_video_buf_1) return 0;
    for (i = 0; i < 2; i++) {
        f->video_buf_Y[i] = i ? f->internal_video_buf_1 : f->internal_video_buf_0;
        f->video_buf_U[i] = f->video_buf_Y[i] + (size_t)w * h;
        f->video_buf_V[i] = f->video_buf_U[i] + (size_t)(w / 2) * (h / 2);
    }
    f->video_buf = f->internal_video_buf_0;
    f->video_buf2 = f->internal_video_buf_1;
    f->free = 1;
    f->v_width = w; f->v_height = h; f->video_size = (int)size;
    return 1;
}

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
    for (i = 0; i < n; i++) { z = z * 6364136223846793005ull + 1442695040888963407ull; p[i] = (uint8_t)(z >> 56); }
}

int main(int argc, char **argv)
{
    const int w = argc > 2 ? atoi(argv[1]) : 720, h = argc > 2 ? atoi(argv[2]) : 576;
    frame_t f;
    TCVHandle tcv;
    uint8_t *src[3], *dst[3], *plain;
    if (!ac_init(AC_ALL & ac_cpuinfo())) return 1;
    if (!(tcv = tcv_init()) || !frame_alloc(&f, w, h)) return 1;
    printf("page_aligned %d %d\n", (int)((size_t)f.internal_video_buf_0 % 4096 == 0), (int)((size_t)f.internal_video_buf_1 % 4096 == 0));
    printf("kind %d %d\n", acgpu_pointer_kind(f.video_buf), acgpu_pointer_kind(f.video_buf_V[1]));

    /* decoder -> YUV420P in buffer 0; ac_imgconvert into buffer 1 through the frame's own plane pointers (decode_lavc.c:307-312) */
    fill(f.video_buf, (size_t)w * h * 3 / 2, 5);
    src[0] = f.video_buf_Y[0]; src[1] = f.video_buf_U[0]; src[2] = f.video_buf_V[0];
    dst[0] = f.video_buf_Y[1]; dst[1] = dst[2] = NULL;
    if (!ac_imgconvert(src, IMG_YUV420P, dst, IMG_RGB24, w, h)) return 1;
    printf("yuv420p_rgb24 %016llx\n", fnv(f.internal_video_buf_1, (size_t)w * h * 3));
    /* a frame operation between the two buffers, then swap (src/video_trans.c:130-150) */
    if (!tcv_flip_v(tcv, f.internal_video_buf_1, f.internal_video_buf_0, w, h, 3)) return 1;
    printf("flip_v %016llx\n", fnv(f.internal_video_buf_0, (size_t)w * h * 3));
    if (!tcv_convert(tcv, f.internal_video_buf_0, f.internal_video_buf_1, w, h, IMG_RGB24, IMG_YUV422P)) return 1;
    printf("rgb24_yuv422p %016llx\n", fnv(f.internal_video_buf_1, (size_t)w * h * 2));

    /* an existing pageable frame: registered once, then used like any other */
    plain = malloc((size_t)w * h * 3 + 8192);
    printf("plain_kind_before %d\n", acgpu_pointer_kind(plain));
    if (!acgpu_host_register(plain, (size_t)w * h * 3)) return 1;
    printf("plain_kind_after %d\n", acgpu_pointer_kind(plain + 100));
    dst[0] = plain;
    if (!ac_imgconvert(src, IMG_YUV420P, dst, IMG_BGR24, w, h)) return 1;
    printf("yuv420p_bgr24 %016llx\n", fnv(plain, (size_t)w * h * 3));
    if (!acgpu_host_unregister(plain)) return 1;
    printf("plain_kind_end %d\n", acgpu_pointer_kind(plain));
    free(plain);
    tc_buffree(f.internal_video_buf_0);
    tc_buffree(f.internal_video_buf_1);

    /* a segment of the frame ring through one chain call: every frame has buffers of its own, results replace the frames */
    {
        enum { RING = 4 };
        frame_t ring[RING];
        const uint8_t *in[RING];
        uint8_t *out[RING];
        acgpu_chain_op ops[2];
        int i;
        memset(ops, 0, sizeof(ops));
        ops[0].kind = ACGPU_CHAIN_DEINTERLACE; ops[0].p[0] = 5;
        ops[1].kind = ACGPU_CHAIN_GAMMA; ops[1].d[0] = 0.8;
        for (i = 0; i < RING; i++) {
            if (!frame_alloc(&ring[i], w, h)) return 1;
            fill(ring[i].video_buf, (size_t)w * h * 3 / 2, 10 + (unsigned)i);
            in[i] = out[i] = ring[i].video_buf;
        }
        if (!acgpu_chain_frame_list_host(in, IMG_YUV420P, w, h, out, ops, 2, RING)) { fprintf(stderr, "%s\n", acgpu_last_error()); return 1; }
        for (i = 0; i < RING; i++) {
            printf("ring_%d %016llx\n", i, fnv(ring[i].video_buf, (size_t)w * h * 3 / 2));
            tc_buffree(ring[i].internal_video_buf_0);
            tc_buffree(ring[i].internal_video_buf_1);
        }
    }
    tcv_free(tcv);
    return 0;
}
