/*
 * legacy_bench.c -- the UNMODIFIED aclib call sequence, from C threads, against libacgpu (profiling aid).
 *
 * What a transcode frame thread does (src/frame_threads.c:174-228 -> libtcvideo/tcvideo.c:1060): one
 * ac_imgconvert(src planes, IMG_YUV420P, dest planes, IMG_RGB24, w, h) per frame on its own frame buffers.
 * T threads, each with its own source/destination frame, buffers either malloc'ed (pageable) or from
 * acgpu_host_alloc (page-locked: SURVEY 8f row 4, frame buffers allocated through the library).
 *
 *   gcc -O2 -I include -o /tmp/legacy_bench tools/legacy_bench.c -L transcode-tcforge_b200 -lacgpu -lpthread \
 *       -Wl,-rpath,$PWD/transcode-tcforge_b200
 *   /tmp/legacy_bench [width height [seconds]]
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ac.h"
#include "acgpu.h"
#include "imgconvert.h"

static int W = 1920, H = 1080, pinned;
static double seconds = 1.5;
static volatile int stop_flag;

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static void *worker(void *arg)
{
    long *count = arg;
    const size_t sb = (size_t)W * H * 3 / 2, db = (size_t)W * H * 3;
    uint8_t *s = pinned ? acgpu_host_alloc(sb) : malloc(sb), *d = pinned ? acgpu_host_alloc(db) : malloc(db);
    uint8_t *sp[3], *dp[3] = {d, NULL, NULL};
    size_t i;
    if (!s || !d) { fprintf(stderr, "allocation failed: %s\n", acgpu_last_error()); exit(1); }
    for (i = 0; i < sb; i++) s[i] = (uint8_t)(i * 2654435761u >> 13);
    memset(d, 0, db);
    YUV_INIT_PLANES(sp, s, IMG_YUV420P, W, H);
    while (!stop_flag) {
        if (!ac_imgconvert(sp, IMG_YUV420P, dp, IMG_RGB24, W, H)) { fprintf(stderr, "ac_imgconvert failed: %s\n", acgpu_last_error()); exit(1); }
        ++*count;
    }
    if (pinned) { acgpu_host_free(s); acgpu_host_free(d); } else { free(s); free(d); }
    return NULL;
}

int main(int argc, char **argv)
{
    static const int threads[] = {1, 2, 4, 8, 16};
    int t, i;
    if (argc > 2) { W = atoi(argv[1]); H = atoi(argv[2]); }
    if (argc > 3) seconds = atof(argv[3]);
    if (!ac_init(AC_ALL)) { fprintf(stderr, "ac_init failed: %s\n", acgpu_last_error()); return 1; }
    for (pinned = 0; pinned <= 1; pinned++)
        for (t = 0; t < 5; t++) {
            const int T = threads[t];
            pthread_t tid[16];
            long counts[16] = {0}, total = 0;
            double t0;
            stop_flag = 0;
            for (i = 0; i < T; i++) pthread_create(&tid[i], NULL, worker, &counts[i]);
            {   /* let every thread finish its first (context-creating) call before the clock starts */
                struct timespec w = {0, 300000000};
                nanosleep(&w, NULL);
            }
            for (i = 0; i < T; i++) total -= counts[i];
            t0 = now();
            while (now() - t0 < seconds) { struct timespec w = {0, 2000000}; nanosleep(&w, NULL); }
            for (i = 0; i < T; i++) total += counts[i];
            t0 = now() - t0;
            stop_flag = 1;
            for (i = 0; i < T; i++) pthread_join(tid[i], NULL);
            printf("ac_imgconvert %dx%d YUV420P->RGB24, %-8s host frames, %2d C threads: %8.1f frames/s\n", W, H,
                   pinned ? "pinned" : "pageable", T, total / t0);
            fflush(stdout);
        }
    return 0;
}
